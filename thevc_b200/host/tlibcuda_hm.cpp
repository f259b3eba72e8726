/*
 * tlibcuda_hm.cpp -- HM 7.2 side of the C ABI (see tlibcuda_hm.h).  Host C++; every computation is a
 * call into libthevc_cuda.so (include/thevc_cuda.h).
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

#define private public
#define protected public
#include "TLibCommon/TComRdCost.h"
#undef private
#undef protected
#include "TLibCommon/TComRom.h"
#include "TLibCommon/ContextTables.h"
#include "TLibCommon/TComPattern.h"
#include "TLibCommon/TComDataCU.h"
#include "TLibCommon/TComPic.h"
#include "TLibCommon/TComPicYuv.h"
#include "TLibCommon/TComSlice.h"
#include "TLibCommon/TComMv.h"
#include "TLibCommon/TComYuv.h"
#include "TLibEncoder/TEncCfg.h"

#include "tlibcuda_hm.h"
#include "thevc_cuda.h"

namespace {

struct DevPic {                 // one registered TComPicYuv: host buffer range -> device slot
  TComPicYuv* yuv = nullptr;
  int poc = -0x7fffffff;
  bool is_org = false;
  unsigned long long stamp = 0;
};

struct State {
  tvc_ctx* h = nullptr;
  bool on_me = true, on_frac = true, on_tq = true, on_rdoq = true, on_mc = true, on_tables = true, verbose = false, disabled = false;
  int w = 0, ht = 0;
  std::vector<DevPic> slots;
  unsigned long long clock = 0;
  int cur_slot = -1;
  int table_refs[8];
  int num_table_refs = 0;
  unsigned long long n_tz = 0, n_frac = 0, n_xt = 0, n_xit = 0, n_dq = 0, n_mc = 0, n_rdoq = 0;
};

State& S()
{
  static State s;
  return s;
}

void die(const char* what, int rc)
{
  fprintf(stderr, "TLibCuda: %s failed (%d): %s\n", what, rc, S().h ? tvc_last_error(S().h) : "no context");
  exit(EXIT_FAILURE);
}
#define CK(call)                      \
  do {                                \
    int _rc = (call);                 \
    if (_rc != TVC_OK) die(#call, _rc); \
  } while (0)

void report()
{
  State& s = S();
  if (s.h)
    fprintf(stderr, "TLibCuda: %llu xTZSearch, %llu xPatternSearchFracDIF, %llu xT, %llu xIT, %llu xDeQuant, %llu xRateDistOptQuant, %llu xPredInterUni calls served; %llu kernel launches\n",
            s.n_tz, s.n_frac, s.n_xt, s.n_xit, s.n_dq, s.n_rdoq, s.n_mc, (unsigned long long)tvc_launch_count(s.h));
}

void parse_env()
{
  State& s = S();
  const char* e = getenv("TVC_HM");
  if (!e) return;
  s.on_me = strstr(e, "me") != nullptr;
  s.on_frac = strstr(e, "frac") != nullptr;
  s.on_tq = strstr(e, "tq") != nullptr;
  s.on_rdoq = strstr(e, "rdoq") != nullptr;
  s.on_mc = strstr(e, "mc") != nullptr;
  s.on_tables = strstr(e, "tables") != nullptr;
  s.verbose = strstr(e, "verbose") != nullptr;
}

void ensure_ctx(int w, int ht)
{
  State& s = S();
  if (s.disabled || (s.h && s.w >= w && s.ht >= ht)) return;
  static bool parsed = false;
  if (!parsed) { parse_env(); parsed = true; atexit(report); }
  if (!s.on_me && !s.on_frac && !s.on_tq && !s.on_rdoq && !s.on_mc) { s.disabled = true; return; }     // TVC_HM=none: the unmodified path
  if (s.h) {                  // the decoder learns the picture size after its first transforms: start over with the real size
    tvc_ctx_destroy(s.h);
    s.h = nullptr;
  }
  tvc_config c;
  c.width = w; c.height = ht;
  c.bit_depth = (int)(g_uiBitDepth + g_uiBitIncrement);
  c.max_cu = (int)g_uiMaxCUWidth;
  c.num_slots = 24;
  c.device = 0;
  int rc = tvc_ctx_create(&c, &s.h);
  if (rc != TVC_OK) {
    fprintf(stderr, "TLibCuda: tvc_ctx_create failed (%d): no CUDA device -- there is no CPU fallback\n", rc);
    exit(EXIT_FAILURE);
  }
  s.w = w; s.ht = ht;
  s.slots.assign(c.num_slots, DevPic());
  s.cur_slot = -1; s.num_table_refs = 0;
}

// slot of a picture buffer; uploads it when the slot does not hold this picture's current content
int slot_for(TComPicYuv* yuv, int poc, bool is_org, bool with_margin)
{
  State& s = S();
  int free_slot = -1;
  unsigned long long oldest = ~0ull;
  for (size_t i = 0; i < s.slots.size(); i++) {
    DevPic& d = s.slots[i];
    if (d.yuv == yuv && d.poc == poc && d.is_org == is_org) { d.stamp = ++s.clock; return (int)i; }
  }
  for (size_t i = 0; i < s.slots.size(); i++)
    if (s.slots[i].stamp < oldest) { oldest = s.slots[i].stamp; free_slot = (int)i; }
  DevPic& d = s.slots[free_slot];
  d.yuv = yuv; d.poc = poc; d.is_org = is_org; d.stamp = ++s.clock;
  CK(tvc_pic_upload(s.h, free_slot, yuv->getLumaAddr(), yuv->getStride(), yuv->getCbAddr(), yuv->getCrAddr(), yuv->getCStride(),
                    with_margin ? 1 : 0));
  return free_slot;
}

// which registered reconstruction does this luma pointer point into, and at which pel
bool locate(const short* p, int& slot, int& x, int& y)
{
  State& s = S();
  for (size_t i = 0; i < s.slots.size(); i++) {
    DevPic& d = s.slots[i];
    if (!d.yuv || d.is_org) continue;
    const short* org = d.yuv->getLumaAddr();
    const int stride = d.yuv->getStride();
    const ptrdiff_t off = p - org;
    if (off < 0 || off >= (ptrdiff_t)stride * d.yuv->getHeight()) continue;
    y = (int)(off / stride); x = (int)(off % stride);
    if (x >= d.yuv->getWidth()) continue;
    slot = (int)i;
    return true;
  }
  return false;
}

}  // namespace

void tlibcuda_picture_start(TComPic* pic, TComSlice* slice)
{
  TComPicYuv* org = pic->getPicYuvOrg();
  ensure_ctx(org->getWidth(), org->getHeight());
  State& s = S();
  if (!s.h || (!s.on_me && !s.on_frac)) return;
  s.cur_slot = slot_for(org, slice->getPOC(), true, false);
  s.num_table_refs = 0;
  for (int l = 0; l < 2; l++) {
    RefPicList e = l ? REF_PIC_LIST_1 : REF_PIC_LIST_0;
    for (int i = 0; i < slice->getNumRefIdx(e); i++) {
      TComPic* r = slice->getRefPic(e, i);
      int slot = slot_for(r->getPicYuvRec(), r->getPOC(), false, true);
      bool seen = false;
      for (int k = 0; k < s.num_table_refs; k++) seen |= s.table_refs[k] == slot;
      if (!seen && s.num_table_refs < 8) s.table_refs[s.num_table_refs++] = slot;
    }
  }
  if (s.on_tables && s.num_table_refs > 0 && g_uiBitIncrement == 0)
    CK(tvc_me_prepass(s.h, s.cur_slot, s.num_table_refs, s.table_refs, nullptr));
  else
    s.num_table_refs = 0;
  if (s.verbose) fprintf(stderr, "TLibCuda: POC %d cur slot %d, %d reference(s) with SAD tables\n", slice->getPOC(), s.cur_slot, s.num_table_refs);
}

bool tlibcuda_tz_search(TComDataCU* cu, TComPattern* key, short* refY, int refStride, TComMv* lt, TComMv* rb, TComMv& rcMv,
                        unsigned& ruiSAD, TComRdCost* rd, TEncCfg* cfg, int searchRange)
{
  State& s = S();
  if (!s.h || !s.on_me || s.cur_slot < 0) return false;
  int slot, x, y;
  if (!locate(refY, slot, x, y)) return false;
  (void)refStride;
  tvc_me_job j;
  memset(&j, 0, sizeof(j));
  j.ref_slot = slot;
  j.ref_index = -1;
  for (int k = 0; k < s.num_table_refs; k++)
    if (s.table_refs[k] == slot) j.ref_index = k;
  j.x = x; j.y = y; j.w = key->getROIYWidth(); j.h = key->getROIYHeight();
  j.mode = TVC_ME_TZ;
  j.fen = cfg->getUseFastEnc() ? 1 : 0;
  j.search_range = searchRange;
  j.lx = lt->getHor(); j.ty = lt->getVer(); j.rx = rb->getHor(); j.by = rb->getVer();
  j.predx = rd->m_mvPredictor.getHor(); j.predy = rd->m_mvPredictor.getVer();
  TComMv start = rcMv;                      // xTZSearch :4311-4312
  cu->clipMv(start);
  start >>= 2;
  j.startx = start.getHor(); j.starty = start.getVer();
  j.lambda_cost = rd->m_uiCost;
  tvc_me_result r;
  CK(tvc_me_search_batch(s.h, s.cur_slot, j.ref_index >= 0 ? 1 : 0, 1, &j, &r));
  rcMv.set(r.mvx, r.mvy);
  ruiSAD = r.sad;
  s.n_tz++;
  return true;
}

bool tlibcuda_frac_search(TComPattern* key, short* refY, int refStride, TComMv* mvInt, TComMv& half, TComMv& qter,
                          unsigned& ruiCost, TComRdCost* rd, TEncCfg* cfg, bool biPred)
{
  State& s = S();
  if (!s.h || !s.on_frac || s.cur_slot < 0 || biPred) return false;     // bi-pred search target is not the original picture
  int slot, x, y;
  if (!locate(refY, slot, x, y)) return false;
  (void)refStride;
  tvc_frac_job j;
  memset(&j, 0, sizeof(j));
  j.ref_slot = slot;
  j.x = x; j.y = y; j.w = key->getROIYWidth(); j.h = key->getROIYHeight();
  j.imvx = mvInt->getHor(); j.imvy = mvInt->getVer();
  j.predx = rd->m_mvPredictor.getHor(); j.predy = rd->m_mvPredictor.getVer();
  j.lambda_cost = rd->m_uiCost;
  j.hadamard = cfg->getUseHADME() ? 1 : 0;
  tvc_frac_result r;
  CK(tvc_me_frac_batch(s.h, s.cur_slot, 1, &j, &r));
  half.set(r.halfx, r.halfy);
  qter.set(r.qtrx, r.qtry);
  ruiCost = r.cost;
  rd->setCostScale(0);                      // side effect of the reference body (:4505)
  s.n_frac++;
  return true;
}

bool tlibcuda_pred_inter_uni(TComDataCU* cu, TComPic* refPic, unsigned partAddr, int mvx, int mvy, int w, int h,
                             TComYuv* dst, bool bi)
{
  TComPicYuv* rec = refPic->getPicYuvRec();
  ensure_ctx(rec->getWidth(), rec->getHeight());
  State& s = S();
  if (!s.h || !s.on_mc) return false;
  const int slot = slot_for(rec, refPic->getPOC(), false, true);
  const unsigned z = cu->getZorderIdxInCU() + partAddr;
  const unsigned raster = g_auiZscanToRaster[z];
  const int ctus_x = (int)cu->getPic()->getFrameWidthInCU();
  const int x = (int)(cu->getAddr() % ctus_x) * (int)g_uiMaxCUWidth + (int)g_auiRasterToPelX[raster];
  const int y = (int)(cu->getAddr() / ctus_x) * (int)g_uiMaxCUHeight + (int)g_auiRasterToPelY[raster];
  CK(tvc_mc_block(s.h, slot, x, y, w, h, mvx, mvy, bi ? 1 : 0, dst->getLumaAddr(partAddr), (int)dst->getStride(),
                  dst->getCbAddr(partAddr), dst->getCrAddr(partAddr), (int)dst->getCStride()));
  s.n_mc++;
  return true;
}

static void ensure_tq_ctx()
{
  if (!S().h && !S().disabled) ensure_ctx(64, 64);        // transforms need no picture slots
}

bool tlibcuda_xT(unsigned mode, short* resi, unsigned stride, int* coef, int w, int h)
{
  ensure_tq_ctx();
  State& s = S();
  if (!s.h || !s.on_tq || w != h) return false;
  CK(tvc_xT(s.h, (w == 4 && mode != REG_DCT) ? 1 : 0, resi, (int)stride, coef, w, h));
  s.n_xt++;
  return true;
}

bool tlibcuda_xIT(unsigned mode, int* coef, short* resi, unsigned stride, int w, int h)
{
  ensure_tq_ctx();
  State& s = S();
  if (!s.h || !s.on_tq || w != h) return false;
  CK(tvc_xIT(s.h, (w == 4 && mode != REG_DCT) ? 1 : 0, coef, resi, (int)stride, w, h));
  s.n_xit++;
  return true;
}

bool tlibcuda_xDeQuant(const int* src, int* dst, int w, int h, int per, int rem)
{
  ensure_tq_ctx();
  State& s = S();
  if (!s.h || !s.on_tq || w != h) return false;
  CK(tvc_xDeQuant(s.h, src, dst, w, h, per, rem));
  s.n_dq++;
  return true;
}

bool tlibcuda_rdoq(TComDataCU* cu, int* src, int* dst, int* arl, unsigned w, unsigned h, unsigned& absSum, int ttype,
                   unsigned absPartIdx, int per, int rem, double lambda, const void* est, bool useArl)
{
  ensure_tq_ctx();
  State& s = S();
  if (!s.h || !s.on_rdoq || w != h) return false;
  static_assert(sizeof(tvc_est_bits) == 254 * sizeof(int), "estBitsSbacStruct layout");
  const bool luma = ttype == TEXT_LUMA;
  const bool intra = cu->isIntra(absPartIdx);
  unsigned scan = cu->getCoefScanIdx(absPartIdx, w, luma, intra);        // :1767-1772
  const int scan_idx = scan == SCAN_HOR ? 1 : (scan == SCAN_VER ? 2 : 0);
  int cbf_ctx;                                                             // :2103-2115
  if (!intra && luma && cu->getTransformIdx(absPartIdx) == 0) cbf_ctx = -1;
  else cbf_ctx = (ttype ? TEXT_CHROMA : ttype) * NUM_QT_CBF_CTX + (int)cu->getCtxQtCbf(absPartIdx, (TextType)ttype, cu->getTransformIdx(absPartIdx));
  if (!useArl) memset(arl, 0, sizeof(int) * w * h);                        // :1783
  CK(tvc_xRateDistOptQuant(s.h, src, dst, arl, (int)w, luma ? 1 : 0, scan_idx, per, rem, cbf_ctx,
                           cu->getSlice()->getPPS()->getSignHideFlag() ? 1 : 0, useArl ? 1 : 0, lambda, (const tvc_est_bits*)est, &absSum));
  s.n_rdoq++;
  return true;
}
