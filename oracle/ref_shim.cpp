/*
 * ref_shim.cpp -- extern "C" wrappers around the UNMODIFIED reference classes.
 *
 * TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile together with the reference's own
 * TLibCommon sources (taken where they lie under /root/reference) into
 * oracle/_ref/libhmref.so.  No reference source is copied into this repository; this file
 * only #includes the reference headers at build time and forwards calls, so that the
 * restatement in hm_oracle*.c and the CUDA path can be compared with what the reference's
 * own compiled code returns.  `private`/`protected` are opened for this translation unit
 * only, to reach TComTrQuant::xT/xIT/xQuant/xDeQuant, TComPrediction::xPredInter*Blk and
 * TEncSearch::xSetSearchRange / xTZSearch / xPatternSearch / xPatternSearchFracDIF.
 */
#include <sstream>
#include <cstring>
#include <cstdlib>
#include <vector>

#define private public
#define protected public
#include "TLibCommon/TypeDef.h"
#include "TLibCommon/CommonDef.h"
#include "TLibCommon/TComRom.h"
#include "TLibCommon/TComRdCost.h"
#include "TLibCommon/TComPattern.h"
#include "TLibCommon/TComInterpolationFilter.h"
#include "TLibCommon/TComTrQuant.h"
#include "TLibCommon/TComDataCU.h"
#include "TLibCommon/TComSlice.h"
#include "TLibCommon/TComPic.h"
#include "TLibCommon/TComPicYuv.h"
#include "TLibCommon/TComYuv.h"
#include "TLibCommon/TComPrediction.h"
#include "TLibCommon/TComPicYuv.h"
#include "TLibEncoder/TEncCfg.h"
#include "TLibEncoder/TEncSearch.h"
#undef private
#undef protected

/* external-linkage free functions of TLibCommon/TComTrQuant.cpp:417-972 */
void partialButterfly4(short* src, short* dst, int shift, int line);
void partialButterfly8(short* src, short* dst, int shift, int line);
void partialButterfly16(short* src, short* dst, int shift, int line);
void partialButterfly32(short* src, short* dst, int shift, int line);
void partialButterflyInverse4(short* src, short* dst, int shift, int line);
void partialButterflyInverse8(short* src, short* dst, int shift, int line);
void partialButterflyInverse16(short* src, short* dst, int shift, int line);
void partialButterflyInverse32(short* src, short* dst, int shift, int line);
void fastForwardDst(short* block, short* coeff, int shift);
void fastInverseDst(short* tmp, short* block, int shift);
void xTrMxN(short* block, short* coeff, int iWidth, int iHeight, UInt uiMode);
void xITrMxN(short* coeff, short* block, int iWidth, int iHeight, UInt uiMode);

static bool s_rom_ready = false;
static TComRdCost* s_rd = 0;

extern "C" {

/* mirrors TAppEncCfg::xSetGlobal (App/TAppEncoder/TAppEncCfg.cpp:922-949) for the globals the
 * hot path reads */
void ref_init(int internal_bit_depth)
{
  if (!s_rom_ready) {
    initROM();
    s_rom_ready = true;
    s_rd = new TComRdCost;
    s_rd->init();
  }
  g_uiMaxCUWidth = 64;
  g_uiMaxCUHeight = 64;
  g_uiMaxCUDepth = 4;
  g_uiAddCUDepth = 1;
  g_uiBitDepth = 8;
  g_uiBitIncrement = (UInt)(internal_bit_depth - 8);
  g_uiBASE_MAX = ((1 << g_uiBitDepth) - 1);
  g_uiIBDI_MAX = ((1 << (g_uiBitDepth + g_uiBitIncrement)) - 1);
}

/* ---- distortion: integer-ME SAD exactly as xTZSearchHelp sets it up (TEncSearch.cpp:312-336) */
unsigned ref_sad_me(const short* org, int so, const short* cur, int sc, int w, int h, int subshift)
{
  TComPattern pat;
  pat.initPattern((Pel*)org, NULL, NULL, w, h, so, 0, 0, 0, 0);
  DistParam dp;
  s_rd->setDistParam(&pat, (Pel*)cur, sc, dp);
  dp.iSubShift = subshift;
  dp.bApplyWeight = false;
  return dp.DistFunc(&dp);
}

/* fractional-ME distortion as xPatternRefinement sets it up (TEncSearch.cpp:721-747) */
unsigned ref_dist_frac(const short* org, int so, const short* cur, int sc, int w, int h, int hadamard)
{
  TComPattern pat;
  pat.initPattern((Pel*)org, NULL, NULL, w, h, so, 0, 0, 0, 0);
  DistParam dp;
  s_rd->setDistParam(&pat, (Pel*)cur, sc, 1, dp, hadamard != 0);
  dp.bApplyWeight = false;
  return dp.DistFunc(&dp);
}

/* TComRdCost::getDistPart, unweighted (TComRdCost.cpp:449-478) */
unsigned ref_get_dist_part(const short* cur, int sc, const short* org, int so, int w, int h, int dfunc)
{
  return s_rd->getDistPart((Pel*)cur, sc, (Pel*)org, so, (UInt)w, (UInt)h, false, (DFunc)dfunc);
}

unsigned ref_calc_had(const short* p0, int s0, const short* p1, int s1, int w, int h)
{
  return s_rd->calcHAD((Pel*)p0, s0, (Pel*)p1, s1, w, h);
}

unsigned ref_component_bits(int v) { return s_rd->xGetComponentBits(v); }

unsigned ref_mv_cost(double lambda, int x, int y, int scale, int predx, int predy, unsigned* lambda_cost)
{
  s_rd->setLambda(lambda);
  s_rd->getMotionCost(1, 0);
  TComMv p(predx, predy);
  s_rd->setPredictor(p);
  s_rd->setCostScale(scale);
  if (lambda_cost) *lambda_cost = s_rd->m_uiCost;
  return s_rd->getCost(x, y);
}

/* ---- interpolation (TComInterpolationFilter public API) */
void ref_filter_hor_luma(short* src, int ss, short* dst, int ds, int w, int h, int frac, int isLast)
{ TComInterpolationFilter f; f.filterHorLuma(src, ss, dst, ds, w, h, frac, isLast != 0); }
void ref_filter_ver_luma(short* src, int ss, short* dst, int ds, int w, int h, int frac, int isFirst, int isLast)
{ TComInterpolationFilter f; f.filterVerLuma(src, ss, dst, ds, w, h, frac, isFirst != 0, isLast != 0); }
void ref_filter_hor_chroma(short* src, int ss, short* dst, int ds, int w, int h, int frac, int isLast)
{ TComInterpolationFilter f; f.filterHorChroma(src, ss, dst, ds, w, h, frac, isLast != 0); }
void ref_filter_ver_chroma(short* src, int ss, short* dst, int ds, int w, int h, int frac, int isFirst, int isLast)
{ TComInterpolationFilter f; f.filterVerChroma(src, ss, dst, ds, w, h, frac, isFirst != 0, isLast != 0); }

/* ---- transforms */
void ref_partial_butterfly(int n, short* src, short* dst, int shift, int line)
{
  switch (n) {
    case 4: partialButterfly4(src, dst, shift, line); break;
    case 8: partialButterfly8(src, dst, shift, line); break;
    case 16: partialButterfly16(src, dst, shift, line); break;
    default: partialButterfly32(src, dst, shift, line); break;
  }
}
void ref_partial_butterfly_inverse(int n, short* src, short* dst, int shift, int line)
{
  switch (n) {
    case 4: partialButterflyInverse4(src, dst, shift, line); break;
    case 8: partialButterflyInverse8(src, dst, shift, line); break;
    case 16: partialButterflyInverse16(src, dst, shift, line); break;
    default: partialButterflyInverse32(src, dst, shift, line); break;
  }
}
void ref_fast_forward_dst(short* b, short* c, int shift) { fastForwardDst(b, c, shift); }
void ref_fast_inverse_dst(short* t, short* b, int shift) { fastInverseDst(t, b, shift); }
void ref_xTrMxN(short* block, short* coeff, int w, int h, int use_dst) { xTrMxN(block, coeff, w, h, use_dst ? 0 : REG_DCT); }
void ref_xITrMxN(short* coeff, short* block, int w, int h, int use_dst) { xITrMxN(coeff, block, w, h, use_dst ? 0 : REG_DCT); }

void ref_dct_matrix(int n, short* out)
{
  for (int k = 0; k < n; k++)
    for (int j = 0; j < n; j++)
      out[k * n + j] = n == 4 ? g_aiT4[k][j] : n == 8 ? g_aiT8[k][j] : n == 16 ? g_aiT16[k][j] : g_aiT32[k][j];
}

/* g_auiSigLastScan[scanIdx][log2-1]; scan_idx 0 diag, 1 hor, 2 ver (ours) */
void ref_scan(int scan_idx, int log2size, unsigned* out)
{
  int ref_idx = scan_idx == 0 ? SCAN_DIAG : (scan_idx == 1 ? SCAN_HOR : SCAN_VER);
  const UInt* s = g_auiSigLastScan[ref_idx][log2size - 1];
  memcpy(out, s, sizeof(UInt) << (2 * log2size));
}

/* ---- TComTrQuant with a minimal CU/slice context --------------------------------------- */
struct TqCtx {
  TComTrQuant tq;
  TComDataCU cu;
  TComSlice slice;
  TComSPS sps;
  TComPPS pps;
  std::vector<UChar> tskip[3];
  std::vector<Char> predmode;
  std::vector<UChar> lumadir, chromadir, depth;
  Bool bypass[256];
  bool inited;
  TqCtx() : inited(false) {}
};
static TqCtx* s_tq = 0;

static TqCtx* tq_get()
{
  if (!s_tq) {
    s_tq = new TqCtx;
    TqCtx* c = s_tq;
    /* TEncTop.cpp:322-339 init order: RDOQ off here so xQuant takes the plain branch */
    c->tq.init(64, 64, 32, 0, NULL, NULL, NULL, false, true, false, true);
    c->tq.setFlatScalingList();
    c->tq.setUseScalingList(false);
    const int nparts = 256;
    for (int i = 0; i < 3; i++) c->tskip[i].assign(nparts, 0);
    c->predmode.assign(nparts, MODE_INTER);
    c->lumadir.assign(nparts, DC_IDX);
    c->chromadir.assign(nparts, DC_IDX);
    c->depth.assign(nparts, 0);
    memset(c->bypass, 0, sizeof(c->bypass));
    c->cu.m_puhTransformSkip[0] = &c->tskip[0][0];
    c->cu.m_puhTransformSkip[1] = &c->tskip[1][0];
    c->cu.m_puhTransformSkip[2] = &c->tskip[2][0];
    c->cu.m_pePredMode = &c->predmode[0];
    c->cu.m_puhLumaIntraDir = &c->lumadir[0];
    c->cu.m_puhChromaIntraDir = &c->chromadir[0];
    c->cu.m_puhDepth = &c->depth[0];
    c->cu.m_CUTransquantBypass = c->bypass;
    c->cu.m_pcSlice = &c->slice;
    c->slice.setSPS(&c->sps);
    c->slice.setPPS(&c->pps);
    c->sps.setMaxTrSize(32);
    c->inited = true;
  }
  return s_tq;
}

static void tq_detach(TqCtx* c)
{
  /* the TComDataCU destructor is never run (static lifetime); nothing to free */
  (void)c;
}

/* TComTrQuant::xT / xIT (private) */
void ref_xT(int use_dst, short* resi, int stride, int* coeff, int w, int h)
{ tq_get()->tq.xT(use_dst ? 0 : REG_DCT, resi, (UInt)stride, coeff, w, h); }
void ref_xIT(int use_dst, int* coeff, short* resi, int stride, int w, int h)
{ tq_get()->tq.xIT(use_dst ? 0 : REG_DCT, coeff, resi, (UInt)stride, w, h); }
void ref_transform_skip(short* resi, int stride, int* coeff, int w, int h)
{ tq_get()->tq.xTransformSkip(resi, (UInt)stride, coeff, w, h); }
void ref_itransform_skip(int* coeff, short* resi, int stride, int w, int h)
{ tq_get()->tq.xITransformSkip(coeff, resi, (UInt)stride, w, h); }

/* setQPforQuant (TComTrQuant.cpp:192-222) */
void ref_set_qp(int qpy, int is_luma, int qp_bd_offset, int chroma_qp_offset, int* per, int* rem)
{
  TqCtx* c = tq_get();
  c->tq.setQPforQuant(qpy, is_luma ? TEXT_LUMA : TEXT_CHROMA_U, qp_bd_offset, chroma_qp_offset);
  *per = c->tq.m_cQP.m_iPer;
  *rem = c->tq.m_cQP.m_iRem;
}

/* xQuant non-RDOQ branch (TComTrQuant.cpp:1102-1270).
 * qpy/base_qpy are luma QPs before the bit-depth offset; is_intra_cu selects the scan via
 * getCoefScanIdx together with luma_dir (intra luma only). */
void ref_quant(int* coef, int* qcoef, int* arl, int w, int h, int qpy, int base_qpy, int qp_bd_offset,
               int is_luma, int is_intra_slice, int is_intra_cu, int luma_dir, int sign_hide, int use_arl,
               unsigned* abs_sum)
{
  TqCtx* c = tq_get();
  c->tq.m_bUseRDOQ = false;
  c->tq.m_bUseAdaptQpSelect = use_arl != 0;
  c->tq.m_useTansformSkipFast = false;
  c->slice.setSliceType(is_intra_slice ? I_SLICE : P_SLICE);
  c->slice.setSliceQpBase(base_qpy);
  c->sps.setQpBDOffsetY(qp_bd_offset);
  c->sps.setQpBDOffsetC(qp_bd_offset);
  c->pps.setSignHideFlag(sign_hide);
  c->predmode[0] = is_intra_cu ? MODE_INTRA : MODE_INTER;
  c->lumadir[0] = (UChar)luma_dir;
  c->chromadir[0] = (UChar)luma_dir;
  c->tq.setQPforQuant(qpy, is_luma ? TEXT_LUMA : TEXT_CHROMA_U, qp_bd_offset, 0);
  UInt acsum = *abs_sum;
  Int* parl = arl;
  c->tq.xQuant(&c->cu, coef, qcoef, parl, w, h, acsum, is_luma ? TEXT_LUMA : TEXT_CHROMA_U, 0);
  *abs_sum = acsum;
  tq_detach(c);
}

/* xRateDistOptQuant (TComTrQuant.cpp:1719-2305, private) through xQuant with RDOQ switched on.
 * est points at 254 ints laid out like estBitsSbacStruct (TComTrQuant.h:59-72) and is copied into the
 * quantiser's own table; the minimal CU carries what the function reads: prediction mode, intra
 * directions (-> getCoefScanIdx), transform index (-> cbf context), PPS sign hiding.  The flat scaling
 * lists are rebuilt on every call because their error scale depends on g_uiBitIncrement. */
void ref_rdoq(int* coef, int* qcoef, int* arl, int w, int qpy, int qp_bd_offset, int is_luma, int is_intra_cu,
              int intra_dir, int tr_idx, int sign_hide, int use_arl, double lambda, const int* est, unsigned* abs_sum)
{
  TqCtx* c = tq_get();
  static std::vector<UChar> tridx(256, 0);
  c->tq.setFlatScalingList();
  c->tq.m_bUseRDOQ = true;
  c->tq.m_bUseAdaptQpSelect = use_arl != 0;
  c->tq.m_useTansformSkipFast = false;
  c->slice.setSliceType(is_intra_cu ? I_SLICE : P_SLICE);
  c->sps.setQpBDOffsetY(qp_bd_offset);
  c->sps.setQpBDOffsetC(qp_bd_offset);
  c->pps.setSignHideFlag(sign_hide);
  c->predmode[0] = is_intra_cu ? MODE_INTRA : MODE_INTER;
  c->lumadir[0] = (UChar)intra_dir;
  c->chromadir[0] = (UChar)intra_dir;
  tridx[0] = (UChar)tr_idx;
  c->cu.m_puhTrIdx = &tridx[0];
  memcpy(c->tq.m_pcEstBitsSbac, est, sizeof(estBitsSbacStruct));
  c->tq.setLambda(lambda, lambda);
  c->tq.selectLambda(is_luma ? TEXT_LUMA : TEXT_CHROMA_U);
  c->tq.setQPforQuant(qpy, is_luma ? TEXT_LUMA : TEXT_CHROMA_U, qp_bd_offset, 0);
  UInt acsum = *abs_sum;
  Int* parl = arl;
  c->tq.xQuant(&c->cu, coef, qcoef, parl, w, w, acsum, is_luma ? TEXT_LUMA : TEXT_CHROMA_U, 0);
  *abs_sum = acsum;
  c->tq.m_bUseRDOQ = false;
}
int ref_est_bits_size(void) { return (int)sizeof(estBitsSbacStruct); }

/* xDeQuant flat branch (TComTrQuant.cpp:1272-1355) */
void ref_dequant(int* qcoef, int* coef, int w, int h, int qpy, int qp_bd_offset, int is_luma)
{
  TqCtx* c = tq_get();
  c->tq.setUseScalingList(false);
  c->tq.setQPforQuant(qpy, is_luma ? TEXT_LUMA : TEXT_CHROMA_U, qp_bd_offset, 0);
  c->tq.xDeQuant(qcoef, coef, w, h, 0);
}

/* ---- glue (TComYuv / TComPicYuv) */
void ref_extend_border(short* pic, int stride, int w, int h, int mx, int my)
{
  TComPicYuv p;
  p.xExtendPicCompBorder(pic, stride, w, h, mx, my);
}


/* ==================================================================================== motion search
 * The reference's own TEncSearch members, driven the way xMotionEstimation drives them
 * (TEncSearch.cpp:4120-4207): getMotionCost(1,0) -> setPredictor -> setCostScale(2) ->
 * xSetSearchRange -> xPatternSearch / xTZSearch -> setCostScale(1) -> xPatternSearchFracDIF.
 * The TComDataCU carries only what clipMv reads (CU origin, SPS picture size). */
static TEncSearch* s_search = 0;
static TEncCfg* s_cfg = 0;
static TComDataCU* s_cu = 0;
static TComSlice* s_slice = 0;
static TComSPS* s_sps = 0;

void ref_me_setup(int pic_w, int pic_h, int search_range, int fen, int hadme)
{
  if (!s_search) {
    s_cfg = new TEncCfg;
    s_search = new TEncSearch;
    s_search->initTempBuff();                 /* m_filteredBlock / m_filteredBlockTmp (TComPrediction.cpp:85-95) */
    s_cu = new TComDataCU;
    s_slice = new TComSlice;
    s_sps = new TComSPS;
    s_slice->setSPS(s_sps);
    s_cu->m_pcSlice = s_slice;
    UInt* tmp = &g_auiZscanToRaster[0];       /* TEncCu::create, TEncCu.cpp:99-105 */
    initZscanToRaster(g_uiMaxCUDepth + 1, 1, 0, tmp);
    initRasterToZscan(g_uiMaxCUWidth, g_uiMaxCUHeight, g_uiMaxCUDepth + 1);
    initRasterToPelXY(g_uiMaxCUWidth, g_uiMaxCUHeight, g_uiMaxCUDepth + 1);
  }
  s_cfg->setUseHADME(hadme != 0);
  s_cfg->setUseFastEnc(fen != 0);
  s_search->m_pcEncCfg = s_cfg;
  s_search->m_pcRdCost = s_rd;
  s_search->m_iSearchRange = search_range;
  s_search->m_bipredSearchRange = 4;
  s_search->m_iFastSearch = 1;
  s_sps->setPicWidthInLumaSamples(pic_w);
  s_sps->setPicHeightInLumaSamples(pic_h);
}

static void me_prepare(int cu_x, int cu_y, double lambda, int predx, int predy, int cost_scale)
{
  s_cu->m_uiCUPelX = cu_x;
  s_cu->m_uiCUPelY = cu_y;
  s_rd->setLambda(lambda);
  s_rd->getMotionCost(1, 0);
  TComMv pred(predx, predy);
  s_rd->setPredictor(pred);
  s_rd->setCostScale(cost_scale);
}

void ref_set_search_range(int cu_x, int cu_y, int predx, int predy, int srange, int* out4)
{
  s_cu->m_uiCUPelX = cu_x;
  s_cu->m_uiCUPelY = cu_y;
  TComMv pred(predx, predy), lt, rb;
  s_search->xSetSearchRange(s_cu, pred, srange, lt, rb);
  out4[0] = lt.getHor(); out4[1] = lt.getVer(); out4[2] = rb.getHor(); out4[3] = rb.getVer();
}

/* mode 0: xPatternSearch (full); 1: xTZSearch.  ref points at the PU's co-located pel. out = mvx, mvy, sad */
void ref_int_search(int mode, short* org, int so, short* ref, int rs, int w, int h, int cu_x, int cu_y,
                    int lx, int ty, int rx, int by, double lambda, int predx, int predy, int startx_q, int starty_q, int* out3)
{
  me_prepare(cu_x, cu_y, lambda, predx, predy, 2);
  TComPattern pat;
  pat.initPattern(org, NULL, NULL, w, h, so, 0, 0, 0, 0);
  TComMv lt(lx, ty), rb(rx, by), mv(startx_q, starty_q);
  UInt sad = 0;
  if (mode == 0) s_search->xPatternSearch(&pat, ref, rs, &lt, &rb, mv, sad);
  else s_search->xTZSearch(s_cu, &pat, ref, rs, &lt, &rb, mv, sad);
  out3[0] = mv.getHor(); out3[1] = mv.getVer(); out3[2] = (int)sad;
}

/* xPatternSearchFracDIF at integer MV (imvx, imvy); out = halfx, halfy, qtrx, qtry, cost */
void ref_frac_search(short* org, int so, short* ref, int rs, int w, int h, int imvx, int imvy, double lambda,
                     int predx, int predy, int* out5)
{
  me_prepare(0, 0, lambda, predx, predy, 1);         /* cost scale 1 for the half-pel pass (:4187) */
  TComPattern pat;
  pat.initPattern(org, NULL, NULL, w, h, so, 0, 0, 0, 0);
  TComMv mvi(imvx, imvy), half, qtr;
  UInt cost = 0;
  s_search->xPatternSearchFracDIF(s_cu, &pat, ref, rs, &mvi, half, qtr, cost, false);
  out5[0] = half.getHor(); out5[1] = half.getVer(); out5[2] = qtr.getHor(); out5[3] = qtr.getVer(); out5[4] = (int)cost;
}

/* xMotionEstimation's integer + fractional stages for every census PU of one CTU against num_refs references, looped
 * here so that the CPU baseline of bench.py times the reference's own compiled TEncSearch code and not Python call
 * overhead.  census: 593 x (x, y, w, h, cu_x, cu_y) relative to the CTU (the caller passes tvc_me_census / orc_census).
 * cur / refs[r]: pel (0,0) of padded luma planes of equal stride.  out_int: num_refs*593 x (mvx, mvy, sad, valid);
 * out_frac: x (halfx, halfy, qtrx, qtry, cost). */
void ref_me_frame_ctu(short* cur, short** refs, int num_refs, int stride, int pic_w, int pic_h, int ctu_x, int ctu_y,
                      const int* pred_qpel, double lambda, int srange, const short* census, int* out_int, int* out_frac)
{
  for (int r = 0; r < num_refs; r++) {
    const int predx = pred_qpel[2 * r], predy = pred_qpel[2 * r + 1];
    for (int k = 0; k < 593; k++) {
      const short* c = census + 6 * k;
      const int x = ctu_x + c[0], y = ctu_y + c[1], w = c[2], h = c[3];
      int* io = out_int + ((size_t)r * 593 + k) * 4;
      int* fo = out_frac + ((size_t)r * 593 + k) * 5;
      io[0] = io[1] = io[2] = io[3] = 0;
      fo[0] = fo[1] = fo[2] = fo[3] = fo[4] = 0;
      if (x + w > pic_w || y + h > pic_h) continue;
      const int cu_x = ctu_x + c[4], cu_y = ctu_y + c[5];
      int rng[4];
      ref_set_search_range(cu_x, cu_y, predx, predy, srange, rng);
      short* o = cur + (ptrdiff_t)y * stride + x;
      short* rf = refs[r] + (ptrdiff_t)y * stride + x;
      ref_int_search(1, o, stride, rf, stride, w, h, cu_x, cu_y, rng[0], rng[1], rng[2], rng[3], lambda, predx, predy, predx, predy, io);
      io[3] = 1;
      ref_frac_search(o, stride, rf, stride, w, h, io[0], io[1], lambda, predx, predy, fo);
    }
  }
}

/* ==================================================================================== motion compensation
 * TComPrediction::xPredInterLumaBlk / xPredInterChromaBlk (TComPrediction.cpp:554-645) on a real
 * TComPicYuv: the caller's padded planes are copied into it (margins included). */
static TComPicYuv* s_refpic = 0;
static TComYuv* s_dst = 0;
static int s_ref_w = 0, s_ref_h = 0;

void ref_mc_set_ref(const short* y, const short* u, const short* v, int w, int h)
{
  /* y/u/v: full padded buffers, luma margin 80, stride w+160 (TComPicYuv layout) */
  if (!s_refpic || s_ref_w != w || s_ref_h != h) {
    s_refpic = new TComPicYuv;
    s_refpic->create(w, h, g_uiMaxCUWidth, g_uiMaxCUHeight, g_uiMaxCUDepth);
    s_ref_w = w; s_ref_h = h;
    if (!s_dst) { s_dst = new TComYuv; s_dst->create(g_uiMaxCUWidth, g_uiMaxCUHeight); }
  }
  memcpy(s_refpic->getBufY(), y, sizeof(short) * (size_t)(w + 160) * (h + 160));
  memcpy(s_refpic->getBufU(), u, sizeof(short) * (size_t)(w / 2 + 80) * (h / 2 + 80));
  memcpy(s_refpic->getBufV(), v, sizeof(short) * (size_t)(w / 2 + 80) * (h / 2 + 80));
}

/* PU at luma (x, y); out_y w*h, out_u/out_v (w/2)*(h/2), dense */
void ref_mc_pu(int x, int y, int w, int h, int mvx, int mvy, int bi, short* out_y, short* out_u, short* out_v)
{
  int ctus_x = (s_ref_w + 63) / 64;
  s_cu->m_uiCUAddr = (y / 64) * ctus_x + (x / 64);
  s_cu->m_uiAbsIdxInLCU = g_auiRasterToZscan[((y % 64) / 4) * 16 + (x % 64) / 4];
  TComMv mv(mvx, mvy);
  TComYuv* dst = s_dst;
  s_search->xPredInterLumaBlk(s_cu, s_refpic, 0, &mv, w, h, dst, bi != 0);
  s_search->xPredInterChromaBlk(s_cu, s_refpic, 0, &mv, w, h, dst, bi != 0);
  for (int r = 0; r < h; r++) memcpy(out_y + r * w, s_dst->getLumaAddr() + r * s_dst->getStride(), sizeof(short) * w);
  for (int r = 0; r < h / 2; r++) {
    memcpy(out_u + r * (w / 2), s_dst->getCbAddr() + r * s_dst->getCStride(), sizeof(short) * (w / 2));
    memcpy(out_v + r * (w / 2), s_dst->getCrAddr() + r * s_dst->getCStride(), sizeof(short) * (w / 2));
  }
}

/* TComYuv::addAvg (TComYuv.cpp:520-581) on two 14-bit predictions */
void ref_add_avg(const short* a_y, const short* a_u, const short* a_v, const short* b_y, const short* b_u, const short* b_v,
                 int w, int h, short* o_y, short* o_u, short* o_v)
{
  static TComYuv *A = 0, *B = 0, *D = 0;
  if (!A) { A = new TComYuv; B = new TComYuv; D = new TComYuv; A->create(64, 64); B->create(64, 64); D->create(64, 64); }
  for (int r = 0; r < h; r++) {
    memcpy(A->getLumaAddr() + r * 64, a_y + r * w, 2 * w);
    memcpy(B->getLumaAddr() + r * 64, b_y + r * w, 2 * w);
  }
  for (int r = 0; r < h / 2; r++) {
    memcpy(A->getCbAddr() + r * 32, a_u + r * (w / 2), w); memcpy(A->getCrAddr() + r * 32, a_v + r * (w / 2), w);
    memcpy(B->getCbAddr() + r * 32, b_u + r * (w / 2), w); memcpy(B->getCrAddr() + r * 32, b_v + r * (w / 2), w);
  }
  D->addAvg(A, B, 0, w, h);
  for (int r = 0; r < h; r++) memcpy(o_y + r * w, D->getLumaAddr() + r * 64, 2 * w);
  for (int r = 0; r < h / 2; r++) { memcpy(o_u + r * (w / 2), D->getCbAddr() + r * 32, w); memcpy(o_v + r * (w / 2), D->getCrAddr() + r * 32, w); }
}

/* ==================================================================================== intra rough search
 * The reference's own predIntraLumaAng (TComPrediction.cpp:337-366, through TComPattern::getPredictorPtr :577-605) and
 * calcHAD (TComRdCost.cpp:404-447), exactly as estIntraPredQT calls them per mode (TEncSearch.cpp:2530-2537).  The
 * caller supplies what initAdiPattern leaves in m_piYuvExt: the unfiltered reference line and the smoothed one
 * (4N+1 samples each: left column bottom to top, corner, row above left to right); initAdiPattern itself needs a
 * live TComDataCU, its smoothing loop is pinned by the vectors dumped from the running encoder (tests/golden). */
void ref_intra_rough(const short* line, const short* line_filtered, const short* org, int so, int log2n, int above, int left,
                     unsigned* sad, short* preds)
{
  static TComPrediction* P = 0;
  static TComPattern pat;
  if (!P) { P = new TComPrediction; P->initTempBuff(); }
  const int n = 1 << log2n, sw = 2 * n + 1, wh = sw * sw;
  Int* adi = P->m_piYuvExt;
  for (int b = 0; b < 2; b++) {
    const short* ln = b ? line_filtered : line;
    Int* a = adi + b * wh;
    int l = 0;
    for (int i = 0; i < 2 * n; i++) a[sw * (2 * n - i)] = ln[l++];
    a[0] = ln[l++];
    for (int i = 0; i < 2 * n; i++) a[1 + i] = ln[l++];
  }
  std::vector<short> tmp((size_t)n * n);
  for (int mode = 0; mode < 35; mode++) {
    short* d = preds ? preds + (size_t)mode * n * n : tmp.data();
    P->predIntraLumaAng(&pat, (UInt)mode, d, (UInt)n, n, n, 0, above != 0, left != 0);
    sad[mode] = s_rd->calcHAD(const_cast<short*>(org), so, d, n, n, n);
  }
}

/* ==================================================================================== picture hashes
 * the reference's own calcMD5 / calcCRC / calcChecksum (TComPicYuvMD5.cpp:119-200) on a TComPicYuv filled from dense planes */
void ref_pic_hash(int method, const short* y, const short* u, const short* v, int w, int h, unsigned char* digest /* [3][16] */)
{
  TComPicYuv pic;
  pic.create(w, h, g_uiMaxCUWidth, g_uiMaxCUHeight, g_uiMaxCUDepth);
  for (int r = 0; r < h; r++) memcpy(pic.getLumaAddr() + (size_t)r * pic.getStride(), y + (size_t)r * w, sizeof(short) * w);
  for (int r = 0; r < h / 2; r++) {
    memcpy(pic.getCbAddr() + (size_t)r * pic.getCStride(), u + (size_t)r * (w / 2), sizeof(short) * (w / 2));
    memcpy(pic.getCrAddr() + (size_t)r * pic.getCStride(), v + (size_t)r * (w / 2), sizeof(short) * (w / 2));
  }
  unsigned char d[3][16];
  memset(d, 0, sizeof(d));
  if (method == 1) calcMD5(pic, d); else if (method == 2) calcCRC(pic, d); else calcChecksum(pic, d);
  memcpy(digest, d, sizeof(d));
  pic.destroy();
}

} /* extern "C" */
