/*
 * tlibcuda_hm.cpp -- HM 7.2 side of the C ABI (see tlibcuda_hm.h).  Host C++; every computation is a
 * call into libthevc_cuda.so (include/thevc_cuda.h).
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <map>
#include <chrono>
#include <algorithm>
#include <unordered_map>
#include <utility>
#include <string>

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
#include "TLibCommon/TComSampleAdaptiveOffset.h"
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

// results of one tvc_me_ctu call: the census-wide searches of a (CTU, reference) group for one predictor
struct GroupEntry {
  int predx, predy, sr, fen, had;
  unsigned lambda;
  std::vector<tvc_me_result> ires;
  std::vector<tvc_frac_result> fres;
};
struct Group {
  std::vector<GroupEntry> e;
  std::map<long long, int> seen;           // predictors that missed: a second miss of the same one builds an entry
};
struct LastHit {                           // the integer search just served by look-up: its fractional stage follows
  bool valid = false;
  int slot, x, y, w, h, mvx, mvy, predx, predy, had;
  unsigned lambda;
  tvc_frac_result fr;
};

// decoder picture batch: what is pending and the CU being announced
struct DecCu { int x, y, w; };
struct DecBatch {
  bool on = false;                         // TVC_HM=...,batch
  bool cu_open = false;                    // between tlibcuda_dec_begin_inter and the next begin / flush
  int cu_x = 0, cu_y = 0;
  const short* resi_base[3] = {nullptr, nullptr, nullptr};
  int resi_stride[3] = {0, 0, 0};
  std::vector<tvc_pu> pus;
  std::vector<tvc_tu> tus;
  std::vector<int32_t> levels;
  std::vector<DecCu> cus;
  bool half = false;                       // first list of a bi-predicted PU seen, waiting for the second
  tvc_pu half_pu;
  std::vector<short> stage[3];             // download staging (picture size, no margin)
  unsigned long long n_flush = 0, n_cus = 0, n_pus = 0, n_tus = 0;
  double seconds = 0.0;
};

// deblocking: the edge-unit records of the picture being filtered
struct Dbk {
  bool on = false, dump = false, active = false;
  int w = 0, h = 0;
  std::vector<tvc_dbk_unit> ver, hor;
  std::vector<short> before[3];
  int dumped = 0;
  unsigned long long n_pics = 0, n_units = 0;
  double seconds = 0.0;
};

// SAO apply: the per-CTU records of the component being filtered
struct Sao {
  bool on = false, dump = false, active = false;
  int comp = 0, w = 0, h = 0;
  std::vector<tvc_sao_unit> units;
  std::vector<short> before;
  int dumped = 0;
  unsigned long long n_planes = 0;
  double seconds = 0.0;
};

// intra rough search: counters and the dump state
struct Intra {
  bool on = false, dump = false;
  int min_width = 16;
  unsigned long long n_pus = 0, n_host = 0;
  double seconds = 0.0;
  // dump mode: the PU being walked by the reference's own loop
  int log2n = 0;
  std::vector<short> line, org;
  unsigned sads[35];
  int per_size[7] = {0, 0, 0, 0, 0, 0, 0}, seen[7] = {0, 0, 0, 0, 0, 0, 0};
  FILE* f = nullptr;
};

struct State {
  tvc_ctx* h = nullptr;
  Intra intra;
  bool on_lookup = true, verify = false;
  bool wp = false;                         // weighted prediction in the current picture: ME / MC hooks stand down
  Sao sao;
  Dbk dbk;
  DecBatch dec;
  std::map<long long, Group> groups;       // key = ctu * 64 + device slot of the reference; cleared per picture
  std::map<unsigned, int> census_index;    // (x, y, w, h) inside the CTU -> census index
  std::vector<tvc_me_center> center_guess; // [table ref][ctu]: first predictor of the previous picture's group (quarter pels)
  LastHit last;
  // a census group asked for ahead of the CU loop (tvc_me_ctu_async): the group of the NEXT CTU with the predictor this CTU's group
  // just started with; one in flight per reference index
  struct Spec { bool valid = false; int ctu = -1, slot = -1, predx = 0, predy = 0, sr = 0, fen = 0, had = 0; unsigned lambda = 0; };
  Spec spec[TVC_ME_CTU_TICKETS];
  bool on_spec = true;                     // off: TVC_HM=...,nospec
  unsigned long long n_spec_hits = 0, n_spec_asked = 0;
  LastHit last_bi;                         // the bi-prediction refinement just run: its fractional stage follows
  unsigned long long n_bi = 0, n_bi_frac = 0, n_bi_host = 0;
  unsigned long long n_tz_lookup = 0, n_frac_lookup = 0, n_groups = 0, sum_n_sads = 0;
  double batch_seconds = 0.0, prepass_seconds = 0.0;      // host wall time spent inside tvc_me_ctu / picture_start
  bool on_me = true, on_frac = true, on_tq = true, on_rdoq = true, on_mc = true, on_tables = true, verbose = false, disabled = false;
  // frame pre-pass consumed by the CU loop (TVC_HM=...,frame): one tvc_me_frame per picture with the predictor guesses; a group whose
  // first real predictor equals its guess adopts these results instead of running tvc_me_ctu
  bool on_frame = false, frame_valid = false;
  std::vector<tvc_me_result> frame_int;
  std::vector<tvc_frac_result> frame_frac;
  std::vector<tvc_me_center> frame_pred;
  unsigned frame_lambda = 0;
  int frame_sr = 0, frame_fen = 0, frame_had = 0, frame_refs = 0, frame_nctu = 0;
  int last_sr = -1, last_fen = 0, last_had = 0;          // what the CU loop asked for last (the next picture's pre-pass uses it)
  unsigned long long n_frame_groups = 0;
  bool on_hash = false;                    // picture hashes on the device (TVC_HM=...,hash)
  unsigned long long n_hash = 0;
  bool on_psnr = false;                    // PSNR sums on the device (TVC_HM=...,psnr)
  unsigned long long n_psnr = 0;
  bool on_cand = false;                    // merge / AMVP candidate evaluation (TVC_HM=...,cand)
  bool on_bipred = false;                  // bi-prediction refinement searches on the device (off: TVC_HM=...,nobipred)
  bool on_cand_grid = false;               // ... served by look-up from CTU-wide cost grids (TVC_HM=...,candgrid)
  struct Grid { uint32_t w[TVC_GRID_WORDS]; };
  std::unordered_map<unsigned long long, Grid> grids;      // (CTU, reference slot, clipped MV) -> prefix-sum grids; cleared per picture
  unsigned long long n_grid_hits = 0, n_grid_fills = 0;
  unsigned long long n_merge = 0, n_merge_cands = 0, n_template = 0, n_merge_host = 0;
  double cand_seconds = 0.0;
  int w = 0, ht = 0;
  std::vector<DevPic> slots;
  unsigned long long clock = 0;
  int cur_slot = -1;
  bool tables_live = false;                // SAD tables of the current picture exist in HBM (TVC_ME_FUSED=0 only)
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
  if (s.h && s.on_frame)
    fprintf(stderr, "TLibCuda frame pre-pass: %llu (CTU, reference) groups took their 593 results from the picture-level tvc_me_frame call\n", s.n_frame_groups);
  if (s.h && s.on_lookup && s.n_spec_asked)
    fprintf(stderr, "TLibCuda look-ahead: %llu (CTU, reference) groups asked for ahead of the CU loop with the left neighbour's predictor, %llu of them taken (tvc_me_ctu_async / _fetch)\n",
            s.n_spec_asked, s.n_spec_hits);
  if (s.h && s.on_lookup)
    fprintf(stderr, "TLibCuda look-up: %llu of %llu xTZSearch and %llu of %llu xPatternSearchFracDIF calls served from %llu census-wide (CTU, reference) batches (%.3f s in tvc_me_ctu, %.3f s in picture uploads + SAD-table pre-passes)\n",
            s.n_tz_lookup, s.n_tz, s.n_frac_lookup, s.n_frac, s.n_groups, s.batch_seconds, s.prepass_seconds);
  if (s.h && s.on_lookup && s.n_tz_lookup)
    fprintf(stderr, "TLibCuda TZ work: %.1f candidates per served xTZSearch on average (the reference's own count, n_sads)\n",
            (double)s.sum_n_sads / (double)s.n_tz_lookup);
  if (s.dbk.on && s.dbk.n_pics)
    fprintf(stderr, "TLibCuda deblocking: %llu pictures, %llu edge units filtered on the device, %.3f s (upload + 2 kernels + download)\n",
            s.dbk.n_pics, s.dbk.n_units, s.dbk.seconds);
  if (s.sao.on && s.sao.n_planes)
    fprintf(stderr, "TLibCuda SAO: %llu planes filtered on the device, %.3f s (upload + kernel + download)\n", s.sao.n_planes, s.sao.seconds);
  if (s.on_hash)
    fprintf(stderr, "TLibCuda picture hash: %llu pictures hashed on the device\n", s.n_hash);
  if (s.on_psnr)
    fprintf(stderr, "TLibCuda PSNR: the squared-difference sums of %llu pictures computed on the device\n", s.n_psnr);
  if (s.on_cand_grid)
    fprintf(stderr, "TLibCuda candidate look-up: %llu of %llu candidate costs served from %llu CTU-wide (CTU, reference, MV) cost grids; %llu bi-predicted merge sets left to the reference's code\n",
            s.n_grid_hits, s.n_grid_hits + s.n_grid_fills, s.n_grid_fills, s.n_merge_host);
  if (s.on_cand)
    fprintf(stderr, "TLibCuda candidate evaluation: %llu xMergeEstimation calls (%llu candidates) and %llu xGetTemplateCost calls on the device, %.3f s in tvc_pred_cost_batch\n",
            s.n_merge, s.n_merge_cands, s.n_template, s.cand_seconds);
  if (s.h && (s.n_bi || s.n_bi_host))
    fprintf(stderr, "TLibCuda bi-prediction refinement: %llu xPatternSearch + %llu xPatternSearchFracDIF calls on the device (tvc_me_bipred), %llu left to the reference's code\n",
            s.n_bi, s.n_bi_frac, s.n_bi_host);
  if (s.intra.on)
    fprintf(stderr, "TLibCuda intra rough search: %llu PUs x 35 modes on the device (width >= %d), %llu smaller PUs by the reference's code, %.3f s in tvc_intra_rough\n",
            s.intra.n_pus, s.intra.min_width, s.intra.n_host, s.intra.seconds);
  if (s.h && s.dec.on)
    fprintf(stderr, "TLibCuda picture batch: %llu inter CUs (%llu PUs, %llu TUs) reconstructed in %llu device batches, %.3f s in the batches\n",
            s.dec.n_cus, s.dec.n_pus, s.dec.n_tus, s.dec.n_flush, s.dec.seconds);
  if (s.h)
    fprintf(stderr, "TLibCuda: %llu xTZSearch, %llu xPatternSearchFracDIF, %llu xT, %llu xIT, %llu xDeQuant, %llu xRateDistOptQuant, %llu xPredInterUni calls served; %llu kernel launches\n",
            s.n_tz, s.n_frac, s.n_xt, s.n_xit, s.n_dq, s.n_rdoq, s.n_mc, (unsigned long long)tvc_launch_count(s.h));
}

// TVC_HM=stats: no CUDA at all -- only count, per (picture, CTU, reference), how many integer searches the
// reference runs and with how many distinct AMVP predictors (sizing data for the frame pre-pass look-up)
struct PredStats {
  bool on = false;
  std::map<std::pair<long long, long long>, std::map<long long, unsigned> > m;    // (poc, ctu*64+ref) -> pred -> calls
  std::map<std::pair<long long, long long>, long long> first;
};
PredStats& PS() { static PredStats p; return p; }

void stats_report()
{
  PredStats& p = PS();
  if (!p.on) return;
  unsigned long long calls = 0, groups = 0, hit_first = 0, hit_mode = 0, distinct = 0, hit_top2 = 0, hit_top4 = 0;
  unsigned long long hist[9] = {0};
  for (auto& g : p.m) {
    groups++;
    unsigned best = 0, tot = 0;
    std::vector<unsigned> cs;
    for (auto& e : g.second) { tot += e.second; if (e.second > best) best = e.second; cs.push_back(e.second); }
    std::sort(cs.begin(), cs.end());
    calls += tot; hit_mode += best; distinct += g.second.size();
    hit_first += g.second[p.first[g.first]];
    unsigned t2 = 0, t4 = 0;
    for (size_t i = 0; i < cs.size() && i < 4; i++) { if (i < 2) t2 += cs[cs.size() - 1 - i]; t4 += cs[cs.size() - 1 - i]; }
    hit_top2 += t2; hit_top4 += t4;
    hist[g.second.size() > 8 ? 8 : g.second.size()]++;
  }
  fprintf(stderr, "TLibCuda stats: %llu integer searches in %llu (picture, CTU, reference) groups; distinct predictors per group: mean %.2f; "
                  "served by the group's FIRST predictor %.1f%%, by its most frequent %.1f%%, by the top 2 %.1f%%, top 4 %.1f%%\n",
          calls, groups, groups ? (double)distinct / groups : 0.0, calls ? 100.0 * hit_first / calls : 0.0,
          calls ? 100.0 * hit_mode / calls : 0.0, calls ? 100.0 * hit_top2 / calls : 0.0, calls ? 100.0 * hit_top4 / calls : 0.0);
  if (const char* dump = getenv("TVC_STATS_DUMP")) {
    // one line per (picture, CTU, reference index) group: its FIRST predictor (quarter pels) and how many searches used it
    if (FILE* f = fopen(dump, "w")) {
      for (auto& g : p.m) {
        const long long pv = p.first[g.first];
        const int px = (int)(pv >> 20), py = (int)((pv & 0xfffff) ^ 0x80000) - 0x80000;
        unsigned tot = 0;
        for (auto& e : g.second) tot += e.second;
        fprintf(f, "%lld %lld %lld %d %d %u %u\n", g.first.first, g.first.second / 64, g.first.second % 64, px, py, g.second[pv], tot);
      }
      fclose(f);
    }
  }
  fprintf(stderr, "TLibCuda stats: groups by number of distinct predictors 1..8+:");
  for (int i = 1; i <= 8; i++) fprintf(stderr, " %llu", hist[i]);
  fprintf(stderr, "\n");
}

// TVC_HM is a comma list; every entry is matched as a WHOLE token ("frame" does not switch "me" on, "candgrid" implies "cand",
// "intraN" = intra rough search for PUs of width >= N).  Unset = me,frac,tq,rdoq,mc,tables (every per-call hook of the round-1 ABI);
// dbk, sao, intra, cand, candgrid, hash, frame, batch, bipred are opt-in.  Unknown tokens are fatal: a typo must not silently
// run the reference path.
void parse_env()
{
  State& s = S();
  const char* e = getenv("TVC_HM");
  if (!e) return;
  s.on_me = s.on_frac = s.on_tq = s.on_rdoq = s.on_mc = s.on_tables = false;
  std::string list(e);
  size_t pos = 0;
  while (pos <= list.size()) {
    size_t q = list.find(',', pos);
    if (q == std::string::npos) q = list.size();
    std::string t = list.substr(pos, q - pos);
    pos = q + 1;
    while (!t.empty() && (t.back() == ' ' || t.back() == '\t')) t.pop_back();
    while (!t.empty() && (t.front() == ' ' || t.front() == '\t')) t.erase(t.begin());
    if (t.empty() || t == "none") continue;
    if (t == "me") s.on_me = true;
    else if (t == "frac") s.on_frac = true;
    else if (t == "tq") s.on_tq = true;
    else if (t == "rdoq") s.on_rdoq = true;
    else if (t == "mc") s.on_mc = true;
    else if (t == "tables") s.on_tables = true;
    else if (t == "verbose") s.verbose = true;
    else if (t == "nolookup") s.on_lookup = false;
    else if (t == "nospec") s.on_spec = false;
    else if (t == "verify") s.verify = true;
    else if (t == "batch") s.dec.on = true;
    else if (t == "saodump") s.sao.dump = true;
    else if (t == "sao") s.sao.on = true;
    else if (t == "dbkdump") s.dbk.dump = true;
    else if (t == "dbk") s.dbk.on = true;
    else if (t == "frame") s.on_frame = true;
    else if (t == "hash") s.on_hash = true;
    else if (t == "psnr") s.on_psnr = true;
    else if (t == "cand") s.on_cand = true;
    else if (t == "candgrid") s.on_cand = s.on_cand_grid = true;
    else if (t == "bipred") s.on_bipred = true;
    else if (t == "nobipred") s.on_bipred = false;
    else if (t == "intradump") s.intra.dump = true;
    else if (t.compare(0, 5, "intra") == 0 && t.find_first_not_of("0123456789", 5) == std::string::npos) {
      s.intra.on = true;
      if (t.size() > 5) s.intra.min_width = atoi(t.c_str() + 5);
    } else if (t == "stats") PS().on = true;
    else { fprintf(stderr, "TLibCuda: unknown TVC_HM entry '%s'\n", t.c_str()); exit(EXIT_FAILURE); }
  }
  if (s.sao.dump) s.sao.on = false;
  if (s.dbk.dump) s.dbk.on = false;
  if (s.intra.dump) s.intra.on = false;
  if (PS().on) { s.on_me = s.on_frac = s.on_tq = s.on_rdoq = s.on_mc = s.on_tables = false; atexit(stats_report); }
}

void init_once()
{
  static bool done = false;
  if (done) return;
  done = true;
  parse_env();
  atexit(report);
}

void ensure_ctx(int w, int ht)
{
  State& s = S();
  if (s.disabled || (s.h && s.w >= w && s.ht >= ht)) return;
  init_once();
  if (!s.on_me && !s.on_frac && !s.on_tq && !s.on_rdoq && !s.on_mc && !s.dbk.on && !s.sao.on && !s.intra.on && !s.on_cand && !s.on_hash && !s.on_psnr) { s.disabled = true; return; }     // TVC_HM=none: the unmodified path
  if (s.h) {                  // the decoder learns the picture size after its first transforms: start over with the real size
    tvc_ctx_destroy(s.h);
    s.h = nullptr;
  }
  if (g_uiMaxCUWidth != 64 || g_uiMaxCUHeight != 64) {
    // the kernels' CTU geometry (census, SAD tables, cost grids) is the 64x64 CTU of every cfg the reference ships: any other
    // MaxCUWidth runs the reference's own code below every hook instead of dying in tvc_ctx_create
    fprintf(stderr, "TLibCuda: MaxCUWidth %u is not supported by the device path (64 only): hooks off, the reference's code runs\n", g_uiMaxCUWidth);
    s.disabled = true;
    return;
  }
  tvc_config c;
  c.width = w; c.height = ht;
  c.bit_depth = (int)(g_uiBitDepth + g_uiBitIncrement);
  c.max_cu = (int)g_uiMaxCUWidth;
  c.num_slots = 24;
  c.device = 0;
  int rc = tvc_ctx_create(&c, &s.h);
  if (rc != TVC_OK) {
    fprintf(stderr, "TLibCuda: tvc_ctx_create failed (%d): %s -- there is no CPU fallback\n", rc,
            rc == TVC_ERR_ARG ? "unsupported picture geometry" : "no usable CUDA device");
    exit(EXIT_FAILURE);
  }
  s.w = w; s.ht = ht;
  s.slots.assign(c.num_slots, DevPic());
  s.cur_slot = -1; s.num_table_refs = 0;
  // the last two slots are the decoder batch's reconstruction and residual planes: never handed out as picture slots
  s.slots[c.num_slots - 1].stamp = s.slots[c.num_slots - 2].stamp = ~0ull;
  s.slots[c.num_slots - 3].stamp = ~0ull;       // the picture being hashed (tlibcuda_pic_hash)
  s.slots[c.num_slots - 4].stamp = ~0ull;       // the bi-prediction search target (tlibcuda_full_search)
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
  // the encoder recycles its picture buffers (TEncTop::xGetNewPicBuffer, TEncTop.cpp:411-452: once the list holds
  // GOPSize + maxDecPicBuffering + 2 pictures the oldest unreferenced TComPic is reused): a slot that still maps this buffer
  // with ANOTHER POC holds a picture that no longer exists on the host.  It is taken over, so that no two slots ever alias one
  // host buffer and locate() cannot resolve a pointer to stale device content.
  for (size_t i = 0; i < s.slots.size(); i++) {
    DevPic& d = s.slots[i];
    if (d.yuv == yuv && d.is_org == is_org && d.stamp != ~0ull) {
      if (free_slot < 0) free_slot = (int)i;
      else d = DevPic();
    }
  }
  if (free_slot < 0)
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
  unsigned long long best = 0;
  bool found = false;
  for (size_t i = 0; i < s.slots.size(); i++) {
    DevPic& d = s.slots[i];
    if (!d.yuv || d.is_org || d.stamp == ~0ull) continue;
    const short* org = d.yuv->getLumaAddr();
    const int stride = d.yuv->getStride();
    const ptrdiff_t off = p - org;
    if (off < 0 || off >= (ptrdiff_t)stride * d.yuv->getHeight()) continue;
    const int yy = (int)(off / stride), xx = (int)(off % stride);
    if (xx >= d.yuv->getWidth()) continue;
    if (found && d.stamp < best) continue;        // slot_for keeps one slot per host buffer; should two ever match, the newest upload wins
    found = true; best = d.stamp;
    slot = (int)i; x = xx; y = yy;
  }
  return found;
}

}  // namespace

void tlibcuda_picture_start(TComPic* pic, TComSlice* slice)
{
  TComPicYuv* org = pic->getPicYuvOrg();
  const auto t_start = std::chrono::steady_clock::now();
  ensure_ctx(org->getWidth(), org->getHeight());
  State& s = S();
  if (!s.h || (!s.on_me && !s.on_frac && !s.on_cand)) return;
  s.wp = slice->getPPS()->getUseWP() || slice->getPPS()->getWPBiPred();
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
  s.groups.clear();
  s.grids.clear();
  s.last.valid = false;
  for (auto& sp : s.spec) sp.valid = false;
  if (s.census_index.empty()) {
    tvc_census_pu cen[TVC_ME_CENSUS];
    CK(tvc_me_census(cen));
    for (int k = TVC_ME_CENSUS - 1; k >= 0; k--)
      s.census_index[(unsigned)cen[k].x | ((unsigned)cen[k].y << 6) | ((unsigned)cen[k].w << 12) | ((unsigned)cen[k].h << 19)] = k;
  }
  const int nctu = ((org->getWidth() + 63) / 64) * ((org->getHeight() + 63) / 64);
  s.tables_live = false;
  if (s.on_tables && s.num_table_refs > 0 && g_uiBitIncrement == 0) {
    const bool hbm_tables = tvc_me_uses_tables(s.h) != 0;       // TVC_ME_FUSED=0: the round-1 form with SAD tables in HBM
    // table centres: the predictor each (reference index, CTU) group started with in the previous picture (steady
    // motion keeps it), clipped like a CTU-level clipMv; zero for the first inter picture
    if ((int)s.center_guess.size() != 8 * nctu) s.center_guess.assign((size_t)8 * nctu, tvc_me_center{0, 0});
    std::vector<tvc_me_center> cen((size_t)s.num_table_refs * nctu);
    const int ctus_x = (org->getWidth() + 63) / 64;
    for (int r = 0; r < s.num_table_refs; r++)
      for (int k = 0; k < nctu; k++) {
        tvc_me_center p = s.center_guess[(size_t)r * nctu + k];
        const int x0 = (k % ctus_x) * 64, y0 = (k / ctus_x) * 64;
        const int hmax = (org->getWidth() + 8 - x0 - 1) * 4, hmin = (-64 - 8 - x0 + 1) * 4;
        const int vmax = (org->getHeight() + 8 - y0 - 1) * 4, vmin = (-64 - 8 - y0 + 1) * 4;
        p.cx = (p.cx < hmin ? hmin : (p.cx > hmax ? hmax : p.cx)) >> 2;
        p.cy = (p.cy < vmin ? vmin : (p.cy > vmax ? vmax : p.cy)) >> 2;
        cen[(size_t)r * nctu + k] = p;
      }
    // a GOP ramps up to 4 references: one allocation instead of four growing ones (0.2-0.6 s each at 1080p); if the
    // device cannot hold four (34.8 GB at 1080p) the tables grow on demand as before
    static bool reserved = false;
    if (hbm_tables && !reserved) { reserved = true; if (tvc_me_reserve(s.h, s.num_table_refs > 4 ? s.num_table_refs : 4) != TVC_OK) (void)0; }
    s.frame_valid = false;
    if (s.on_frame && s.on_lookup && s.on_frac && s.last_sr > 0 && !s.wp && slice->getLambdaLuma() > 0.0) {
      // the whole census of the whole picture in one call: SAD tables around the guesses, TZ + fractional search of every PU with the
      // guess as predictor (TComRdCost::setLambda, TComRdCost.cpp:167-173, for the rate weight)
      double lam = slice->getLambdaLuma();
      const double root = sqrt(lam);
      const unsigned lc = (unsigned)floor(65536.0 * root);
      s.frame_pred.resize((size_t)s.num_table_refs * nctu);
      for (int r = 0; r < s.num_table_refs; r++)
        for (int k = 0; k < nctu; k++) s.frame_pred[(size_t)r * nctu + k] = s.center_guess[(size_t)r * nctu + k];
      const size_t n = (size_t)s.num_table_refs * nctu * TVC_ME_CENSUS;
      s.frame_int.resize(n); s.frame_frac.resize(n);
      tvc_me_frame_cfg fc = {s.last_sr, s.last_fen, s.last_had, 1, 1, lc};
      CK(tvc_me_frame(s.h, s.cur_slot, s.num_table_refs, s.table_refs, s.frame_pred.data(), &fc, s.frame_int.data(), s.frame_frac.data()));
      s.frame_valid = true; s.frame_lambda = lc; s.frame_sr = s.last_sr; s.frame_fen = s.last_fen; s.frame_had = s.last_had;
      s.frame_refs = s.num_table_refs; s.frame_nctu = nctu;
      s.tables_live = hbm_tables;
    } else if (hbm_tables) {
      CK(tvc_me_prepass(s.h, s.cur_slot, s.num_table_refs, s.table_refs, cen.data()));
      s.tables_live = true;
    }
  } else {
    s.num_table_refs = 0;
    s.frame_valid = false;
  }
  CK(tvc_sync(s.h));
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
  s.prepass_seconds += dt;
  if (s.verbose) fprintf(stderr, "TLibCuda: POC %d cur slot %d, %d reference(s) with SAD tables, %.1f ms (uploads + tables)\n", slice->getPOC(), s.cur_slot, s.num_table_refs, dt * 1e3);
}

bool tlibcuda_tz_search(TComDataCU* cu, TComPattern* key, short* refY, int refStride, TComMv* lt, TComMv* rb, TComMv& rcMv,
                        unsigned& ruiSAD, TComRdCost* rd, TEncCfg* cfg, int searchRange)
{
  State& s = S();
  if (PS().on) {
    const long long ctu = cu->getAddr();
    long long ref = 0;
    for (int i = 0; i < cu->getSlice()->getNumRefIdx(REF_PIC_LIST_0); i++) {     // which reference: by buffer range
      TComPicYuv* r = cu->getSlice()->getRefPic(REF_PIC_LIST_0, i)->getPicYuvRec();
      const ptrdiff_t off = refY - r->getLumaAddr();
      if (off >= 0 && off < (ptrdiff_t)r->getStride() * r->getHeight()) { ref = i; break; }
    }
    const std::pair<long long, long long> key(cu->getSlice()->getPOC(), ctu * 64 + ref);
    const long long pv = ((long long)rd->m_mvPredictor.getHor() << 20) ^ (rd->m_mvPredictor.getVer() & 0xfffff);
    if (!PS().m.count(key)) PS().first[key] = pv;
    PS().m[key][pv]++;
    return false;
  }
  if (!s.h || !s.on_me || s.cur_slot < 0) return false;
  if (searchRange < 1 || searchRange > TVC_ME_RANGE) return false;      // SearchRange beyond the kernels' +-64 window: the reference's own search
  // weighted prediction changes the distortion itself (setWpScalingDistParam, TEncSearch.cpp:4177): not ported, reference path
  if (cu->getSlice()->getPPS()->getUseWP() || cu->getSlice()->getPPS()->getWPBiPred()) return false;
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
  s.n_tz++;
  s.last.valid = false;
  if (s.on_lookup && s.on_frac) {
    // census-wide batch per (CTU, reference, predictor): the first search of a group pays for all of them
    const int ctus_x = (s.w + 63) / 64, ctu = (y >> 6) * ctus_x + (x >> 6);
    Group& g = s.groups[(long long)ctu * 64 + slot];
    const int had = cfg->getUseHADME() ? 1 : 0;
    GroupEntry* hit = nullptr;
    const bool first_of_group = g.e.empty();            // this search opens the group: whatever serves it also triggers the look-ahead
    for (auto& e : g.e)
      if (e.predx == j.predx && e.predy == j.predy && e.lambda == j.lambda_cost && e.sr == searchRange && e.fen == j.fen && e.had == had) { hit = &e; break; }
    s.last_sr = searchRange; s.last_fen = j.fen; s.last_had = had;
    if (!hit && g.e.empty() && s.frame_valid && j.ref_index >= 0 && j.ref_index < s.frame_refs && ctu < s.frame_nctu &&
        s.frame_lambda == j.lambda_cost && s.frame_sr == searchRange && s.frame_fen == j.fen && s.frame_had == had) {
      const size_t gi = (size_t)j.ref_index * s.frame_nctu + ctu;
      if (s.frame_pred[gi].cx == j.predx && s.frame_pred[gi].cy == j.predy) {       // the guess was right: the pre-pass already holds this group
        g.e.emplace_back();
        GroupEntry& e = g.e.back();
        e.predx = j.predx; e.predy = j.predy; e.lambda = j.lambda_cost; e.sr = searchRange; e.fen = j.fen; e.had = had;
        e.ires.assign(s.frame_int.begin() + gi * TVC_ME_CENSUS, s.frame_int.begin() + (gi + 1) * TVC_ME_CENSUS);
        e.fres.assign(s.frame_frac.begin() + gi * TVC_ME_CENSUS, s.frame_frac.begin() + (gi + 1) * TVC_ME_CENSUS);
        s.n_frame_groups++;
        hit = &e;
      }
    }
    if (!hit && g.e.empty() && s.on_spec && j.ref_index >= 0) {
      // asked for ahead of time while the previous CTU was coded?
      State::Spec& sp = s.spec[j.ref_index % TVC_ME_CTU_TICKETS];
      if (sp.valid) {
        sp.valid = false;
        if (sp.ctu == ctu && sp.slot == slot && sp.predx == j.predx && sp.predy == j.predy && sp.lambda == j.lambda_cost &&
            sp.sr == searchRange && sp.fen == j.fen && sp.had == had) {
          g.e.emplace_back();
          GroupEntry& e = g.e.back();
          e.predx = j.predx; e.predy = j.predy; e.lambda = j.lambda_cost; e.sr = searchRange; e.fen = j.fen; e.had = had;
          e.ires.resize(TVC_ME_CENSUS); e.fres.resize(TVC_ME_CENSUS);
          const auto t0 = std::chrono::steady_clock::now();
          CK(tvc_me_ctu_fetch(s.h, j.ref_index % TVC_ME_CTU_TICKETS, e.ires.data(), e.fres.data()));
          s.batch_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
          s.n_spec_hits++;
          if (j.ref_index >= 0 && (size_t)(j.ref_index + 1) * (s.center_guess.size() / 8) <= s.center_guess.size() && !s.center_guess.empty())
            s.center_guess[(size_t)j.ref_index * (s.center_guess.size() / 8) + ctu] = tvc_me_center{j.predx, j.predy};
          hit = &e;
        }
      }
    }
    if (!hit) {
      const long long pv = ((long long)j.predx << 32) ^ (unsigned)j.predy;
      if (g.e.empty() || (g.e.size() < 4 && g.seen[pv]++ >= 1)) {
        g.e.emplace_back();
        GroupEntry& e = g.e.back();
        e.predx = j.predx; e.predy = j.predy; e.lambda = j.lambda_cost; e.sr = searchRange; e.fen = j.fen; e.had = had;
        e.ires.resize(TVC_ME_CENSUS); e.fres.resize(TVC_ME_CENSUS);
        tvc_me_frame_cfg fc = {searchRange, j.fen, had, j.ref_index >= 0 ? 1 : 0, 1, j.lambda_cost};
        const auto t0 = std::chrono::steady_clock::now();
        CK(tvc_me_ctu(s.h, s.cur_slot, j.ref_index, slot, ctu, tvc_me_center{j.predx, j.predy}, &fc, e.ires.data(), e.fres.data()));
        s.batch_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        s.n_groups++;
        if (g.e.size() == 1 && j.ref_index >= 0 && (size_t)(j.ref_index + 1) * (s.center_guess.size() / 8) <= s.center_guess.size())
          s.center_guess[(size_t)j.ref_index * (s.center_guess.size() / 8) + ctu] = tvc_me_center{j.predx, j.predy};
        hit = &e;
      }
    }
    if (hit && first_of_group && g.e.size() == 1 && s.on_spec && j.ref_index >= 0 && ctu + 1 < ((s.w + 63) / 64) * ((s.ht + 63) / 64)) {
      // this group just got its first predictor: the next CTU's group for the same reference starts with the same one in 96.6 % of
      // HM's own 1080p groups -- ask for it now, on the side stream, while the host codes this CTU
      const size_t gi = (size_t)j.ref_index * s.frame_nctu + (ctu + 1);
      const bool frame_has_it = s.frame_valid && j.ref_index < s.frame_refs && ctu + 1 < s.frame_nctu && s.frame_lambda == j.lambda_cost &&
                                s.frame_sr == searchRange && s.frame_fen == j.fen && s.frame_had == had &&
                                s.frame_pred[gi].cx == j.predx && s.frame_pred[gi].cy == j.predy;
      if (!frame_has_it) {
        tvc_me_frame_cfg fc = {searchRange, j.fen, had, 1, 1, j.lambda_cost};
        CK(tvc_me_ctu_async(s.h, j.ref_index % TVC_ME_CTU_TICKETS, s.cur_slot, j.ref_index, slot, ctu + 1, tvc_me_center{j.predx, j.predy}, &fc));
        State::Spec& sp = s.spec[j.ref_index % TVC_ME_CTU_TICKETS];
        sp.valid = true; sp.ctu = ctu + 1; sp.slot = slot; sp.predx = j.predx; sp.predy = j.predy; sp.lambda = j.lambda_cost;
        sp.sr = searchRange; sp.fen = j.fen; sp.had = had;
        s.n_spec_asked++;
      }
    }
    if (hit) {
      auto it = s.census_index.find((unsigned)(x & 63) | ((unsigned)(y & 63) << 6) | ((unsigned)j.w << 12) | ((unsigned)j.h << 19));
      if (it != s.census_index.end() && hit->ires[it->second].n_sads > 0) {
        const tvc_me_result& r = hit->ires[it->second];
        if (s.verify) {
          tvc_me_result r2;
          CK(tvc_me_search_batch(s.h, s.cur_slot, (s.tables_live && j.ref_index >= 0) ? 1 : 0, 1, &j, &r2));
          if (r2.mvx != r.mvx || r2.mvy != r.mvy || r2.sad != r.sad) {
            fprintf(stderr, "TLibCuda verify: look-up (%d,%d,%u) != single search (%d,%d,%u) for PU %dx%d at (%d,%d)\n", r.mvx, r.mvy, r.sad,
                    r2.mvx, r2.mvy, r2.sad, j.w, j.h, x, y);
            exit(EXIT_FAILURE);
          }
        }
        rcMv.set(r.mvx, r.mvy);
        ruiSAD = r.sad;
        LastHit& l = s.last;
        l.valid = true; l.slot = slot; l.x = x; l.y = y; l.w = j.w; l.h = j.h; l.mvx = r.mvx; l.mvy = r.mvy;
        l.predx = j.predx; l.predy = j.predy; l.lambda = j.lambda_cost; l.had = had; l.fr = hit->fres[it->second];
        s.n_tz_lookup++;
        s.sum_n_sads += r.n_sads;
        return true;
      }
    }
  }
  tvc_me_result r;
  CK(tvc_me_search_batch(s.h, s.cur_slot, (s.tables_live && j.ref_index >= 0) ? 1 : 0, 1, &j, &r));
  rcMv.set(r.mvx, r.mvy);
  ruiSAD = r.sad;
  return true;
}

bool tlibcuda_frac_search(TComPattern* key, short* refY, int refStride, TComMv* mvInt, TComMv& half, TComMv& qter,
                          unsigned& ruiCost, TComRdCost* rd, TEncCfg* cfg, bool biPred)
{
  State& s = S();
  if (!s.h || !s.on_frac || s.cur_slot < 0 || s.wp) return false;
  int slot, x, y;
  if (!locate(refY, slot, x, y)) return false;
  if (biPred) {
    // the search target is 2 * org - pred(other list), not the picture: tlibcuda_full_search ran the integer AND the fractional
    // stage against it in one device call (tvc_me_bipred); without the `bipred` hook the reference's own code runs
    const LastHit& b = s.last_bi;
    if (b.valid && b.slot == slot && b.x == x && b.y == y && b.w == key->getROIYWidth() && b.h == key->getROIYHeight() &&
        b.mvx == mvInt->getHor() && b.mvy == mvInt->getVer() && b.predx == rd->m_mvPredictor.getHor() && b.predy == rd->m_mvPredictor.getVer() &&
        b.lambda == rd->m_uiCost && b.had == (cfg->getUseHADME() ? 1 : 0)) {
      half.set(b.fr.halfx, b.fr.halfy);
      qter.set(b.fr.qtrx, b.fr.qtry);
      ruiCost = b.fr.cost;
      rd->setCostScale(0);
      s.last_bi.valid = false;
      s.n_bi_frac++;
      return true;
    }
    s.n_bi_host++;
    return false;
  }
  (void)refStride;
  tvc_frac_job j;
  memset(&j, 0, sizeof(j));
  j.ref_slot = slot;
  j.x = x; j.y = y; j.w = key->getROIYWidth(); j.h = key->getROIYHeight();
  j.imvx = mvInt->getHor(); j.imvy = mvInt->getVer();
  j.predx = rd->m_mvPredictor.getHor(); j.predy = rd->m_mvPredictor.getVer();
  j.lambda_cost = rd->m_uiCost;
  j.hadamard = cfg->getUseHADME() ? 1 : 0;
  s.n_frac++;
  tvc_frac_result r;
  const LastHit& l = s.last;
  if (l.valid && l.slot == slot && l.x == x && l.y == y && l.w == j.w && l.h == j.h && l.mvx == j.imvx && l.mvy == j.imvy &&
      l.predx == j.predx && l.predy == j.predy && l.lambda == j.lambda_cost && l.had == j.hadamard) {
    r = l.fr;
    if (s.verify) {
      tvc_frac_result r2;
      CK(tvc_me_frac_batch(s.h, s.cur_slot, 1, &j, &r2));
      if (memcmp(&r, &r2, sizeof(r)) != 0) { fprintf(stderr, "TLibCuda verify: fractional look-up differs for PU %dx%d at (%d,%d)\n", j.w, j.h, x, y); exit(EXIT_FAILURE); }
    }
    s.n_frac_lookup++;
  } else
    CK(tvc_me_frac_batch(s.h, s.cur_slot, 1, &j, &r));
  s.last.valid = false;
  half.set(r.halfx, r.halfy);
  qter.set(r.qtrx, r.qtry);
  ruiCost = r.cost;
  rd->setCostScale(0);                      // side effect of the reference body (:4505)
  return true;
}

bool tlibcuda_full_search(TComPattern* key, short* refY, int refStride, TComMv* lt, TComMv* rb, TComMv& rcMv, unsigned& ruiSAD,
                          TComRdCost* rd, TEncCfg* cfg)
{
  State& s = S();
  s.last_bi.valid = false;
  if (!s.h || !s.on_bipred || !s.on_me || !s.on_frac || s.cur_slot < 0 || s.wp) return false;
  int slot, x, y;
  if (!locate(refY, slot, x, y)) return false;
  (void)refStride;
  tvc_me_job j;
  memset(&j, 0, sizeof(j));
  j.ref_slot = slot; j.ref_index = -1;
  j.x = x; j.y = y; j.w = key->getROIYWidth(); j.h = key->getROIYHeight();
  j.mode = TVC_ME_FULL;
  j.fen = cfg->getUseFastEnc() ? 1 : 0;
  j.search_range = TVC_ME_RANGE;
  j.lx = lt->getHor(); j.ty = lt->getVer(); j.rx = rb->getHor(); j.by = rb->getVer();
  j.predx = rd->m_mvPredictor.getHor(); j.predy = rd->m_mvPredictor.getVer();
  j.lambda_cost = rd->m_uiCost;
  if (j.rx - j.lx > 2 * TVC_ME_RANGE || j.by - j.ty > 2 * TVC_ME_RANGE) return false;
  const int had = cfg->getUseHADME() ? 1 : 0;
  tvc_me_result ri;
  LastHit& b = s.last_bi;
  CK(tvc_me_bipred(s.h, (int)s.slots.size() - 4, key->getROIY(), key->getPatternLStride(), &j, had, &ri, &b.fr));
  rcMv.set(ri.mvx, ri.mvy);
  ruiSAD = ri.sad;
  b.valid = true; b.slot = slot; b.x = x; b.y = y; b.w = j.w; b.h = j.h; b.mvx = ri.mvx; b.mvy = ri.mvy;
  b.predx = j.predx; b.predy = j.predy; b.lambda = j.lambda_cost; b.had = had;
  s.n_bi++;
  return true;
}

bool tlibcuda_pred_inter_uni(TComDataCU* cu, TComPic* refPic, unsigned partAddr, int mvx, int mvy, int w, int h,
                             TComYuv* dst, bool bi)
{
  TComPicYuv* rec = refPic->getPicYuvRec();
  ensure_ctx(rec->getWidth(), rec->getHeight());
  State& s = S();
  if (!s.h || !s.on_mc) return false;
  if (cu->getSlice()->getPPS()->getUseWP() || cu->getSlice()->getPPS()->getWPBiPred()) return false;   // weighted prediction: reference path
  const int slot = slot_for(rec, refPic->getPOC(), false, true);
  const unsigned z = cu->getZorderIdxInCU() + partAddr;
  const unsigned raster = g_auiZscanToRaster[z];
  const int ctus_x = (int)cu->getPic()->getFrameWidthInCU();
  const int x = (int)(cu->getAddr() % ctus_x) * (int)g_uiMaxCUWidth + (int)g_auiRasterToPelX[raster];
  const int y = (int)(cu->getAddr() / ctus_x) * (int)g_uiMaxCUHeight + (int)g_auiRasterToPelY[raster];
  if (s.dec.cu_open) {
    // picture batch: record the PU; the two lists of a bi-predicted PU arrive as two consecutive calls (xPredInterBi)
    DecBatch& d = s.dec;
    if (!bi) {
      d.pus.push_back(tvc_pu{x, y, w, h, slot, mvx, mvy, -1, 0, 0});
    } else if (d.half && d.half_pu.x == x && d.half_pu.y == y && d.half_pu.w == w && d.half_pu.h == h) {
      d.half_pu.ref_slot1 = slot; d.half_pu.mvx1 = mvx; d.half_pu.mvy1 = mvy;
      d.pus.push_back(d.half_pu);
      d.half = false;
    } else {
      d.half_pu = tvc_pu{x, y, w, h, slot, mvx, mvy, -1, 0, 0};
      d.half = true;
    }
    s.n_mc++;
    return true;
  }
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

// ---- merge / AMVP candidate evaluation (TEncSearch::xMergeEstimation :3096-3149, xGetTemplateCost :4057-4118)
static void pu_origin(TComDataCU* cu, unsigned partAddr, int& x, int& y)
{
  const unsigned raster = g_auiZscanToRaster[cu->getZorderIdxInCU() + partAddr];
  const int ctus_x = (int)cu->getPic()->getFrameWidthInCU();
  x = (int)(cu->getAddr() % ctus_x) * (int)g_uiMaxCUWidth + (int)g_auiRasterToPelX[raster];
  y = (int)(cu->getAddr() / ctus_x) * (int)g_uiMaxCUHeight + (int)g_auiRasterToPelY[raster];
}

// cost of the PU at (px, py, w, h) inside its CTU with the clipped motion (slot, mvx, mvy): look-up in the CTU-wide grids, filled by one
// tvc_ctu_cost_grids call on the first request of a (CTU, reference, MV)
static unsigned grid_cost(TComDataCU* cu, unsigned partAddr, int w, int h, int slot, int mvx, int mvy, bool hadamard)
{
  State& s = S();
  const unsigned raster = g_auiZscanToRaster[cu->getZorderIdxInCU() + partAddr];
  const int px = (int)g_auiRasterToPelX[raster], py = (int)g_auiRasterToPelY[raster];
  const unsigned long long key = ((unsigned long long)cu->getAddr() << 38) | ((unsigned long long)(slot & 63) << 32) |
                                 ((unsigned long long)(mvx & 0xffff) << 16) | (unsigned long long)(mvy & 0xffff);
  auto it = s.grids.find(key);
  if (it == s.grids.end()) {
    const int ctus_x = (int)cu->getPic()->getFrameWidthInCU();
    tvc_grid_job j = {slot, (int)(cu->getAddr() % ctus_x) * 64, (int)(cu->getAddr() / ctus_x) * 64, mvx, mvy};
    State::Grid g;
    const auto t0 = std::chrono::steady_clock::now();
    CK(tvc_ctu_cost_grids(s.h, s.cur_slot, 1, &j, g.w));
    s.cand_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    it = s.grids.emplace(key, g).first;
    s.n_grid_fills++;
  } else
    s.n_grid_hits++;
  const uint32_t* g = it->second.w;
  const bool h8 = hadamard && !(w & 7) && !(h & 7);
  const uint32_t* I = !hadamard ? g : (h8 ? g + 578 : g + 289);
  const int n = h8 ? 9 : 17, sh = h8 ? 3 : 2;
  const int x0 = px >> sh, y0 = py >> sh, x1 = (px + w) >> sh, y1 = (py + h) >> sh;
  const unsigned sum = I[y1 * n + x1] - I[y0 * n + x1] - I[y1 * n + x0] + I[y0 * n + x0];
  return sum >> g_uiBitIncrement;
}

bool tlibcuda_merge_costs(TComDataCU* cu, int puIdx, TComMvField* cands, const unsigned char* interDir, int numCand, bool hadamard,
                          unsigned* dist)
{
  State& s = S();
  if (!s.h || !s.on_cand || s.cur_slot < 0 || s.wp || numCand <= 0 || numCand > 5) return false;
  (void)interDir;
  UInt partAddr = 0;
  Int w = 0, h = 0;
  cu->getPartIndexAndSize((UInt)puIdx, partAddr, w, h);
  int x, y;
  pu_origin(cu, partAddr, x, y);
  TComSlice* sl = cu->getSlice();
  tvc_pu pus[5];
  for (int i = 0; i < numCand; i++) {
    tvc_pu& p = pus[i];
    p.x = x; p.y = y; p.w = w; p.h = h;
    p.ref_slot0 = p.ref_slot1 = -1; p.mvx0 = p.mvy0 = p.mvx1 = p.mvy1 = 0;
    const int r0 = cands[2 * i].getRefIdx(), r1 = cands[2 * i + 1].getRefIdx();
    if (r0 < 0 && r1 < 0) return false;
    TComMv m0 = cands[2 * i].getMv(), m1 = cands[2 * i + 1].getMv();
    // xCheckIdenticalMotion (TComPrediction.cpp:392-408): same picture, same vector in both lists -> uni-prediction from list 0
    bool use1 = r1 >= 0;
    if (r0 >= 0 && r1 >= 0 && sl->isInterB() && sl->getRefPic(REF_PIC_LIST_0, r0)->getPOC() == sl->getRefPic(REF_PIC_LIST_1, r1)->getPOC() && m0 == m1)
      use1 = false;
    if (r0 >= 0) {
      TComPic* rp = sl->getRefPic(REF_PIC_LIST_0, r0);
      cu->clipMv(m0);                        // xPredInterUni :483-490
      p.ref_slot0 = slot_for(rp->getPicYuvRec(), rp->getPOC(), false, true);
      p.mvx0 = m0.getHor(); p.mvy0 = m0.getVer();
    }
    if (use1) {
      TComPic* rp = sl->getRefPic(REF_PIC_LIST_1, r1);
      cu->clipMv(m1);
      p.ref_slot1 = slot_for(rp->getPicYuvRec(), rp->getPOC(), false, true);
      p.mvx1 = m1.getHor(); p.mvy1 = m1.getVer();
    }
  }
  s.n_merge++; s.n_merge_cands += (unsigned long long)numCand;
  bool all_uni = true;
  for (int i = 0; i < numCand; i++) all_uni &= (pus[i].ref_slot0 < 0) != (pus[i].ref_slot1 < 0);
  if (s.on_cand_grid && all_uni && g_uiMaxCUWidth == 64 && (((x & 7) == 0 && (y & 7) == 0) || (w & 7) || (h & 7))) {
    for (int i = 0; i < numCand; i++) {
      const tvc_pu& p = pus[i];
      dist[i] = p.ref_slot0 >= 0 ? grid_cost(cu, partAddr, w, h, p.ref_slot0, p.mvx0, p.mvy0, hadamard)
                                 : grid_cost(cu, partAddr, w, h, p.ref_slot1, p.mvx1, p.mvy1, hadamard);
    }
    return true;
  }
  // bi-predicted candidate sets (B slices) have no look-up form: the average of two predictions is not a sum of per-list block
  // costs.  In look-up mode they stay with the reference's own xGetInterPredictionError (a per-call device evaluation costs a
  // launch + synchronise per xMergeEstimation: measured at 1080p random access, 267 s of a 527 s encode); `cand` alone keeps
  // every candidate set on the device (parity configuration of the tests).
  if (s.on_cand_grid) { s.n_merge--; s.n_merge_cands -= (unsigned long long)numCand; s.n_merge_host++; return false; }
  const auto t0 = std::chrono::steady_clock::now();
  CK(tvc_pred_cost_batch(s.h, s.cur_slot, hadamard ? TVC_DIST_HADS : TVC_DIST_SAD, numCand, pus, dist));
  s.cand_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return true;
}

bool tlibcuda_template_sad(TComDataCU* cu, TComPic* refPic, unsigned partAddr, int mvx, int mvy, int w, int h, unsigned& sad)
{
  State& s = S();
  if (!s.h || !s.on_cand || s.cur_slot < 0 || s.wp) return false;
  tvc_pu p;
  pu_origin(cu, partAddr, p.x, p.y);
  p.w = w; p.h = h;
  p.ref_slot0 = slot_for(refPic->getPicYuvRec(), refPic->getPOC(), false, true);
  if (s.on_cand_grid && g_uiMaxCUWidth == 64) {
    sad = grid_cost(cu, partAddr, w, h, p.ref_slot0, mvx, mvy, false);
    s.n_template++;
    return true;
  }
  p.mvx0 = mvx; p.mvy0 = mvy;
  p.ref_slot1 = -1; p.mvx1 = p.mvy1 = 0;
  const auto t0 = std::chrono::steady_clock::now();
  CK(tvc_pred_cost_batch(s.h, s.cur_slot, TVC_DIST_SAD, 1, &p, &sad));
  s.cand_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  s.n_template++;
  return true;
}

bool tlibcuda_pic_hash(TComPicYuv& pic, int method, unsigned char digest[3][16])
{
  init_once();
  if (!S().on_hash) return false;
  ensure_ctx(pic.getWidth(), pic.getHeight());
  State& s = S();
  if (!s.h) return false;
  const int slot = (int)s.slots.size() - 3;
  s.slots[slot].yuv = nullptr;                // content changes per call: never matched by slot_for
  CK(tvc_pic_upload(s.h, slot, pic.getLumaAddr(), pic.getStride(), pic.getCbAddr(), pic.getCrAddr(), pic.getCStride(), 0));
  CK(tvc_pic_hash(s.h, slot, method, &digest[0][0]));
  s.n_hash++;
  return true;
}

bool tlibcuda_pic_ssd(TComPicYuv* org, TComPicYuv* rec, int padx, int pady, unsigned long long ssd[3])
{
  init_once();
  State& s = S();
  if (!s.on_psnr || padx || pady) return false;
  ensure_ctx(rec->getWidth(), rec->getHeight());
  if (!s.h) return false;
  // two of the reserved slots: the original and the reconstruction as they are on the host now (after the in-loop filters)
  const int a = (int)s.slots.size() - 3, b = (int)s.slots.size() - 4;
  s.slots[a].yuv = nullptr; s.slots[b].yuv = nullptr;
  CK(tvc_pic_upload(s.h, a, org->getLumaAddr(), org->getStride(), org->getCbAddr(), org->getCrAddr(), org->getCStride(), 0));
  CK(tvc_pic_upload(s.h, b, rec->getLumaAddr(), rec->getStride(), rec->getCbAddr(), rec->getCrAddr(), rec->getCStride(), 0));
  uint64_t v[3];
  CK(tvc_pic_ssd(s.h, a, b, v));
  ssd[0] = v[0]; ssd[1] = v[1]; ssd[2] = v[2];
  s.n_psnr++;
  return true;
}

unsigned tlibcuda_poc_offset()
{
  static int off = -1;
  if (off < 0) { const char* e = getenv("TVC_POC_OFFSET"); off = e ? atoi(e) : 0; if (off < 0) off = 0; }
  return (unsigned)off;
}

// ---- intra rough search (TEncSearch::estIntraPredQT, TEncSearch.cpp:2530-2543)
static void adi_to_line(const int* adi, int n, short* line)
{
  // initAdiPattern's own walk of m_piYuvExt (TComPattern.cpp:277-288): left column bottom to top, corner, row above
  const int sw = 2 * n + 1;
  int l = 0;
  for (int i = 0; i < 2 * n; i++) line[l++] = (short)adi[sw * (2 * n - i)];
  line[l++] = (short)adi[0];
  for (int i = 0; i < 2 * n; i++) line[l++] = (short)adi[1 + i];
}

bool tlibcuda_intra_rough(const int* adiBuf, unsigned width, const short* org, unsigned orgStride, bool above, bool left, unsigned* sad35)
{
  init_once();
  State& s = S();
  Intra& a = s.intra;
  if (!a.on && !a.dump) return false;
  const int n = (int)width;
  int log2n = 0;
  while ((1 << log2n) < n) log2n++;
  if (log2n < 2 || log2n > 6 || (1 << log2n) != n) return false;
  short line[4 * 64 + 1];
  if (a.dump) {
    a.log2n = 0;
    const int limit = getenv("TVC_INTRA_DUMP_PER_SIZE") ? atoi(getenv("TVC_INTRA_DUMP_PER_SIZE")) : 40;
    const int stride = getenv("TVC_INTRA_DUMP_STRIDE") ? atoi(getenv("TVC_INTRA_DUMP_STRIDE")) : 37;
    const int step = std::max(1, stride >> (2 * (log2n - 2)));      // a picture has 4x fewer PUs per size step: sample denser
    if (!getenv("TVC_INTRA_DUMP") || a.per_size[log2n] >= limit || (a.seen[log2n]++ % step) != 0) return false;
    adi_to_line(adiBuf, n, line);
    a.log2n = log2n;
    a.line.assign(line, line + 4 * n + 1);
    a.org.resize((size_t)n * n);
    for (int y = 0; y < n; y++) memcpy(&a.org[(size_t)y * n], org + (size_t)y * orgStride, sizeof(short) * n);
    return false;           // the reference's loop runs and reports every uiSad through tlibcuda_intra_note
  }
  if (n < a.min_width) { a.n_host++; return false; }
  if (!s.h && !s.disabled) ensure_ctx(64, 64);
  if (!s.h) return false;
  adi_to_line(adiBuf, n, line);
  const auto t0 = std::chrono::steady_clock::now();
  CK(tvc_intra_rough(s.h, log2n, line, org, (int)orgStride, above ? 1 : 0, left ? 1 : 0, sad35, nullptr));
  a.seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  a.n_pus++;
  return true;
}

void tlibcuda_intra_note(unsigned mode, unsigned sad)
{
  Intra& a = S().intra;
  if (!a.dump || !a.log2n || mode >= 35) return;
  a.sads[mode] = sad;
  if (mode != 34) return;
  // record: log2n, bit depth, line[4N+1], org[N*N], sad[35]; raw little-endian
  if (!a.f) {
    char name[1024];
    snprintf(name, sizeof(name), "%s/intra_rough.bin", getenv("TVC_INTRA_DUMP"));
    a.f = fopen(name, "wb");
    if (!a.f) { fprintf(stderr, "TLibCuda: cannot write %s\n", name); exit(EXIT_FAILURE); }
  }
  const int hdr[2] = {a.log2n, (int)(g_uiBitDepth + g_uiBitIncrement)};
  fwrite(hdr, sizeof(int), 2, a.f);
  fwrite(a.line.data(), sizeof(short), a.line.size(), a.f);
  fwrite(a.org.data(), sizeof(short), a.org.size(), a.f);
  fwrite(a.sads, sizeof(unsigned), 35, a.f);
  fflush(a.f);
  a.per_size[a.log2n]++;
  a.log2n = 0;
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

// ---------------------------------------------------------------------------------------------- decoder picture batch
bool tlibcuda_dec_begin_inter(TComDataCU* cu, TComYuv* resi)
{
  State& s = S();
  TComPicYuv* rec = cu->getPic()->getPicYuvRec();
  ensure_ctx(rec->getWidth(), rec->getHeight());
  DecBatch& d = s.dec;
  if (!s.h || !d.on || !s.on_mc || !s.on_tq) return false;
  if (cu->getSlice()->getPPS()->getUseWP() || cu->getSlice()->getPPS()->getWPBiPred()) return false;     // weighted prediction: host reconstruction
  // scaling lists: the batch dequantises with the flat per / rem tables only, so such a stream's CUs are never opened and the
  // reference's own xDeQuant( scalingListType ) reconstructs them (TComTrQuant.cpp:1272-1355)
  if (cu->getSlice()->getSPS()->getScalingListFlag()) { d.cu_open = false; return false; }
  d.cu_open = true;
  d.cu_x = (int)cu->getCUPelX(); d.cu_y = (int)cu->getCUPelY();
  d.resi_base[0] = resi->getLumaAddr(); d.resi_base[1] = resi->getCbAddr(); d.resi_base[2] = resi->getCrAddr();
  d.resi_stride[0] = (int)resi->getStride(); d.resi_stride[1] = d.resi_stride[2] = (int)resi->getCStride();
  d.cus.push_back(DecCu{d.cu_x, d.cu_y, (int)cu->getWidth(0)});
  return true;
}

bool tlibcuda_defer_itransform(bool bypass, int ttype, short* resi, unsigned stride, int* coeff, unsigned w, unsigned h, int per, int rem,
                               bool transformSkip, bool scalingList)
{
  State& s = S();
  DecBatch& d = s.dec;
  if (!d.cu_open || w != h) return false;
  if (scalingList && !bypass) {            // cannot happen: tlibcuda_dec_begin_inter does not open a CU of a scaling-list stream
    fprintf(stderr, "TLibCuda: scaling-list TU inside the decoder picture batch\n");
    exit(EXIT_FAILURE);
  }
  const int plane = ttype == TEXT_LUMA ? 0 : (ttype == TEXT_CHROMA_U ? 1 : 2);
  const ptrdiff_t off = resi - d.resi_base[plane];
  if (off < 0 || (int)stride != d.resi_stride[plane]) return false;      // not this CU's residual buffer: leave it to the host
  const int sh = plane ? 1 : 0;
  tvc_tu t;
  memset(&t, 0, sizeof(t));
  t.plane = plane;
  t.x = (d.cu_x >> sh) + (int)(off % stride);
  t.y = (d.cu_y >> sh) + (int)(off / stride);
  t.log2_size = w == 4 ? 2 : (w == 8 ? 3 : (w == 16 ? 4 : 5));
  t.flags = bypass ? TVC_TU_BYPASS : (transformSkip ? TVC_TU_SKIP : 0);
  t.qp_per = per; t.qp_rem = rem; t.base_per = per;
  t.coef_offset = (int)d.levels.size();
  d.levels.insert(d.levels.end(), coeff, coeff + (size_t)w * h);
  d.tus.push_back(t);
  return true;
}

void tlibcuda_dec_flush(TComPic* pic)
{
  State& s = S();
  DecBatch& d = s.dec;
  d.cu_open = false;
  if (!s.h || d.cus.empty()) return;
  const auto t0 = std::chrono::steady_clock::now();
  if (d.half) { d.pus.push_back(d.half_pu); d.half = false; }          // cannot happen without weighted prediction
  const int rec_slot = (int)s.slots.size() - 1, resi_slot = (int)s.slots.size() - 2;
  if (!d.pus.empty()) CK(tvc_mc_batch(s.h, rec_slot, (int)d.pus.size(), d.pus.data()));
  if (!d.tus.empty()) {
    // the ABI wants the TU list grouped by ascending size
    std::stable_sort(d.tus.begin(), d.tus.end(), [](const tvc_tu& a, const tvc_tu& b) { return a.log2_size < b.log2_size; });
    CK(tvc_inv_tq_batch(s.h, resi_slot, rec_slot, rec_slot, (int)d.tus.size(), d.tus.data(), d.levels.data(), d.levels.size()));
  }
  TComPicYuv* rec = pic->getPicYuvRec();
  const int W = rec->getWidth(), H = rec->getHeight();
  if (d.stage[0].size() != (size_t)W * H) {
    d.stage[0].assign((size_t)W * H, 0);
    d.stage[1].assign((size_t)(W / 2) * (H / 2), 0);
    d.stage[2].assign((size_t)(W / 2) * (H / 2), 0);
  }
  CK(tvc_pic_download(s.h, rec_slot, d.stage[0].data(), W, d.stage[1].data(), d.stage[2].data(), W / 2, 0));
  for (const DecCu& c : d.cus) {
    for (int pl = 0; pl < 3; pl++) {
      const int sh = pl ? 1 : 0, st = pl ? W / 2 : W;
      short* dst = (pl == 0 ? rec->getLumaAddr() : (pl == 1 ? rec->getCbAddr() : rec->getCrAddr()));
      const int ds = pl ? rec->getCStride() : rec->getStride();
      const int x = c.x >> sh, y = c.y >> sh, n = c.w >> sh;
      for (int r = 0; r < n; r++)
        memcpy(dst + (ptrdiff_t)(y + r) * ds + x, d.stage[pl].data() + (size_t)(y + r) * st + x, (size_t)n * sizeof(short));
    }
  }
  d.n_flush++; d.n_cus += d.cus.size(); d.n_pus += d.pus.size(); d.n_tus += d.tus.size();
  d.pus.clear(); d.tus.clear(); d.levels.clear(); d.cus.clear();
  d.seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---------------------------------------------------------------------------------------------- deblocking
void tlibcuda_dbk_begin(TComPic* pic)
{
  init_once();
  State& s = S();
  Dbk& d = s.dbk;
  d.active = false;
  if (!d.on && !d.dump) return;
  TComPicYuv* rec = pic->getPicYuvRec();
  if (d.on) { ensure_ctx(rec->getWidth(), rec->getHeight()); if (!s.h) return; }
  d.w = rec->getWidth(); d.h = rec->getHeight();
  d.ver.assign((size_t)((d.w + 7) >> 3) * ((d.h + 3) >> 2), tvc_dbk_unit{0, 0, 0, 0});
  d.hor.assign((size_t)((d.w + 3) >> 2) * ((d.h + 7) >> 3), tvc_dbk_unit{0, 0, 0, 0});
  d.active = true;
  if (d.dump) {
    for (int pl = 0; pl < 3; pl++) {
      const int w = pl ? d.w / 2 : d.w, h = pl ? d.h / 2 : d.h, st = pl ? rec->getCStride() : rec->getStride();
      const short* src = pl == 0 ? rec->getLumaAddr() : (pl == 1 ? rec->getCbAddr() : rec->getCrAddr());
      d.before[pl].resize((size_t)w * h);
      for (int r = 0; r < h; r++) memcpy(&d.before[pl][(size_t)r * w], src + (ptrdiff_t)r * st, (size_t)w * sizeof(short));
    }
  }
}

bool tlibcuda_dbk_unit(TComDataCU* cu, unsigned absZorderIdx, int dir, int edge, unsigned idx, unsigned bs, int qp, TComDataCU* cuP,
                       unsigned partP, TComDataCU* cuQ, unsigned partQ)
{
  Dbk& d = S().dbk;
  if (!d.active) return false;
  const int x0 = (int)cu->getCUPelX() + (int)g_auiRasterToPelX[g_auiZscanToRaster[absZorderIdx]];
  const int y0 = (int)cu->getCUPelY() + (int)g_auiRasterToPelY[g_auiZscanToRaster[absZorderIdx]];
  const int x = dir == 0 ? x0 + edge * 4 : x0 + (int)idx * 4, y = dir == 0 ? y0 + (int)idx * 4 : y0 + edge * 4;
  // xEdgeFilterLuma :651-658
  const bool pcm = cu->getSlice()->getSPS()->getUsePCM() && cu->getSlice()->getSPS()->getPCMFilterDisableFlag();
  const bool keepP = (pcm && cuP->getIPCMFlag(partP)) || cuP->isLosslessCoded(partP);
  const bool keepQ = (pcm && cuQ->getIPCMFlag(partQ)) || cuQ->isLosslessCoded(partQ);
  tvc_dbk_unit u = {(uint8_t)bs, (uint8_t)qp, (uint8_t)((keepP ? 1 : 0) | (keepQ ? 2 : 0)), 0};
  if (dir == 0) d.ver[(size_t)(y >> 2) * ((d.w + 7) >> 3) + (x >> 3)] = u;
  else d.hor[(size_t)(y >> 3) * ((d.w + 3) >> 2) + (x >> 2)] = u;
  d.n_units++;
  return d.on;            // device mode: the host does not filter; dump mode: the reference filters as usual
}

bool tlibcuda_dbk_skip_chroma() { const Dbk& d = S().dbk; return d.active && d.on; }

void tlibcuda_dbk_end(TComPic* pic, int betaOffsetDiv2, int tcOffsetDiv2)
{
  State& s = S();
  Dbk& d = s.dbk;
  if (!d.active) return;
  d.active = false;
  TComPicYuv* rec = pic->getPicYuvRec();
  if (d.on) {
    const auto t0 = std::chrono::steady_clock::now();
    const int slot = (int)s.slots.size() - 1;       // the reserved reconstruction slot
    CK(tvc_pic_upload(s.h, slot, rec->getLumaAddr(), rec->getStride(), rec->getCbAddr(), rec->getCrAddr(), rec->getCStride(), 0));
    CK(tvc_deblock_pic(s.h, slot, d.ver.data(), d.hor.data(), betaOffsetDiv2, tcOffsetDiv2));
    CK(tvc_pic_download(s.h, slot, rec->getLumaAddr(), rec->getStride(), rec->getCbAddr(), rec->getCrAddr(), rec->getCStride(), 0));
    d.n_pics++;
    d.seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return;
  }
  // dump mode: picture before / after the reference's own filtering + the records, raw little-endian
  const char* dir = getenv("TVC_DBK_DUMP");
  const int limit = getenv("TVC_DBK_DUMP_PICS") ? atoi(getenv("TVC_DBK_DUMP_PICS")) : 4;
  if (!dir || d.dumped >= limit) return;
  char name[512];
  snprintf(name, sizeof(name), "%s/dbk_%02d.bin", dir, d.dumped++);
  FILE* f = fopen(name, "wb");
  if (!f) return;
  const int hdr[6] = {d.w, d.h, (int)(g_uiBitDepth + g_uiBitIncrement), betaOffsetDiv2, tcOffsetDiv2, pic->getPOC()};
  fwrite(hdr, sizeof(int), 6, f);
  fwrite(d.ver.data(), sizeof(tvc_dbk_unit), d.ver.size(), f);
  fwrite(d.hor.data(), sizeof(tvc_dbk_unit), d.hor.size(), f);
  for (int pl = 0; pl < 3; pl++) fwrite(d.before[pl].data(), sizeof(short), d.before[pl].size(), f);
  for (int pl = 0; pl < 3; pl++) {
    const int w = pl ? d.w / 2 : d.w, h = pl ? d.h / 2 : d.h, st = pl ? rec->getCStride() : rec->getStride();
    const short* src = pl == 0 ? rec->getLumaAddr() : (pl == 1 ? rec->getCbAddr() : rec->getCrAddr());
    for (int r = 0; r < h; r++) fwrite(src + (ptrdiff_t)r * st, sizeof(short), (size_t)w, f);
  }
  fclose(f);
}

// ---------------------------------------------------------------------------------------------- SAO apply
static short* plane_of(TComPicYuv* rec, int comp) { return comp == 0 ? rec->getLumaAddr() : (comp == 1 ? rec->getCbAddr() : rec->getCrAddr()); }

void tlibcuda_sao_begin(TComPic* pic, int yCbCr, bool useNIF)
{
  init_once();
  State& s = S();
  Sao& a = s.sao;
  a.active = false;
  if ((!a.on && !a.dump) || useNIF) return;
  TComPicYuv* rec = pic->getPicYuvRec();
  if (a.on) { ensure_ctx(rec->getWidth(), rec->getHeight()); if (!s.h) return; }
  a.comp = yCbCr;
  a.w = rec->getWidth() >> (yCbCr ? 1 : 0); a.h = rec->getHeight() >> (yCbCr ? 1 : 0);
  tvc_sao_unit off;
  memset(&off, 0, sizeof(off));
  off.type = -1;
  a.units.assign((size_t)pic->getNumCUsInFrame(), off);
  a.active = true;
  if (a.dump) {
    const int st = yCbCr ? rec->getCStride() : rec->getStride();
    a.before.resize((size_t)a.w * a.h);
    for (int r = 0; r < a.h; r++) memcpy(&a.before[(size_t)r * a.w], plane_of(rec, yCbCr) + (ptrdiff_t)r * st, (size_t)a.w * sizeof(short));
  }
}

bool tlibcuda_sao_unit(int addr, int typeIdx, const int* offsetEo, const int* offsetBands)
{
  Sao& a = S().sao;
  if (!a.active) return false;
  tvc_sao_unit& u = a.units[addr];
  memset(&u, 0, sizeof(u));
  u.type = (int16_t)typeIdx;
  if (typeIdx == SAO_BO) for (int k = 0; k < 32; k++) u.bo[k] = (int16_t)offsetBands[k + 1];
  else for (int k = 0; k < 5; k++) u.eo[k] = (int16_t)offsetEo[k];
  return a.on;
}

void tlibcuda_sao_end(TComPic* pic, int yCbCr)
{
  State& s = S();
  Sao& a = s.sao;
  if (!a.active) return;
  a.active = false;
  TComPicYuv* rec = pic->getPicYuvRec();
  const int st = yCbCr ? rec->getCStride() : rec->getStride();
  if (a.on) {
    const auto t0 = std::chrono::steady_clock::now();
    const int src = (int)s.slots.size() - 1, dst = (int)s.slots.size() - 2;     // the two reserved slots
    if (yCbCr == 0)     // the chroma planes are still unfiltered when the luma pass runs: one upload serves all three components
      CK(tvc_pic_upload(s.h, src, rec->getLumaAddr(), rec->getStride(), rec->getCbAddr(), rec->getCrAddr(), rec->getCStride(), 0));
    CK(tvc_sao_plane(s.h, src, dst, yCbCr, a.units.data()));
    CK(tvc_pic_download(s.h, dst, yCbCr == 0 ? rec->getLumaAddr() : nullptr, rec->getStride(), yCbCr == 1 ? rec->getCbAddr() : nullptr,
                        yCbCr == 2 ? rec->getCrAddr() : nullptr, rec->getCStride(), 0));
    a.n_planes++;
    a.seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return;
  }
  const char* dir = getenv("TVC_SAO_DUMP");
  const int limit = getenv("TVC_SAO_DUMP_PLANES") ? atoi(getenv("TVC_SAO_DUMP_PLANES")) : 6;
  if (!dir || a.dumped >= limit) return;
  char name[512];
  snprintf(name, sizeof(name), "%s/sao_%02d.bin", dir, a.dumped++);
  FILE* f = fopen(name, "wb");
  if (!f) return;
  const int hdr[8] = {a.w, a.h, (int)(g_uiBitDepth + g_uiBitIncrement), yCbCr, (int)g_uiMaxCUWidth >> (yCbCr ? 1 : 0), (int)pic->getFrameWidthInCU(),
                      (int)a.units.size(), pic->getPOC()};
  fwrite(hdr, sizeof(int), 8, f);
  fwrite(a.units.data(), sizeof(tvc_sao_unit), a.units.size(), f);
  fwrite(a.before.data(), sizeof(short), a.before.size(), f);
  for (int r = 0; r < a.h; r++) fwrite(plane_of(rec, yCbCr) + (ptrdiff_t)r * st, sizeof(short), (size_t)a.w, f);
  fclose(f);
}
