// tvc_rdoq.cu -- rate-distortion optimised quantisation, batched over TUs.
//
// Replaces TComTrQuant::xRateDistOptQuant (TComTrQuant.cpp:1719-2305) and its helpers (xGetCodedLevel :2446,
// xGetICRateCost :2508, xGetICRate :2577, xGetRateLast :2652, getSigCtxInc :2349, calcPatternSigCtx :2315,
// getSigCoeffGroupCtxInc :2707, setErrScaleCoeff :2794 for flat lists).
//
// The algorithm is a backward walk over the coding scan whose flag contexts (c1, c2, Rice parameter, context
// set) depend on the levels chosen so far, and whose Lagrangian sums are IEEE doubles that must be added in
// scan order to reproduce the reference's decisions bit for bit.  The parallelism is therefore ACROSS TUs
// (a frame holds thousands) and, inside a TU, in everything that does not carry that state:
//
//   * one warp per TU, 4 warps per CTA; the TU's bit-estimate table (1016 B) sits in shared memory;
//   * per coefficient group (16 scan positions) lanes 0..15 quantise their coefficient in parallel
//     (Int64 product, level candidates, zero cost, significance context and its two lambda-weighted costs),
//     then the 16-step dependent chain runs warp-uniformly out of shared memory (no divergence, broadcast reads);
//     all-zero groups in front of the last significant coefficient only add their zero costs;
//   * the last-position search stages each group the same way; the output pass (sign, clean-up, uiAbsSum)
//     and sign-data hiding (independent per 16-coefficient subset) run lane-parallel.
//
// Every double operation is written with __dmul_rn/__dadd_rn/__dsub_rn/__ddiv_rn so that no FMA contraction can
// change a cost (the reference is x86-64 SSE2 code without FMA).
#include "tvc_internal.cuh"
#include <thread>
#include <sched.h>
#include <algorithm>

#include <cfloat>
#include <climits>
#include <cmath>

namespace tvc {

static __constant__ int c_rq_quantScales[6] = {26214, 23302, 20560, 18396, 16384, 14564};   // g_quantScales
static __constant__ int c_rq_invQuantScales[6] = {40, 45, 51, 57, 64, 72};                  // g_invQuantScales
static __constant__ uint8_t c_rq_groupIdx[32] = {0, 1, 2, 3, 4, 4, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7,
                                                 8, 8, 8, 8, 8, 8, 8, 8, 9, 9, 9, 9, 9, 9, 9, 9};   // g_uiGroupIdx
static __constant__ int c_rq_riceRange[5] = {7, 14, 26, 46, 78};                            // g_auiGoRiceRange
static __constant__ int c_rq_ricePrefix[5] = {8, 7, 6, 5, 4};                               // g_auiGoRicePrefixLen
static __constant__ uint8_t c_rq_map4[16] = {0, 1, 4, 5, 2, 3, 4, 5, 6, 6, 8, 8, 7, 7, 8, 8};   // ctxIndMap :2362

constexpr int kRdoqWarps = 4;

// int offsets of the members of tvc_est_bits
constexpr int EB_SIG_CG = 0, EB_SIG = 4, EB_LAST_X = 88, EB_LAST_Y = 120, EB_GT1 = 152, EB_GT2 = 200, EB_CBF = 212,
              EB_ROOT = 242, EB_INTS = 254;
static_assert(sizeof(tvc_est_bits) == EB_INTS * 4, "tvc_est_bits must mirror estBitsSbacStruct");

// per-coefficient working arrays of one launch, indexed [coef_offset + scan position]
struct RdoqScratch {
  double* coded;      // pdCostCoeff
  double* csig;       // pdCostSig
  double* cost0;      // pdCostCoeff0
  int32_t* level;     // chosen level (unsigned during the walk, signed after the output pass)
  int32_t* ctxw;      // flag-context state the level was chosen in (packed): rateIncUp / rateIncDown are derived from it where
  int32_t* orig;      // sign-data hiding needs them, together with the level chosen THEN (zeroed groups keep their increments)
  int32_t* sigd;      // sigRateDelta
  int32_t* du;        // deltaU
};
constexpr size_t kRdoqScratchBytesPerCoef = 3 * 8 + 5 * 4;

struct RdoqWarp {
  int32_t est[EB_INTS + 2];
  double cost0[16], sig0[16], sig1[16], coded[16], csig[16], lastc[16];
  double cg_sig[64];
  int32_t lvl_dbl[16], max_lvl[16], sigd_in[16], sigd[16], level[16], ctxw[16], orig[16], du[16];
};

struct LvlState { int ctx_set, c1, c2, rice, c1_idx, c2_idx; };

__device__ __forceinline__ int rq_base_level(const LvlState& s) { return s.c1_idx < 8 ? (2 + (s.c2_idx < 1)) : 1; }

// xGetICRate
__device__ __forceinline__ int rq_level_rate_int(const int32_t* est, int lvl, int one_ctx, int abs_ctx, const LvlState& s)
{
  const int base = rq_base_level(s);
  int rate = 0;
  if (lvl >= base) {
    unsigned sym = (unsigned)(lvl - base);
    const unsigned max_vlc = (unsigned)c_rq_riceRange[s.rice];
    if (sym > max_vlc) {
      const unsigned rest = sym - max_vlc;
      int egs = 1;
      for (unsigned m = 2; rest >= m; m <<= 1) egs += 2;
      rate += egs << 15;
      sym = min(sym, max_vlc + 1);
    }
    const unsigned pref = (sym >> s.rice) + 1;
    rate += (int)((min(pref, (unsigned)c_rq_ricePrefix[s.rice]) + (unsigned)s.rice) & 0xffffu) << 15;
    if (s.c1_idx < 8) {
      rate += est[EB_GT1 + 2 * one_ctx + 1];
      if (s.c2_idx < 1) rate += est[EB_GT2 + 2 * abs_ctx + 1];
    }
  } else if (lvl == 1) rate += est[EB_GT1 + 2 * one_ctx];
  else if (lvl == 2) rate += est[EB_GT1 + 2 * one_ctx + 1] + est[EB_GT2 + 2 * abs_ctx];
  return rate;
}

// xGetICRateCost: the rate is a sum of integers below 2^53, exact in double whatever the order
__device__ __forceinline__ double rq_level_rate_cost(const int32_t* est, double lambda, unsigned lvl, int one_ctx, int abs_ctx,
                                                     const LvlState& s)
{
  long long rate = 32768;
  const unsigned base = (unsigned)rq_base_level(s);
  if (lvl >= base) {
    unsigned sym = lvl - base, len;
    if (sym < (3u << s.rice)) {
      len = sym >> s.rice;
      rate += (int)((len + 1 + (unsigned)s.rice) << 15);
    } else {
      len = (unsigned)s.rice;
      sym -= 3u << s.rice;
      while (sym >= (1u << len)) sym -= 1u << (len++);
      rate += (int)((3 + len + 1 - (unsigned)s.rice + len) << 15);
    }
    if (s.c1_idx < 8) {
      rate += est[EB_GT1 + 2 * one_ctx + 1];
      if (s.c2_idx < 1) rate += est[EB_GT2 + 2 * abs_ctx + 1];
    }
  } else if (lvl == 1) rate += est[EB_GT1 + 2 * one_ctx];
  else rate += (long long)est[EB_GT1 + 2 * one_ctx + 1] + est[EB_GT2 + 2 * abs_ctx];
  return __dmul_rn(lambda, (double)rate);
}

// getSigCtxInc (REMOVAL_8x2_2x8_CG)
__device__ __forceinline__ int rq_sig_ctx(int pattern, int scan_idx, int px, int py, int log2, int is_luma)
{
  if (px + py == 0) return 0;
  if (log2 == 2) return c_rq_map4[4 * py + px];
  const int offset = log2 == 3 ? (scan_idx == 0 ? 9 : 15) : (is_luma ? 21 : 12);
  const int sx = px & 3, sy = py & 3;
  int cnt;
  if (pattern == 0) cnt = sx + sy <= 2 ? (sx + sy == 0 ? 2 : 1) : 0;
  else if (pattern == 1) cnt = sy <= 1 ? (sy == 0 ? 2 : 1) : 0;
  else if (pattern == 2) cnt = sx <= 1 ? (sx == 0 ? 2 : 1) : 0;
  else cnt = 2;
  return ((is_luma && ((px >> 2) + (py >> 2)) > 0) ? 3 : 0) + offset + cnt;
}

// xGetRateLast
__device__ __forceinline__ double rq_last_cost(const int32_t* est, double lambda, int px, int py)
{
  const unsigned cx = c_rq_groupIdx[px], cy = c_rq_groupIdx[py];
  long long r = (long long)est[EB_LAST_X + cx] + est[EB_LAST_Y + cy];     // Int sum in the reference: no overflow for bit estimates
  if (cx > 3) r += 32768ll * ((cx - 2) >> 1);
  if (cy > 3) r += 32768ll * ((cy - 2) >> 1);
  return __dmul_rn(lambda, (double)r);
}

// rateIncUp / rateIncDown of a position (:1935-1944) from the packed context state and the level chosen there
__device__ __forceinline__ void rq_rate_increments(const int32_t* est, int ctxw, int orig, int& rup, int& rdn)
{
  rup = 0; rdn = 0;
  if (ctxw >= 0) return;                        // bit 31 clear: in front of the last significant coefficient, never initialised (memset 0)
  const int one_ctx = ctxw & 31, abs_ctx = (ctxw >> 5) & 7;
  LvlState s = {0, 0, 0, (ctxw >> 8) & 7, (ctxw >> 11) & 31, (ctxw >> 16) & 31};
  if (orig > 0) {
    const int now = rq_level_rate_int(est, orig, one_ctx, abs_ctx, s);
    rup = rq_level_rate_int(est, orig + 1, one_ctx, abs_ctx, s) - now;
    rdn = rq_level_rate_int(est, orig - 1, one_ctx, abs_ctx, s) - now;
  } else rup = est[EB_GT1 + 2 * one_ctx];
}

__global__ void __launch_bounds__(kRdoqWarps * 32)
k_rdoq(int n, const tvc_rdoq_tu* __restrict__ tus, const tvc_est_bits* __restrict__ est_tab, int sign_hide, int use_arl, int bd,
       ScanTables scans, const int32_t* __restrict__ coef, int32_t* __restrict__ levels, int32_t* __restrict__ arl,
       uint32_t* __restrict__ abs_sum, RdoqScratch G)
{
  __shared__ RdoqWarp smem[kRdoqWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // TU lists are grouped by ascending size: walk them from the back so that the long 32x32 chains (1024 dependent
  // steps) start first and the short 4x4 ones fill the machine behind them instead of leaving a tail
  const int t = n - 1 - (blockIdx.x * kRdoqWarps + warp);
  if (t < 0) return;
  RdoqWarp& S = smem[warp];
  const tvc_rdoq_tu tu = tus[t];
  const int log2 = tu.log2_size, w = 1 << log2, ncoef = w * w, ncg = ncoef >> 4, cgw = w >> 2;
  const int is_luma = tu.is_luma, scan_idx = tu.scan_idx;
  const double lambda = tu.lambda;
  const uint16_t* __restrict__ scan = scans.s[scan_idx][log2 - 2];
  const size_t off = (size_t)tu.coef_offset;
  const int32_t* __restrict__ src = coef + off;

  {
    const int32_t* e = reinterpret_cast<const int32_t*>(est_tab + tu.est_index);
    for (int i = lane; i < EB_INTS; i += 32) S.est[i] = e[i];
  }
  const int tshift = 15 - bd - log2;
  const int qbits = 14 + tu.qp_per + tshift;
  const int qscale = c_rq_quantScales[tu.qp_rem];
  const int qbitsC = qbits - 7, addC = 1 << (qbitsC - 1);
  // setErrScaleCoeff: 2^15 * 2^(-2*tshift) / q / q / 2^(2*bi); the power of two is exact
  const double err_scale = __ddiv_rn(__ddiv_rn(__ddiv_rn(ldexp(32768.0, -2 * tshift), (double)qscale), (double)qscale),
                                     (double)(1 << (2 * (bd - 8))));
  __syncwarp();
  const int32_t* est = S.est;

  LvlState st = {0, 1, 0, 0, 0, 0};
  double base_cost = 0.0, uncoded_cost = 0.0;
  int last_pos = -1, last_cg = -1;
  unsigned long long cg_flag = 0ull;

  for (int cg = ncg - 1; cg >= 0; cg--) {
    const unsigned first = scan[cg << 4];
    const int cgy = (int)(first >> log2) >> 2, cgx = (int)(first & (unsigned)(w - 1)) >> 2;
    const int cgpos = cgy * cgw + cgx;
    int right = 0, lower = 0;
    if (cgx < cgw - 1) right = (int)((cg_flag >> (cgy * cgw + cgx + 1)) & 1ull);
    if (cgy < cgw - 1) lower = (int)((cg_flag >> ((cgy + 1) * cgw + cgx)) & 1ull);
    const int pattern = log2 == 2 ? -1 : right + (lower << 1);

    // ---- stage A: lane k quantises scan position cg*16 + k
    unsigned my_max = 0;
    if (lane < 16) {
      const int sp = (cg << 4) + lane;
      const unsigned bp = scan[sp];
      const int c = src[bp];
      const long long scaled = (long long)abs(c) * qscale;
      const long long cap = 2147483647ll - (1ll << (qbits - 1));
      const int lvl_dbl = (int)min(scaled, cap);
      if (use_arl && arl) arl[off + bp] = (lvl_dbl + addC) >> qbitsC;
      my_max = (unsigned)((lvl_dbl + (1 << (qbits - 1))) >> qbits);
      const double e0 = (double)lvl_dbl;
      S.lvl_dbl[lane] = lvl_dbl;
      S.max_lvl[lane] = (int)my_max;
      S.cost0[lane] = __dmul_rn(__dmul_rn(e0, e0), err_scale);
      const int ctx = rq_sig_ctx(pattern, scan_idx, (int)(bp & (unsigned)(w - 1)), (int)(bp >> log2), log2, is_luma);
      const int b0 = est[EB_SIG + 2 * ctx], b1 = est[EB_SIG + 2 * ctx + 1];
      S.sig0[lane] = __dmul_rn(lambda, (double)b0);
      S.sig1[lane] = __dmul_rn(lambda, (double)b1);
      S.sigd_in[lane] = b1 - b0;
    }
    const unsigned nz_mask = __ballot_sync(0xffffffffu, my_max > 0);
    __syncwarp();

    if (last_pos < 0 && nz_mask == 0) {
      // nothing significant yet: both running sums take the zero costs in scan order (:1896-1897, 2006)
#pragma unroll
      for (int k = 15; k >= 0; k--) {
        const double c0 = S.cost0[k];
        uncoded_cost = __dadd_rn(uncoded_cost, c0);
        base_cost = __dadd_rn(base_cost, c0);
      }
      __syncwarp();
      continue;
    }

    // ---- stage B: the dependent chain, warp-uniform
    double st_sig = 0.0, st_sig0 = 0.0, st_coded = 0.0, st_uncoded = 0.0;
    int st_nnz_before0 = 0;
    bool cg_nz = false;
    for (int k = 15; k >= 0; k--) {
      const int sp = (cg << 4) + k;
      const int lvl_dbl = S.lvl_dbl[k];
      const unsigned max_lvl = (unsigned)S.max_lvl[k];
      const double c0 = S.cost0[k];
      uncoded_cost = __dadd_rn(uncoded_cost, c0);
      unsigned best = 0;
      double coded = 0.0, csig = 0.0;
      int ctxw = 0, du = 0, sigd = 0;
      if (max_lvl > 0 && last_pos < 0) {
        last_pos = sp;
        st.ctx_set = (sp < 16 || !is_luma) ? 0 : 2;
        last_cg = cg;
      }
      if (last_pos >= 0) {
        const int one_ctx = 4 * st.ctx_set + st.c1, abs_ctx = st.ctx_set + st.c2;
        const bool is_last = sp == last_pos;
        double cur_sig = 0.0;
        bool decided = false;
        if (!is_last && max_lvl < 3) {
          csig = S.sig0[k];
          coded = __dadd_rn(c0, csig);
          decided = max_lvl == 0;
        } else coded = DBL_MAX;
        if (!decided) {
          if (!is_last) cur_sig = S.sig1[k];
          const unsigned min_lvl = max_lvl > 1 ? max_lvl - 1 : 1;
          for (int l = (int)max_lvl; l >= (int)min_lvl; l--) {
            const double err = (double)(lvl_dbl - (int)((unsigned)l << qbits));
            double cc = __dadd_rn(__dmul_rn(__dmul_rn(err, err), err_scale), rq_level_rate_cost(est, lambda, (unsigned)l, one_ctx, abs_ctx, st));
            cc = __dadd_rn(cc, cur_sig);
            if (cc < coded) { best = (unsigned)l; coded = cc; csig = cur_sig; }
          }
        }
        if (!is_last) sigd = S.sigd_in[k];
        du = (lvl_dbl - (int)(best << qbits)) >> (qbits - 8);
        // rateIncUp / rateIncDown (:1935-1944) are only read by sign-data hiding: keep the state they depend on and derive
        // them there, lane-parallel, for the positions that are actually examined
        ctxw = (int)(0x80000000u | (unsigned)one_ctx | ((unsigned)abs_ctx << 5) | ((unsigned)st.rice << 8) | ((unsigned)st.c1_idx << 11) |
                     ((unsigned)st.c2_idx << 16));
        base_cost = __dadd_rn(base_cost, coded);
        if ((int)best >= rq_base_level(st) && best > (3u << st.rice)) st.rice = min(st.rice + 1, 4);
        if (best >= 1) st.c1_idx++;
        if (best > 1) { st.c1 = 0; st.c2 += st.c2 < 2; st.c2_idx++; }
        else if (st.c1 < 3 && st.c1 > 0 && best) st.c1++;
        if ((sp & 15) == 0 && sp > 0) {
          st.c2 = 0; st.rice = 0; st.c1_idx = 0; st.c2_idx = 0;
          st.ctx_set = (sp == 16 || !is_luma) ? 0 : 2;
          if (st.c1 == 0) st.ctx_set++;
          st.c1 = 1;
        }
      } else base_cost = __dadd_rn(base_cost, c0);

      st_sig = __dadd_rn(st_sig, csig);
      if (k == 0) st_sig0 = csig;
      if (best) {
        cg_nz = true;
        st_coded = __dadd_rn(st_coded, __dsub_rn(coded, csig));
        st_uncoded = __dadd_rn(st_uncoded, c0);
        if (k != 0) st_nnz_before0++;
      }
      if (lane == 0) {
        S.level[k] = (int)best; S.coded[k] = coded; S.csig[k] = csig;
        S.ctxw[k] = ctxw; S.orig[k] = (int)best; S.du[k] = du; S.sigd[k] = sigd;
      }
    }
    if (cg_nz) cg_flag |= 1ull << cgpos;

    // ---- coefficient-group significance (:2025-2091)
    bool zero_out = false;
    if (last_cg >= 0) {
      if (cg) {
        const int cctx = (right | lower) ? 1 : 0;     // getSigCoeffGroupCtxInc: same neighbours as the pattern
        const double cg0 = __dmul_rn(lambda, (double)est[EB_SIG_CG + 2 * cctx]);
        const double cg1 = __dmul_rn(lambda, (double)est[EB_SIG_CG + 2 * cctx + 1]);
        double cgs = 0.0;
        bool have = false;
        if (!cg_nz) {
          base_cost = __dadd_rn(base_cost, __dsub_rn(cg0, st_sig));
          cgs = cg0; have = true;
        } else if (cg < last_cg) {
          if (st_nnz_before0 == 0) { base_cost = __dsub_rn(base_cost, st_sig0); st_sig = __dsub_rn(st_sig, st_sig0); }
          double zero_cg = base_cost;
          base_cost = __dadd_rn(base_cost, cg1);
          zero_cg = __dadd_rn(zero_cg, cg0);
          cgs = cg1; have = true;
          zero_cg = __dadd_rn(zero_cg, st_uncoded);
          zero_cg = __dsub_rn(zero_cg, st_coded);
          zero_cg = __dsub_rn(zero_cg, st_sig);
          if (zero_cg < base_cost) {
            cg_flag &= ~(1ull << cgpos);
            base_cost = zero_cg;
            cgs = cg0;
            zero_out = true;
          }
        }
        if (have && lane == 0) S.cg_sig[cg] = cgs;
        if (!have && lane == 0) S.cg_sig[cg] = 0.0;
      } else {
        cg_flag |= 1ull << cgpos;
        if (lane == 0) S.cg_sig[cg] = 0.0;
      }
    } else if (lane == 0) S.cg_sig[cg] = 0.0;
    __syncwarp();

    // ---- write the group's decisions to the per-coefficient arrays
    if (lane < 16) {
      const size_t g = off + (size_t)((cg << 4) + lane);
      int lv = S.level[lane];
      double cd = S.coded[lane], cs = S.csig[lane];
      if (zero_out && lv) { lv = 0; cd = S.cost0[lane]; cs = 0.0; }
      G.level[g] = lv; G.coded[g] = cd; G.csig[g] = cs; G.cost0[g] = S.cost0[lane];
      G.ctxw[g] = S.ctxw[lane]; G.orig[g] = S.orig[lane]; G.du[g] = S.du[lane]; G.sigd[g] = S.sigd[lane];
    }
    __syncwarp();
  }

  if (last_pos < 0) {
    // all-zero TU (:2095-2098): levels are the zeros written as uiMaxAbsLevel, uiAbsSum untouched
    for (int i = lane; i < ncoef; i += 32) levels[off + i] = 0;
    if (lane == 0 && abs_sum) abs_sum[t] = 0;
    return;
  }

  // ---- best last position (:2100-2162)
  double best_cost;
  {
    const int row = tu.cbf_ctx < 0 ? EB_ROOT : EB_CBF + 2 * tu.cbf_ctx;
    best_cost = __dadd_rn(uncoded_cost, __dmul_rn(lambda, (double)est[row]));
    base_cost = __dadd_rn(base_cost, __dmul_rn(lambda, (double)est[row + 1]));
  }
  int best_last_p1 = 0;
  bool found = false;
  for (int cg = last_cg; cg >= 0 && !found; cg--) {
    const unsigned first = scan[cg << 4];
    const int cgpos = ((int)(first >> log2) >> 2) * cgw + ((int)(first & (unsigned)(w - 1)) >> 2);
    base_cost = __dsub_rn(base_cost, S.cg_sig[cg]);
    if (!((cg_flag >> cgpos) & 1ull)) continue;
    __syncwarp();
    if (lane < 16) {
      const int sp = (cg << 4) + lane;
      const size_t g = off + (size_t)sp;
      const int lv = G.level[g];
      S.level[lane] = lv;
      S.coded[lane] = G.coded[g]; S.csig[lane] = G.csig[g]; S.cost0[lane] = G.cost0[g];
      if (lv) {
        const unsigned bp = scan[sp];
        const int py = (int)(bp >> log2), px = (int)(bp & (unsigned)(w - 1));
        S.lastc[lane] = scan_idx == 2 ? rq_last_cost(est, lambda, py, px) : rq_last_cost(est, lambda, px, py);
      }
    }
    __syncwarp();
    for (int k = 15; k >= 0; k--) {
      const int sp = (cg << 4) + k;
      if (sp > last_pos) continue;
      const int lv = S.level[k];
      if (lv) {
        const double total = __dsub_rn(__dadd_rn(base_cost, S.lastc[k]), S.csig[k]);
        if (total < best_cost) { best_last_p1 = sp + 1; best_cost = total; }
        if (lv > 1) { found = true; break; }
        base_cost = __dsub_rn(base_cost, S.coded[k]);
        base_cost = __dadd_rn(base_cost, S.cost0[k]);
      } else base_cost = __dsub_rn(base_cost, S.csig[k]);
    }
  }
  __syncwarp();

  // ---- output pass: sign, clean-up behind the chosen last position, uiAbsSum (:2164-2176)
  unsigned sum = 0;
  int top_subset = -1;
  for (int sp = lane; sp < ncoef; sp += 32) {
    int lv = 0;
    const unsigned bp = scan[sp];
    if (sp < best_last_p1) {
      lv = G.level[off + sp];
      sum += (unsigned)lv;
      if (src[bp] < 0) lv = -lv;
    }
    if (sp <= last_pos) G.level[off + sp] = lv;
    levels[off + bp] = lv;
    if (lv) top_subset = sp >> 4;
  }
#pragma unroll
  for (int m = 16; m; m >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, m);
    top_subset = max(top_subset, __shfl_xor_sync(0xffffffffu, top_subset, m));
  }
  if (lane == 0 && abs_sum) abs_sum[t] = sum;
  if (!sign_hide || sum < 2) return;
  __syncwarp();

  // ---- sign-data hiding (:2178-2304): subsets are independent, one lane each
  const double inv = (double)c_rq_invQuantScales[tu.qp_rem];
  const double rdf = __dadd_rn(__ddiv_rn(__ddiv_rn(__ddiv_rn(__dmul_rn(__dmul_rn(inv, inv), (double)(1 << (2 * tu.qp_per))), lambda), 16.0),
                                         (double)(1 << (2 * (bd - 8)))), 0.5);
  const long long rd_factor = (long long)rdf;
  for (int sub = lane; sub <= top_subset; sub += 32) {
    const int base = sub << 4;
    const int32_t* lev = G.level + off + base;
    int first_nz = 16, last_nz = -1, asum = 0;
    for (int k = 15; k >= 0; --k) if (lev[k]) { last_nz = k; break; }
    for (int k = 0; k < 16; k++) if (lev[k]) { first_nz = k; break; }
    if (last_nz - first_nz < 4) continue;
    for (int k = first_nz; k <= last_nz; k++) asum += lev[k];
    const unsigned signbit = lev[first_nz] > 0 ? 0u : 1u;
    if (signbit == (unsigned)(asum & 1)) continue;
    const bool is_top = sub == top_subset;
    long long min_cost = LLONG_MAX, cur = LLONG_MAX;
    int min_k = -1, final_change = 0, change = 0;
    for (int k = is_top ? last_nz : 15; k >= 0; --k) {
      const size_t g = off + (size_t)(base + k);
      const int lv = lev[k], du = G.du[g];
      int rup, rdn;
      rq_rate_increments(est, G.ctxw[g], G.orig[g], rup, rdn);
      if (lv != 0) {
        const bool one = abs(lv) == 1;
        const long long up = rd_factor * (long long)(-du) + rup;
        long long down = rd_factor * (long long)du + rdn - (one ? ((1 << 15) + G.sigd[g]) : 0);
        if (is_top && last_nz == k && one) down -= 4 << 15;
        if (up < down) { cur = up; change = 1; }
        else { change = -1; cur = (k == first_nz && one) ? LLONG_MAX : down; }
      } else {
        cur = rd_factor * (-(long long)abs(du)) + (1 << 15) + rup + G.sigd[g];
        change = 1;
        if (k < first_nz) {
          const unsigned s = src[scan[base + k]] >= 0 ? 0u : 1u;
          if (s != signbit) cur = LLONG_MAX;
        }
      }
      if (cur < min_cost) { min_cost = cur; final_change = change; min_k = k; }
    }
    const unsigned bp = scan[base + min_k];
    // :2283 tests the flat quantiser coefficient against +-32768: never true for g_quantScales
    if (src[bp] >= 0) levels[off + bp] += final_change; else levels[off + bp] -= final_change;
  }
}

static int ensure_rdoq_scratch(tvc_ctx* c, size_t elems, RdoqScratch& G)
{
  if (c->rdoq_scratch_elems < elems) {
    if (c->rdoq_scratch) { cudaStreamSynchronize(c->stream); cudaFree(c->rdoq_scratch); c->rdoq_scratch = nullptr; c->rdoq_scratch_elems = 0; }
    size_t cap = (elems + 4095) & ~(size_t)4095;
    TVC_CUDA(c, cudaMalloc(&c->rdoq_scratch, cap * kRdoqScratchBytesPerCoef));
    c->rdoq_scratch_elems = cap;
  }
  const size_t cap = c->rdoq_scratch_elems;
  char* p = (char*)c->rdoq_scratch;
  G.coded = (double*)p; p += cap * 8;
  G.csig = (double*)p; p += cap * 8;
  G.cost0 = (double*)p; p += cap * 8;
  G.level = (int32_t*)p; p += cap * 4;
  G.ctxw = (int32_t*)p; p += cap * 4;
  G.orig = (int32_t*)p; p += cap * 4;
  G.sigd = (int32_t*)p; p += cap * 4;
  G.du = (int32_t*)p;
  return TVC_OK;
}

static int launch_rdoq(tvc_ctx* c, int n, const tvc_rdoq_tu* tus_dev, const tvc_est_bits* est_dev, const tvc_quant_cfg& qc,
                       const int32_t* coef_dev, int32_t* levels_dev, int32_t* arl_dev, size_t coef_elems, uint32_t* abs_dev)
{
  ScanTables st;
  int r = ensure_scans(c, st);
  if (r) return r;
  RdoqScratch G;
  if ((r = ensure_rdoq_scratch(c, coef_elems, G))) return r;
  ProfScope ps(c, TVC_PH_RDOQ);
  k_rdoq<<<(n + kRdoqWarps - 1) / kRdoqWarps, kRdoqWarps * 32, 0, c->stream>>>(n, tus_dev, est_dev, qc.sign_hide, qc.use_arl, c->cfg.bit_depth, st,
                                                                             coef_dev, levels_dev, qc.use_arl ? arl_dev : nullptr, abs_dev, G);
  TVC_LAUNCH_CHECK(c);
  return TVC_OK;
}

static inline bool rdoq_record_ok(const tvc_rdoq_tu& t, int n_est, size_t coef_elems)
{
  const size_t nn = (size_t)1 << (2 * (t.log2_size & 7));
  return !(t.log2_size < 2 || t.log2_size > 5 || t.scan_idx < 0 || t.scan_idx > 2 || (t.scan_idx != 0 && t.log2_size > 3) ||
           t.qp_rem < 0 || t.qp_rem > 5 || t.qp_per < 0 || t.qp_per > 12 || t.cbf_ctx > 14 || t.est_index < 0 || t.est_index >= n_est ||
           t.coef_offset < 0 || (size_t)t.coef_offset + nn > coef_elems || !(t.lambda > 0.0) || !std::isfinite(t.lambda));
}

static int validate_rdoq(tvc_ctx* c, int n, const tvc_rdoq_tu* tus, int n_est, size_t coef_elems)
{
  for (int i = 0; i < n; i++)
    if (!rdoq_record_ok(tus[i], n_est, coef_elems)) return set_err(c, TVC_ERR_ARG, "RDOQ TU %d invalid", i);
  return TVC_OK;
}

// Both lists of a picture-sized batch in ONE pass, cut into ranges for a few host threads (2.6 x 10^5 records = 20 MB at 1080p: the
// three serial passes cost 1.5-2.5 ms of host time).  Returns false on any finding; the caller then runs the serial validators, which
// name the offending record.
static bool validate_pair_fast(const Pic& p, int n, const tvc_tu* tus, const tvc_rdoq_tu* rtus, int n_est, size_t coef_elems, int counts[4])
{
  struct Part { int counts[4] = {0, 0, 0, 0}; int first = 0, last = 0; bool ok = true; };
  // as many threads as this process may run on (a rank pinned to its share of the host cores gets its share), at most 8
  unsigned hw = std::thread::hardware_concurrency();
  cpu_set_t cpus;
  if (sched_getaffinity(0, sizeof(cpus), &cpus) == 0 && CPU_COUNT(&cpus) > 0) hw = (unsigned)CPU_COUNT(&cpus);
  const int nt = n < 32768 ? 1 : (int)std::min<unsigned>(8, hw ? hw : 1);
  std::vector<Part> parts(nt);
  auto work = [&](int k) {
    Part& P = parts[k];
    const int i0 = (int)((long long)n * k / nt), i1 = (int)((long long)n * (k + 1) / nt);
    int prev = i0 < i1 ? tus[i0].log2_size : 2;
    P.first = prev;
    for (int i = i0; i < i1; i++) {
      const tvc_tu& t = tus[i];
      const tvc_rdoq_tu& r = rtus[i];
      if (!tu_record_ok(p, t, coef_elems) || !rdoq_record_ok(r, n_est, coef_elems) || t.log2_size != r.log2_size ||
          t.coef_offset != r.coef_offset || t.log2_size < prev) { P.ok = false; return; }
      prev = t.log2_size;
      P.counts[t.log2_size - 2]++;
    }
    P.last = prev;
  };
  std::vector<std::thread> th;
  for (int k = 1; k < nt; k++) th.emplace_back(work, k);
  work(0);
  for (auto& t : th) t.join();
  counts[0] = counts[1] = counts[2] = counts[3] = 0;
  int prev = 2;
  for (int k = 0; k < nt; k++) {
    const Part& P = parts[k];
    if (!P.ok) return false;
    const bool empty = P.counts[0] + P.counts[1] + P.counts[2] + P.counts[3] == 0;
    if (!empty) { if (P.first < prev) return false; prev = P.last; }
    for (int l = 0; l < 4; l++) counts[l] += P.counts[l];
  }
  return true;
}

}  // namespace tvc

using namespace tvc;

extern "C" {

int tvc_rdoq_batch_dev(tvc_ctx* c, int n, const tvc_rdoq_tu* tus_dev, int n_est, const tvc_est_bits* est_dev, const tvc_quant_cfg* qc,
                       const int32_t* coef_dev, int32_t* levels_dev, int32_t* arl_dev, size_t coef_elems, uint32_t* abs_sum_dev)
{
  if (!c || n < 0 || !qc || n_est < 1 || (n && (!tus_dev || !est_dev || !coef_dev || !levels_dev)) || (qc->use_arl && n && !arl_dev))
    return set_err(c, TVC_ERR_ARG, "tvc_rdoq_batch_dev: bad argument");
  if (n == 0) return TVC_OK;
  return launch_rdoq(c, n, tus_dev, est_dev, *qc, coef_dev, levels_dev, arl_dev, coef_elems, abs_sum_dev);
}

int tvc_rdoq_batch(tvc_ctx* c, int n, const tvc_rdoq_tu* tus, int n_est, const tvc_est_bits* est, const tvc_quant_cfg* qc,
                   const int32_t* coef, int32_t* levels, int32_t* arl, size_t coef_elems, uint32_t* abs_sum)
{
  if (!c || n < 0 || !qc || (n && (!tus || !est || !coef || !levels || n_est < 1)) || (qc && qc->use_arl && n && !arl))
    return set_err(c, TVC_ERR_ARG, "tvc_rdoq_batch: bad argument");
  if (n == 0) return TVC_OK;
  int r;
  if ((r = validate_rdoq(c, n, tus, n_est, coef_elems))) return r;
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t tu_b = up((size_t)n * sizeof(tvc_rdoq_tu)), est_b = up((size_t)n_est * sizeof(tvc_est_bits)), coef_b = up(coef_elems * 4),
               abs_b = up((size_t)n * 4);
  const bool want_arl = qc->use_arl != 0;
  if ((r = ensure_scratch(c, c->in, tu_b + est_b + coef_b))) return r;
  if ((r = ensure_scratch(c, c->out, 2 * coef_b + abs_b + 512))) return r;
  char* hi = (char*)c->in.host;
  memcpy(hi, tus, (size_t)n * sizeof(tvc_rdoq_tu));
  memcpy(hi + tu_b, est, (size_t)n_est * sizeof(tvc_est_bits));
  memcpy(hi + tu_b + est_b, coef, coef_elems * 4);
  TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, hi, tu_b + est_b + coef_elems * 4, cudaMemcpyHostToDevice, c->stream));
  char* di = (char*)c->in.dev;
  char* dout = (char*)c->out.dev;
  if ((r = launch_rdoq(c, n, (const tvc_rdoq_tu*)di, (const tvc_est_bits*)(di + tu_b), *qc, (const int32_t*)(di + tu_b + est_b),
                       (int32_t*)dout, (int32_t*)(dout + coef_b + abs_b), coef_elems, (uint32_t*)(dout + coef_b))))
    return r;
  const size_t back = coef_b + abs_b + (want_arl ? coef_b : 0);
  TVC_CUDA(c, cudaMemcpyAsync(c->out.host, c->out.dev, back, cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  const char* ho = (const char*)c->out.host;
  // only the elements a TU of the list covers are defined
  for (int i = 0; i < n; i++) {
    const size_t nn = (size_t)1 << (2 * tus[i].log2_size), o = (size_t)tus[i].coef_offset;
    memcpy(levels + o, ho + o * 4, nn * 4);
    if (want_arl) memcpy(arl + o, ho + coef_b + abs_b + o * 4, nn * 4);
  }
  if (abs_sum) memcpy(abs_sum, ho + coef_b, (size_t)n * 4);
  return TVC_OK;
}

static int fwd_rdoq_host(tvc_ctx* c, int resi_slot, int n, const tvc_tu* tus, const tvc_rdoq_tu* rtus, int n_est, const tvc_est_bits* est,
                         const tvc_quant_cfg* qc, int32_t* levels, int32_t* arl, size_t coef_elems, uint32_t* abs_sum, int inv_resi_slot,
                         int pred_slot, int recon_slot, int16_t* levels16 = nullptr);

__global__ void k_levels_to_i16(const int32_t* __restrict__ src, int16_t* __restrict__ dst, size_t n, int* __restrict__ overflow)
{
  // two levels per thread and step: 8-byte loads, 4-byte stores
  const size_t n2 = n >> 1, stride = (size_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    const int2 v = reinterpret_cast<const int2*>(src)[i];
    bad |= v.x < -32768 || v.x > 32767 || v.y < -32768 || v.y > 32767;
    reinterpret_cast<uint32_t*>(dst)[i] = ((uint32_t)v.x & 0xffffu) | ((uint32_t)v.y << 16);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) { const int v = src[n - 1]; bad |= v < -32768 || v > 32767; dst[n - 1] = (int16_t)v; }
  if (bad) atomicOr(overflow, 1);
}

int tvc_fwd_rdoq_batch(tvc_ctx* c, int resi_slot, int n, const tvc_tu* tus, const tvc_rdoq_tu* rtus, int n_est, const tvc_est_bits* est,
                       const tvc_quant_cfg* qc, int32_t* levels, int32_t* arl, size_t coef_elems, uint32_t* abs_sum)
{
  return fwd_rdoq_host(c, resi_slot, n, tus, rtus, n_est, est, qc, levels, arl, coef_elems, abs_sum, -1, -1, -1);
}

int tvc_fwd_rdoq_recon_batch16(tvc_ctx* c, int resi_slot, int inv_resi_slot, int pred_slot, int recon_slot, int n, const tvc_tu* tus,
                               const tvc_rdoq_tu* rtus, int n_est, const tvc_est_bits* est, const tvc_quant_cfg* qc, int16_t* levels16,
                               size_t coef_elems, uint32_t* abs_sum)
{
  if (!c || !valid_slot(c, inv_resi_slot) || !valid_slot(c, pred_slot) || !valid_slot(c, recon_slot) || !levels16)
    return set_err(c, TVC_ERR_ARG, "tvc_fwd_rdoq_recon_batch16: bad argument");
  if (qc && qc->use_arl) return set_err(c, TVC_ERR_ARG, "tvc_fwd_rdoq_recon_batch16: no ARL output in the 16-bit form");
  return fwd_rdoq_host(c, resi_slot, n, tus, rtus, n_est, est, qc, (int32_t*)levels16, nullptr, coef_elems, abs_sum, inv_resi_slot, pred_slot,
                       recon_slot, levels16);
}

int tvc_fwd_rdoq_recon_batch(tvc_ctx* c, int resi_slot, int inv_resi_slot, int pred_slot, int recon_slot, int n, const tvc_tu* tus,
                             const tvc_rdoq_tu* rtus, int n_est, const tvc_est_bits* est, const tvc_quant_cfg* qc, int32_t* levels,
                             size_t coef_elems, uint32_t* abs_sum)
{
  if (!c || !valid_slot(c, inv_resi_slot) || !valid_slot(c, pred_slot) || !valid_slot(c, recon_slot))
    return set_err(c, TVC_ERR_ARG, "tvc_fwd_rdoq_recon_batch: bad slot");
  return fwd_rdoq_host(c, resi_slot, n, tus, rtus, n_est, est, qc, levels, nullptr, coef_elems, abs_sum, inv_resi_slot, pred_slot, recon_slot);
}

static int fwd_rdoq_host(tvc_ctx* c, int resi_slot, int n, const tvc_tu* tus, const tvc_rdoq_tu* rtus, int n_est, const tvc_est_bits* est,
                         const tvc_quant_cfg* qc, int32_t* levels, int32_t* arl, size_t coef_elems, uint32_t* abs_sum, int inv_resi_slot,
                         int pred_slot, int recon_slot, int16_t* levels16)
{
  if (!c || !valid_slot(c, resi_slot) || n < 0 || !qc || (n && (!tus || !rtus || !est || !levels || n_est < 1)) || (qc && qc->use_arl && n && !arl))
    return set_err(c, TVC_ERR_ARG, "tvc_fwd_rdoq_batch: bad argument");
  if (n == 0) return TVC_OK;
  int counts[4], r;
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t tu_b = up((size_t)n * sizeof(tvc_tu)), rtu_b = up((size_t)n * sizeof(tvc_rdoq_tu)), est_b = up((size_t)n_est * sizeof(tvc_est_bits)),
               coef_b = up(coef_elems * 4), abs_b = up((size_t)n * 4);
  const bool want_arl = qc->use_arl != 0;
  // in: [tus][rdoq tus][est][device-only coefficients]; out: [levels][abs sums][arl]
  if ((r = ensure_scratch(c, c->in, tu_b + rtu_b + est_b + coef_b))) return r;
  if ((r = ensure_scratch(c, c->out, 2 * coef_b + abs_b))) return r;
  // The lists go up first and are validated while the copy is in flight (a 1080p picture has 2.6 x 10^5 TUs: 20 MB of records and
  // ~1 ms of checks); page-locked caller lists are copied where they lie, pageable ones through the pinned staging area.
  if (is_pinned(tus) && is_pinned(rtus) && is_pinned(est)) {
    TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, tus, (size_t)n * sizeof(tvc_tu), cudaMemcpyHostToDevice, c->stream));
    TVC_CUDA(c, cudaMemcpyAsync((char*)c->in.dev + tu_b, rtus, (size_t)n * sizeof(tvc_rdoq_tu), cudaMemcpyHostToDevice, c->stream));
    TVC_CUDA(c, cudaMemcpyAsync((char*)c->in.dev + tu_b + rtu_b, est, (size_t)n_est * sizeof(tvc_est_bits), cudaMemcpyHostToDevice, c->stream));
  } else {
    char* hi = (char*)c->in.host;
    memcpy(hi, tus, (size_t)n * sizeof(tvc_tu));
    memcpy(hi + tu_b, rtus, (size_t)n * sizeof(tvc_rdoq_tu));
    memcpy(hi + tu_b + rtu_b, est, (size_t)n_est * sizeof(tvc_est_bits));
    TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, hi, tu_b + rtu_b + est_b, cudaMemcpyHostToDevice, c->stream));
  }
  r = TVC_OK;
  if (!validate_pair_fast(c->pics[resi_slot], n, tus, rtus, n_est, coef_elems, counts)) {
    r = validate_tus(c, resi_slot, n, tus, coef_elems, counts);
    if (!r) r = validate_rdoq(c, n, rtus, n_est, coef_elems);
    for (int i = 0; i < n && !r; i++)
      if (tus[i].log2_size != rtus[i].log2_size || tus[i].coef_offset != rtus[i].coef_offset)
        r = set_err(c, TVC_ERR_ARG, "tvc_fwd_rdoq_batch: TU %d of the two lists differ", i);
    if (!r) r = set_err(c, TVC_ERR_ARG, "tvc_fwd_rdoq_batch: invalid TU list");
  }
  if (r) { cudaStreamSynchronize(c->stream); return r; }      // nothing but the copy was queued
  char* di = (char*)c->in.dev;
  char* dout = (char*)c->out.dev;
  int32_t* d_coef = (int32_t*)(di + tu_b + rtu_b + est_b);
  if ((r = tvc_fwd_transform_batch_dev(c, resi_slot, n, (const tvc_tu*)di, counts, d_coef))) return r;
  if ((r = launch_rdoq(c, n, (const tvc_rdoq_tu*)(di + tu_b), (const tvc_est_bits*)(di + tu_b + rtu_b), *qc, d_coef, (int32_t*)dout,
                       (int32_t*)(dout + coef_b + abs_b), coef_elems, (uint32_t*)(dout + coef_b))))
    return r;
  // round trip: dequant + inverse transform + reconstruction from the levels where they are
  if (inv_resi_slot >= 0 && (r = launch_inv(c, inv_resi_slot, pred_slot, recon_slot, counts, (const tvc_tu*)di, (const int32_t*)dout, 1))) return r;
  if (levels16) {
    // the levels are what CABAC codes: 16-bit values (TComTrQuant clips them to [-32768, 32767], TComTrQuant.cpp:1298); they go back
    // as int16 -- half the bytes of the largest device-to-host transfer of a picture.  A level beyond 16 bits raises an error.
    int16_t* d16 = (int16_t*)(dout + coef_b + abs_b);              // the ARL area is unused in this form
    int* d_flag = (int*)(dout + coef_b + abs_b + up(coef_elems * 2));
    TVC_CUDA(c, cudaMemsetAsync(d_flag, 0, sizeof(int), c->stream));
    k_levels_to_i16<<<kNumSM * 8, 256, 0, c->stream>>>((const int32_t*)dout, d16, coef_elems, d_flag);
    TVC_LAUNCH_CHECK(c);
    int* h_flag = (int*)((char*)c->out.host + coef_b + abs_b + up(coef_elems * 2));
    TVC_CUDA(c, cudaMemcpyAsync(h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (is_pinned(levels16) && (!abs_sum || is_pinned(abs_sum))) {
      TVC_CUDA(c, cudaMemcpyAsync(levels16, d16, coef_elems * 2, cudaMemcpyDeviceToHost, c->stream));
      if (abs_sum) TVC_CUDA(c, cudaMemcpyAsync(abs_sum, dout + coef_b, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
      TVC_CUDA(c, cudaStreamSynchronize(c->stream));
    } else {
      TVC_CUDA(c, cudaMemcpyAsync((char*)c->out.host + coef_b, dout + coef_b, abs_b + coef_elems * 2, cudaMemcpyDeviceToHost, c->stream));
      TVC_CUDA(c, cudaStreamSynchronize(c->stream));
      memcpy(levels16, (char*)c->out.host + coef_b + abs_b, coef_elems * 2);
      if (abs_sum) memcpy(abs_sum, (char*)c->out.host + coef_b, (size_t)n * 4);
    }
    if (*h_flag) return set_err(c, TVC_ERR_ARG, "tvc_fwd_rdoq_recon_batch16: a level does not fit 16 bits (use the 32-bit form)");
    return TVC_OK;
  }
  const bool pin_out = is_pinned(levels) && (!want_arl || is_pinned(arl)) && (!abs_sum || is_pinned(abs_sum));
  if (pin_out) {
    TVC_CUDA(c, cudaMemcpyAsync(levels, dout, coef_elems * 4, cudaMemcpyDeviceToHost, c->stream));
    if (want_arl) TVC_CUDA(c, cudaMemcpyAsync(arl, dout + coef_b + abs_b, coef_elems * 4, cudaMemcpyDeviceToHost, c->stream));
    if (abs_sum) TVC_CUDA(c, cudaMemcpyAsync(abs_sum, dout + coef_b, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    TVC_CUDA(c, cudaStreamSynchronize(c->stream));
    return TVC_OK;
  }
  TVC_CUDA(c, cudaMemcpyAsync(c->out.host, c->out.dev, coef_b + abs_b + (want_arl ? coef_b : 0), cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  const char* ho = (const char*)c->out.host;
  memcpy(levels, ho, coef_elems * 4);
  if (want_arl) memcpy(arl, ho + coef_b + abs_b, coef_elems * 4);
  if (abs_sum) memcpy(abs_sum, ho + coef_b, (size_t)n * 4);
  return TVC_OK;
}

int tvc_xRateDistOptQuant(tvc_ctx* c, const int32_t* coef, int32_t* qcoef, int32_t* arl, int w, int is_luma, int scan_idx, int qp_per,
                          int qp_rem, int cbf_ctx, int sign_hide, int use_arl, double lambda, const tvc_est_bits* est, uint32_t* abs_sum)
{
  if (!c || (w != 4 && w != 8 && w != 16 && w != 32)) return set_err(c, TVC_ERR_ARG, "tvc_xRateDistOptQuant: unsupported size");
  tvc_rdoq_tu tu;
  memset(&tu, 0, sizeof(tu));
  tu.log2_size = w == 4 ? 2 : w == 8 ? 3 : w == 16 ? 4 : 5;
  tu.is_luma = is_luma; tu.scan_idx = scan_idx; tu.qp_per = qp_per; tu.qp_rem = qp_rem; tu.cbf_ctx = cbf_ctx;
  tu.est_index = 0; tu.coef_offset = 0; tu.lambda = lambda;
  tvc_quant_cfg qc = {0, sign_hide, use_arl};
  uint32_t s = 0;
  int r = tvc_rdoq_batch(c, 1, &tu, 1, est, &qc, coef, qcoef, arl, (size_t)w * w, &s);
  if (r) return r;
  if (abs_sum) *abs_sum += s;      // uiAbsSum accumulates in the reference (:2168)
  return TVC_OK;
}

}  // extern "C"
