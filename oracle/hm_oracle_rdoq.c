/*
 * hm_oracle_rdoq.c -- CPU restatement of HM-7.2 rate-distortion optimised quantisation
 * (TComTrQuant::xRateDistOptQuant and its helpers).  TEST INFRASTRUCTURE ONLY (see hm_oracle.h).
 * Citations: /root/reference/source/Lib/TLibCommon/TComTrQuant.cpp unless stated otherwise.
 *
 * Compile-time switches of the reference that shape this code (TypeDef.h): REMOVE_NSQT 1 (square
 * TUs only), REMOVAL_8x2_2x8_CG 1 (every coefficient group is 4x4), REMOVE_NUM_GREATER1 1,
 * COEF_REMAIN_BIN_REDUCTION 3, C1FLAG_NUMBER 8, C2FLAG_NUMBER 1, SBH_THRESHOLD 4,
 * ADAPTIVE_QP_SELECTION 1 (ARL_C_PRECISION 7), flat scaling lists (ScalingList 0 in every cfg).
 *
 * All costs are IEEE doubles evaluated in the reference's order of operations (the reference is
 * built for x86-64/SSE2 without FMA contraction; oracle/Makefile passes -ffp-contract=off).
 */
#include "hm_oracle.h"
#include <float.h>
#include <string.h>

static const int k_quant_scales[6] = { 26214, 23302, 20560, 18396, 16384, 14564 };   /* TComRom.cpp:293-296 */
static const int k_inv_quant_scales[6] = { 40, 45, 51, 57, 64, 72 };                 /* TComRom.cpp:298-301 */
static const uint8_t k_group_idx[32] = { 0, 1, 2, 3, 4, 4, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7,          /* TComRom.cpp:504 */
                                         8, 8, 8, 8, 8, 8, 8, 8, 9, 9, 9, 9, 9, 9, 9, 9 };
static const int k_rice_range[5] = { 7, 14, 26, 46, 78 };                             /* TComRom.cpp:507-510 */
static const int k_rice_prefix[5] = { 8, 7, 6, 5, 4 };                                /* TComRom.cpp:512-515 */

static inline int iabs(int v) { return v < 0 ? -v : v; }

/* setErrScaleCoeff :2794-2817 with the flat quantiser coefficient (xsetFlatScalingList :2904-2920) */
double orc_rdoq_err_scale(int log2_size, int qp_rem, int bd)
{
  int tshift = 15 - bd - log2_size;
  double q = (double)k_quant_scales[qp_rem];
  double s = (double)(1 << 15);
  /* pow(2.0, -2.0*tshift) is an exact power of two */
  int e = -2 * tshift;
  double p2 = 1.0;
  for (int i = 0; i < (e < 0 ? -e : e); i++) p2 = e < 0 ? p2 * 0.5 : p2 * 2.0;
  s = s * p2;
  return s / q / q / (double)(1 << (2 * (bd - 8)));
}

/* state of the level-flag contexts while walking the scan backwards (:1827-1838) */
typedef struct { int ctx_set, c1, c2, rice, c1_idx, c2_idx; } lvl_state;

static inline int base_level(const lvl_state* s) { return s->c1_idx < 8 ? (2 + (s->c2_idx < 1)) : 1; }

/* xGetICRate :2577-2636 -- integer rate of an absolute level (scaled by 2^15) */
static int level_rate_int(const orc_est_bits* est, int lvl, int one_ctx, int abs_ctx, const lvl_state* s)
{
  int base = base_level(s), rate = 0;
  if (lvl >= base) {
    unsigned sym = (unsigned)(lvl - base);
    unsigned max_vlc = (unsigned)k_rice_range[s->rice];
    if (sym > max_vlc) {
      unsigned rest = sym - max_vlc;
      int egs = 1;
      for (unsigned m = 2; rest >= m; m <<= 1) egs += 2;
      rate += egs << 15;
      sym = sym < max_vlc + 1 ? sym : max_vlc + 1;
    }
    unsigned pref = (sym >> s->rice) + 1;
    unsigned cap = (unsigned)k_rice_prefix[s->rice];
    rate += (int)(((pref < cap ? pref : cap) + (unsigned)s->rice) & 0xffffu) << 15;
    if (s->c1_idx < 8) {
      rate += est->greater_one[one_ctx][1];
      if (s->c2_idx < 1) rate += est->level_abs[abs_ctx][1];
    }
  } else if (lvl == 1) rate += est->greater_one[one_ctx][0];
  else if (lvl == 2) rate += est->greater_one[one_ctx][1] + est->level_abs[abs_ctx][0];
  else return 0;      /* lvl == 0 */
  return rate;
}

/* xGetICRateCost :2508-2575 -- lambda * rate including the sign bin (xGetIEPRate = 32768) */
static double level_rate_cost(const orc_est_bits* est, double lambda, unsigned lvl, int one_ctx, int abs_ctx, const lvl_state* s)
{
  double rate = 32768.0;
  unsigned base = (unsigned)base_level(s);
  if (lvl >= base) {
    unsigned sym = lvl - base, len;
    if (sym < (3u << s->rice)) {
      len = sym >> s->rice;
      rate += (double)(int)((len + 1 + (unsigned)s->rice) << 15);
    } else {
      len = (unsigned)s->rice;
      sym -= 3u << s->rice;
      while (sym >= (1u << len)) sym -= 1u << (len++);
      rate += (double)(int)((3 + len + 1 - (unsigned)s->rice + len) << 15);
    }
    if (s->c1_idx < 8) {
      rate += (double)est->greater_one[one_ctx][1];
      if (s->c2_idx < 1) rate += (double)est->level_abs[abs_ctx][1];
    }
  } else if (lvl == 1) rate += (double)est->greater_one[one_ctx][0];
  else {               /* lvl == 2 */
    rate += (double)est->greater_one[one_ctx][1];
    rate += (double)est->level_abs[abs_ctx][0];
  }
  return lambda * rate;
}

/* getSigCtxInc :2349-2428 (REMOVAL_8x2_2x8_CG branch); scan_idx: 0 diag, 1 hor, 2 ver */
static int sig_ctx_inc(int pattern, int scan_idx, int px, int py, int log2_size, int is_luma)
{
  static const uint8_t map4[16] = { 0, 1, 4, 5, 2, 3, 4, 5, 6, 6, 8, 8, 7, 7, 8, 8 };
  if (px + py == 0) return 0;
  if (log2_size == 2) return map4[4 * py + px];
  int offset = log2_size == 3 ? (scan_idx == 0 ? 9 : 15) : (is_luma ? 21 : 12);
  int sx = px & 3, sy = py & 3, cnt;
  if (pattern == 0) cnt = sx + sy <= 2 ? (sx + sy == 0 ? 2 : 1) : 0;
  else if (pattern == 1) cnt = sy <= 1 ? (sy == 0 ? 2 : 1) : 0;
  else if (pattern == 2) cnt = sx <= 1 ? (sx == 0 ? 2 : 1) : 0;
  else cnt = 2;
  return ((is_luma && ((px >> 2) + (py >> 2)) > 0) ? 3 : 0) + offset + cnt;
}

/* xGetRateLast :2652-2668 */
static double last_pos_cost(const orc_est_bits* est, double lambda, int px, int py)
{
  unsigned cx = k_group_idx[px], cy = k_group_idx[py];
  double r = (double)(est->last_x[cx] + est->last_y[cy]);
  if (cx > 3) r += 32768.0 * (double)((cx - 2) >> 1);
  if (cy > 3) r += 32768.0 * (double)((cy - 2) >> 1);
  return lambda * r;
}

/* xRateDistOptQuant :1719-2305.  coef: w*w raster of xT output; qcoef: levels out; arl: ARL
 * coefficients (written when p->use_arl) or NULL; scan: the coding scan of (scan_idx, log2).
 * *abs_sum accumulates like uiAbsSum (the caller zeroes it, :1394).                           */
void orc_rdoq(const int32_t* coef, int32_t* qcoef, int32_t* arl, const orc_rdoq_param* p,
              const orc_est_bits* est, const uint32_t* scan, uint32_t* abs_sum)
{
  const int log2 = p->log2_size, w = 1 << log2, ncoef = w * w, ncg = ncoef >> 4, cgw = w >> 2;
  const int bi = p->bd - 8;
  const int tshift = 15 - p->bd - log2;
  const int qbits = 14 + p->qp_per + tshift;                       /* :1757 */
  const int qscale = k_quant_scales[p->qp_rem];
  const double err_scale = orc_rdoq_err_scale(log2, p->qp_rem, p->bd);
  const double lambda = p->lambda;
  const int qbitsC = qbits - 7, addC = 1 << (qbitsC - 1);          /* :1764-1765 */

  double cost_coded[1024], cost_sig[1024], cost_zero[1024], cost_cg_sig[64];
  int rate_up[1024], rate_down[1024], sig_delta[1024], delta_u[1024];
  unsigned cg_flag[64];
  memset(cost_coded, 0, sizeof(double) * (size_t)ncoef);
  memset(cost_sig, 0, sizeof(double) * (size_t)ncoef);
  memset(rate_up, 0, sizeof(int) * (size_t)ncoef);
  memset(rate_down, 0, sizeof(int) * (size_t)ncoef);
  memset(sig_delta, 0, sizeof(int) * (size_t)ncoef);
  memset(delta_u, 0, sizeof(int) * (size_t)ncoef);
  memset(cost_cg_sig, 0, sizeof(cost_cg_sig));
  memset(cg_flag, 0, sizeof(cg_flag));
  if (arl) memset(arl, 0, sizeof(int32_t) * (size_t)ncoef);      /* :1783 (the reference clears it whether or not ARL is on) */

  lvl_state st = { 0, 1, 0, 0, 0, 0 };
  double base_cost = 0.0, uncoded_cost = 0.0;
  int last_pos = -1, last_cg = -1;

  for (int cg = ncg - 1; cg >= 0; cg--) {
    /* REMOVAL_8x2_2x8_CG: the 16 scan positions of a group lie in one 4x4 block; the group's block
     * position (scanCG[], :1800-1814, 1864-1866) is therefore that of its first coefficient */
    const unsigned first = scan[cg << 4];
    const int cgy = (int)(first >> log2) >> 2, cgx = (int)(first & (unsigned)(w - 1)) >> 2;
    const int cgpos = cgy * cgw + cgx;
    /* calcPatternSigCtx :2315-2337 */
    int pattern = -1;
    if (log2 != 2) {
      int right = 0, lower = 0;
      if (cgx < cgw - 1) right = cg_flag[cgy * cgw + cgx + 1] != 0;
      if (cgy < cgw - 1) lower = cg_flag[(cgy + 1) * cgw + cgx] != 0;
      pattern = right + (lower << 1);
    }
    /* coeffGroupRDStats */
    double st_sig = 0.0, st_sig0 = 0.0, st_coded = 0.0, st_uncoded = 0.0;
    int st_nnz_before0 = 0;

    for (int k = 15; k >= 0; k--) {
      const int sp = (cg << 4) + k;
      const unsigned bp = scan[sp];
      int64_t scaled = (int64_t)iabs(coef[bp]) * qscale;
      const int64_t cap = 2147483647LL - (1LL << (qbits - 1));
      const int lvl_dbl = (int)(scaled < cap ? scaled : cap);    /* :1886 */
      if (p->use_arl && arl) arl[bp] = (lvl_dbl + addC) >> qbitsC;
      const unsigned max_lvl = (unsigned)((lvl_dbl + (1 << (qbits - 1))) >> qbits);
      const double e0 = (double)lvl_dbl;
      cost_zero[sp] = e0 * e0 * err_scale;
      uncoded_cost += cost_zero[sp];
      qcoef[bp] = (int32_t)max_lvl;

      if (max_lvl > 0 && last_pos < 0) {
        last_pos = sp;
        st.ctx_set = (sp < 16 || !p->is_luma) ? 0 : 2;
        last_cg = cg;
      }
      if (last_pos >= 0) {
        const int one_ctx = 4 * st.ctx_set + st.c1, abs_ctx = st.ctx_set + st.c2;
        const int is_last = sp == last_pos;
        int sig_ctx = 0;
        if (!is_last) sig_ctx = sig_ctx_inc(pattern, p->scan_idx, (int)(bp & (unsigned)(w - 1)), (int)(bp >> log2), log2, p->is_luma);
        /* xGetCodedLevel :2446-2499 */
        unsigned best = 0;
        double cur_sig = 0.0;
        int decided = 0;
        if (!is_last && max_lvl < 3) {
          cost_sig[sp] = lambda * (double)est->sig[sig_ctx][0];
          cost_coded[sp] = cost_zero[sp] + cost_sig[sp];
          if (max_lvl == 0) decided = 1;
        } else cost_coded[sp] = DBL_MAX;
        if (!decided) {
          if (!is_last) cur_sig = lambda * (double)est->sig[sig_ctx][1];
          const unsigned min_lvl = max_lvl > 1 ? max_lvl - 1 : 1;
          for (int l = (int)max_lvl; l >= (int)min_lvl; l--) {
            double err = (double)(lvl_dbl - (int)((unsigned)l << qbits));
            double c = err * err * err_scale + level_rate_cost(est, lambda, (unsigned)l, one_ctx, abs_ctx, &st);
            c += cur_sig;
            if (c < cost_coded[sp]) { best = (unsigned)l; cost_coded[sp] = c; cost_sig[sp] = cur_sig; }
          }
        }
        if (!is_last) sig_delta[bp] = est->sig[sig_ctx][1] - est->sig[sig_ctx][0];
        delta_u[bp] = (lvl_dbl - (int)(best << qbits)) >> (qbits - 8);
        if (best > 0) {
          int now = level_rate_int(est, (int)best, one_ctx, abs_ctx, &st);
          rate_up[bp] = level_rate_int(est, (int)best + 1, one_ctx, abs_ctx, &st) - now;
          rate_down[bp] = level_rate_int(est, (int)best - 1, one_ctx, abs_ctx, &st) - now;
        } else rate_up[bp] = est->greater_one[one_ctx][0];
        qcoef[bp] = (int32_t)best;
        base_cost += cost_coded[sp];

        /* Rice parameter and flag-context updates :1949-2002 */
        if ((int)best >= base_level(&st) && best > (3u << st.rice)) st.rice = st.rice + 1 < 4 ? st.rice + 1 : 4;
        if (best >= 1) st.c1_idx++;
        if (best > 1) { st.c1 = 0; st.c2 += st.c2 < 2; st.c2_idx++; }
        else if (st.c1 < 3 && st.c1 > 0 && best) st.c1++;
        if ((sp & 15) == 0 && sp > 0) {
          st.c2 = 0; st.rice = 0; st.c1_idx = 0; st.c2_idx = 0;
          st.ctx_set = (sp == 16 || !p->is_luma) ? 0 : 2;
          if (st.c1 == 0) st.ctx_set++;
          st.c1 = 1;
        }
      } else base_cost += cost_zero[sp];

      st_sig += cost_sig[sp];
      if (k == 0) st_sig0 = cost_sig[sp];
      if (qcoef[bp]) {
        cg_flag[cgpos] = 1;
        st_coded += cost_coded[sp] - cost_sig[sp];
        st_uncoded += cost_zero[sp];
        if (k != 0) st_nnz_before0++;
      }
    }

    /* coefficient-group significance decision :2025-2091 */
    if (last_cg >= 0) {
      if (cg) {
        /* getSigCoeffGroupCtxInc :2707-2743 */
        int right = 0, lower = 0;
        if (cgx < cgw - 1) right = cg_flag[cgy * cgw + cgx + 1] != 0;
        if (cgy < cgw - 1) lower = cg_flag[(cgy + 1) * cgw + cgx] != 0;
        const int cctx = right || lower;
        if (cg_flag[cgpos] == 0) {
          base_cost += lambda * (double)est->sig_cg[cctx][0] - st_sig;
          cost_cg_sig[cg] = lambda * (double)est->sig_cg[cctx][0];
        } else if (cg < last_cg) {
          if (st_nnz_before0 == 0) { base_cost -= st_sig0; st_sig -= st_sig0; }
          double zero_cg = base_cost;
          base_cost += lambda * (double)est->sig_cg[cctx][1];
          zero_cg += lambda * (double)est->sig_cg[cctx][0];
          cost_cg_sig[cg] = lambda * (double)est->sig_cg[cctx][1];
          zero_cg += st_uncoded;
          zero_cg -= st_coded;
          zero_cg -= st_sig;
          if (zero_cg < base_cost) {
            cg_flag[cgpos] = 0;
            base_cost = zero_cg;
            cost_cg_sig[cg] = lambda * (double)est->sig_cg[cctx][0];
            for (int k = 15; k >= 0; k--) {
              const int sp = (cg << 4) + k;
              const unsigned bp = scan[sp];
              if (qcoef[bp]) { qcoef[bp] = 0; cost_coded[sp] = cost_zero[sp]; cost_sig[sp] = 0.0; }
            }
          }
        }
      } else cg_flag[cgpos] = 1;
    }
  }

  if (last_pos < 0) return;                                        /* :2095-2098 */

  /* coded-block-flag cost and the best last position :2100-2162 */
  double best_cost;
  if (p->cbf_ctx < 0) {
    best_cost = uncoded_cost + lambda * (double)est->block_root_cbp[0][0];
    base_cost += lambda * (double)est->block_root_cbp[0][1];
  } else {
    best_cost = uncoded_cost + lambda * (double)est->block_cbp[p->cbf_ctx][0];
    base_cost += lambda * (double)est->block_cbp[p->cbf_ctx][1];
  }
  int best_last_p1 = 0, found = 0;
  for (int cg = last_cg; cg >= 0 && !found; cg--) {
    const unsigned first = scan[cg << 4];
    const int cgpos = ((int)(first >> log2) >> 2) * cgw + ((int)(first & (unsigned)(w - 1)) >> 2);
    base_cost -= cost_cg_sig[cg];
    if (!cg_flag[cgpos]) continue;
    for (int k = 15; k >= 0; k--) {
      const int sp = (cg << 4) + k;
      if (sp > last_pos) continue;
      const unsigned bp = scan[sp];
      if (qcoef[bp]) {
        const int py = (int)(bp >> log2), px = (int)(bp & (unsigned)(w - 1));
        double cl = p->scan_idx == 2 ? last_pos_cost(est, lambda, py, px) : last_pos_cost(est, lambda, px, py);
        double total = base_cost + cl - cost_sig[sp];
        if (total < best_cost) { best_last_p1 = sp + 1; best_cost = total; }
        if (qcoef[bp] > 1) { found = 1; break; }
        base_cost -= cost_coded[sp];
        base_cost += cost_zero[sp];
      } else base_cost -= cost_sig[sp];
    }
  }

  uint32_t sum = *abs_sum;
  for (int sp = 0; sp < best_last_p1; sp++) {
    const unsigned bp = scan[sp];
    int lvl = qcoef[bp];
    sum += (uint32_t)lvl;
    qcoef[bp] = coef[bp] < 0 ? -lvl : lvl;
  }
  for (int sp = best_last_p1; sp <= last_pos; sp++) qcoef[scan[sp]] = 0;
  *abs_sum = sum;

  /* sign-data hiding on the RDOQ levels :2178-2304 */
  if (p->sign_hide && sum >= 2) {
    const double inv = (double)k_inv_quant_scales[p->qp_rem];
    const int64_t rd_factor = (int64_t)(inv * inv * (double)(1 << (2 * p->qp_per)) / lambda / 16 / (double)(1 << (2 * bi)) + 0.5);
    int last_flag = -1;
    for (int sub = (ncoef - 1) >> 4; sub >= 0; sub--) {
      const int base = sub << 4;
      int first_nz = 16, last_nz = -1, asum = 0, n;
      for (n = 15; n >= 0; --n) if (qcoef[scan[n + base]]) { last_nz = n; break; }
      for (n = 0; n < 16; n++) if (qcoef[scan[n + base]]) { first_nz = n; break; }
      for (n = first_nz; n <= last_nz; n++) asum += qcoef[scan[n + base]];
      if (last_nz >= 0 && last_flag == -1) last_flag = 1;
      if (last_nz - first_nz >= 4) {
        const unsigned signbit = qcoef[scan[base + first_nz]] > 0 ? 0u : 1u;
        if (signbit != (unsigned)(asum & 1)) {
          int64_t min_cost = INT64_MAX, cur = INT64_MAX;
          int min_pos = -1, final_change = 0, change = 0;
          for (n = (last_flag == 1 ? last_nz : 15); n >= 0; --n) {
            const unsigned bp = scan[n + base];
            if (qcoef[bp] != 0) {
              int64_t up = rd_factor * (-delta_u[bp]) + rate_up[bp];
              int64_t down = rd_factor * delta_u[bp] + rate_down[bp] - (iabs(qcoef[bp]) == 1 ? ((1 << 15) + sig_delta[bp]) : 0);
              if (last_flag == 1 && last_nz == n && iabs(qcoef[bp]) == 1) down -= 4 << 15;
              if (up < down) { cur = up; change = 1; }
              else {
                change = -1;
                cur = (n == first_nz && iabs(qcoef[bp]) == 1) ? INT64_MAX : down;
              }
            } else {
              cur = rd_factor * (-(int64_t)iabs(delta_u[bp])) + (1 << 15) + rate_up[bp] + sig_delta[bp];
              change = 1;
              if (n < first_nz) {
                const unsigned s = coef[bp] >= 0 ? 0u : 1u;
                if (s != signbit) cur = INT64_MAX;
              }
            }
            if (cur < min_cost) { min_cost = cur; final_change = change; min_pos = (int)bp; }
          }
          /* :2283 compares the flat QUANTISER coefficient (never +-32768) -- no effect with flat lists */
          if (coef[min_pos] >= 0) qcoef[min_pos] += final_change; else qcoef[min_pos] -= final_change;
        }
      }
      if (last_flag == 1) last_flag = 0;
    }
  }
}
