/*
 * hm_oracle_deblock.c -- CPU restatement of the HM-7.2 deblocking filter's sample work.
 * TEST INFRASTRUCTURE ONLY (see hm_oracle.h).  Citations: /root/reference/source/Lib/TLibCommon/TComLoopFilter.cpp.
 *
 * The boundary strengths are an input (they come out of the CU tree, :266-569); the unit records are the ones
 * include/thevc_cuda.h describes.  Pinned by tests/golden/deblock_golden.npz: pictures before / after the reference's
 * own loopFilterPic inside the reference encoder, with the unit records it acted on.
 */
#include "hm_oracle.h"
#include <stddef.h>
#include <stdlib.h>

static const uint8_t k_tc[54] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,5,5,6,6,7,8,9,10,11,13,14,16,18,20,22,24 };   /* :56-59 */
static const uint8_t k_beta[52] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,6,7,8,9,10,11,12,13,14,15,16,17,18,20,22,24,26,28,30,32,34,36,38,40,42,44,46,48,50,52,54,56,58,60,62,64 };   /* :61-64 */
static const uint8_t k_chroma_scale[58] = { 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,29,30,31,32,
                                            33,33,34,34,35,35,36,36,37,37,38,39,40,41,42,43,44,45,46,47,48,49,50,51 };   /* TComRom.cpp:380-386 */

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

/* xPelFilterLuma :799-867 on one line; s points at m4, `o` is the step across the edge */
static void luma_line(Pel* s, int o, int tc, int sw, int keep_p, int keep_q, int thr_cut, int fp, int fq, int maxv)
{
  int m4 = s[0], m3 = s[-o], m5 = s[o], m2 = s[-2 * o], m6 = s[2 * o], m1 = s[-3 * o], m7 = s[3 * o], m0 = s[-4 * o];
  if (sw) {
    s[-o] = (Pel)clip3(m3 - 2 * tc, m3 + 2 * tc, (m1 + 2 * m2 + 2 * m3 + 2 * m4 + m5 + 4) >> 3);
    s[0] = (Pel)clip3(m4 - 2 * tc, m4 + 2 * tc, (m2 + 2 * m3 + 2 * m4 + 2 * m5 + m6 + 4) >> 3);
    s[-2 * o] = (Pel)clip3(m2 - 2 * tc, m2 + 2 * tc, (m1 + m2 + m3 + m4 + 2) >> 2);
    s[o] = (Pel)clip3(m5 - 2 * tc, m5 + 2 * tc, (m3 + m4 + m5 + m6 + 2) >> 2);
    s[-3 * o] = (Pel)clip3(m1 - 2 * tc, m1 + 2 * tc, (2 * m0 + 3 * m1 + m2 + m3 + m4 + 4) >> 3);
    s[2 * o] = (Pel)clip3(m6 - 2 * tc, m6 + 2 * tc, (m3 + m4 + m5 + 3 * m6 + 2 * m7 + 4) >> 3);
  } else {
    int delta = (9 * (m4 - m3) - 3 * (m5 - m2) + 8) >> 4;
    if (abs(delta) < thr_cut) {
      delta = clip3(-tc, tc, delta);
      s[-o] = (Pel)clip3(0, maxv, m3 + delta);
      s[0] = (Pel)clip3(0, maxv, m4 - delta);
      int tc2 = tc >> 1;
      if (fp) s[-2 * o] = (Pel)clip3(0, maxv, m2 + clip3(-tc2, tc2, ((((m1 + m3 + 1) >> 1) - m2 + delta) >> 1)));
      if (fq) s[o] = (Pel)clip3(0, maxv, m5 + clip3(-tc2, tc2, ((((m6 + m4 + 1) >> 1) - m5 - delta) >> 1)));
    }
  }
  if (keep_p) { s[-o] = (Pel)m3; s[-2 * o] = (Pel)m2; s[-3 * o] = (Pel)m1; }
  if (keep_q) { s[0] = (Pel)m4; s[o] = (Pel)m5; s[2 * o] = (Pel)m6; }
}

static int strong(const Pel* s, int o, int d, int beta, int tc)      /* xUseStrongFiltering :901-911 */
{
  int m4 = s[0], m3 = s[-o], m7 = s[3 * o], m0 = s[-4 * o];
  int ds = abs(m0 - m3) + abs(m7 - m4);
  return ds < (beta >> 3) && d < (beta >> 2) && abs(m3 - m4) < ((tc * 5 + 1) >> 1);
}
static int calc_dp(const Pel* s, int o) { return abs(s[-3 * o] - 2 * s[-2 * o] + s[-o]); }      /* :913 */
static int calc_dq(const Pel* s, int o) { return abs(s[0] - 2 * s[o] + s[2 * o]); }            /* :918 */

/* one edge unit: 4 luma lines (xEdgeFilterLuma :608-676) and, on the 16-pel grid with bs > 1, 2 chroma lines per plane
 * (xEdgeFilterChroma :746-795).  (x, y): luma position of the unit's first Q sample; dir 0 vertical edge, 1 horizontal */
static void filter_unit(Pel* Y, int sy, Pel* U, Pel* V, int sc, int x, int y, int dir, const orc_dbk_unit* u, int beta_off2,
                        int tc_off2, int bd)
{
  const int scale = 1 << (bd - 8), maxv = (1 << bd) - 1;
  const int keep_p = u->flags & 1, keep_q = (u->flags >> 1) & 1;
  {
    const int o = dir == 0 ? 1 : sy, step = dir == 0 ? sy : 1;
    Pel* s = Y + (ptrdiff_t)y * sy + x;
    int idx_tc = clip3(0, 53, u->qp + 2 * (u->bs - 1) + (tc_off2 << 1));
    int idx_b = clip3(0, 51, u->qp + (beta_off2 << 1));
    int tc = k_tc[idx_tc] * scale, beta = k_beta[idx_b] * scale;
    int side = (beta + (beta >> 1)) >> 3, thr_cut = tc * 10;
    int dp0 = calc_dp(s, o), dq0 = calc_dq(s, o), dp3 = calc_dp(s + 3 * step, o), dq3 = calc_dq(s + 3 * step, o);
    int d0 = dp0 + dq0, d3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3, d = d0 + d3;
    if (d < beta) {
      int fp = dp < side, fq = dq < side;
      int sw = strong(s, o, 2 * d0, beta, tc) && strong(s + 3 * step, o, 2 * d3, beta, tc);
      for (int i = 0; i < 4; i++) luma_line(s + i * step, o, tc, sw, keep_p, keep_q, thr_cut, fp, fq, maxv);
    }
  }
  if (u->bs > 1 && ((dir == 0 ? x : y) & 15) == 0) {
    const int o = dir == 0 ? 1 : sc, step = dir == 0 ? sc : 1;
    int qpc = k_chroma_scale[clip3(0, 51, u->qp)];
    int tc = k_tc[clip3(0, 53, qpc + 2 * (u->bs - 1) + (tc_off2 << 1))] * scale;
    for (int pl = 0; pl < 2; pl++) {
      Pel* c = (pl ? V : U) + (ptrdiff_t)(y >> 1) * sc + (x >> 1);
      for (int i = 0; i < 2; i++) {                                  /* xPelFilterChroma :869-892 */
        Pel* s = c + i * step;
        int m4 = s[0], m3 = s[-o], m5 = s[o], m2 = s[-2 * o];
        int delta = clip3(-tc, tc, ((((m4 - m3) << 2) + m2 - m5 + 4) >> 3));
        if (!keep_p) s[-o] = (Pel)clip3(0, maxv, m3 + delta);
        if (!keep_q) s[0] = (Pel)clip3(0, maxv, m4 - delta);
      }
    }
  }
}

/* loopFilterPic :153-191: every vertical edge of the picture, then every horizontal edge.  Planes point at pel (0,0). */
void orc_deblock_pic(Pel* Y, int sy, Pel* U, Pel* V, int sc, int width, int height, const orc_dbk_unit* ver,
                     const orc_dbk_unit* hor, int beta_off2, int tc_off2, int bd)
{
  const int vw = (width + 7) >> 3, vh = (height + 3) >> 2, hw = (width + 3) >> 2, hh = (height + 7) >> 3;
  if (ver)
    for (int uy = 0; uy < vh; uy++)
      for (int ux = 1; ux < vw; ux++) {
        const orc_dbk_unit* u = ver + (size_t)uy * vw + ux;
        if (u->bs) filter_unit(Y, sy, U, V, sc, ux * 8, uy * 4, 0, u, beta_off2, tc_off2, bd);
      }
  if (hor)
    for (int uy = 1; uy < hh; uy++)
      for (int ux = 0; ux < hw; ux++) {
        const orc_dbk_unit* u = hor + (size_t)uy * hw + ux;
        if (u->bs) filter_unit(Y, sy, U, V, sc, ux * 4, uy * 8, 1, u, beta_off2, tc_off2, bd);
      }
}

/* ------------------------------------------------------------------ SAO apply
 * TComSampleAdaptiveOffset::processSaoUnitAll / processSaoCuOrg (TComSampleAdaptiveOffset.cpp:781-1003, 1072-1236) for one
 * colour component of a single-slice picture.  The reference works in place and reads the unfiltered left column / upper
 * row out of line buffers, the right / lower neighbours are not filtered yet: every edge class is therefore taken on the
 * deblocked picture, restated here as src -> dst.  Pinned by tests/golden/sao_golden.npz (pictures before / after the
 * reference's own SAOProcess with the per-CTU records it used). */
void orc_sao_plane(const Pel* src, Pel* dst, int stride, int w, int h, int ctu, int ctus_x, const orc_sao_unit* units, int bd)
{
  const int maxv = (1 << bd) - 1;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const orc_sao_unit* u = units + (y / ctu) * ctus_x + (x / ctu);
      const ptrdiff_t o = (ptrdiff_t)y * stride + x;
      int c = src[o], v = c;
      if (u->type == 4) v = c + u->bo[c >> (bd - 5)];                           /* SAO_BO :985-996 */
      else if (u->type >= 0) {
        /* neighbour pair of the class: EO_0 left/right :846-862, EO_1 up/down :864-887, EO_2 135 deg :889-925, EO_3 45 deg :927-962 */
        int ax, ay;
        switch (u->type) { case 0: ax = -1; ay = 0; break; case 1: ax = 0; ay = -1; break; case 2: ax = -1; ay = -1; break; default: ax = -1; ay = 1; break; }
        int okx = ax == 0 || (x > 0 && x < w - 1), oky = ay == 0 || (y > 0 && y < h - 1);
        if (okx && oky) {
          int a = src[o + ax + (ptrdiff_t)ay * stride], b = src[o - ax - (ptrdiff_t)ay * stride];
          int e = (c > a) - (c < a) + (c > b) - (c < b) + 2;
          v = c + u->eo[e];
        }
      }
      dst[o] = (Pel)clip3(0, maxv, v);                                           /* m_pClipTable :192-215 */
    }
}
