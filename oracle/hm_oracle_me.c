/*
 * hm_oracle_me.c -- CPU restatement of HM-7.2 integer and fractional motion search.
 * TEST INFRASTRUCTURE ONLY (see hm_oracle.h).  Citations: /root/reference/source/Lib.
 */
#include "hm_oracle.h"
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_UINT 0xFFFFFFFFu

/* TLibCommon/TComDataCU.cpp:3505-3517  clipMv (quarter-pel units) */
void orc_clip_mv(const orc_cu_geom* g, int* mvx, int* mvy)
{
  const int sh = 2, off = 8;
  int hmax = (g->pic_w + off - g->cu_x - 1) << sh;
  int hmin = (-g->max_cu - off - g->cu_x + 1) * (1 << sh);
  int vmax = (g->pic_h + off - g->cu_y - 1) << sh;
  int vmin = (-g->max_cu - off - g->cu_y + 1) * (1 << sh);
  int x = *mvx, y = *mvy;
  x = x > hmin ? x : hmin; x = x < hmax ? x : hmax;
  y = y > vmin ? y : vmin; y = y < vmax ? y : vmax;
  *mvx = x; *mvy = y;
}

/* TLibEncoder/TEncSearch.cpp:4209-4225  xSetSearchRange: window = clip(pred) +- SR, clipped,
 * returned in integer pels (arithmetic >> 2). */
void orc_set_search_range(const orc_cu_geom* g, int predx, int predy, int srange,
                          int* lx, int* ty, int* rx, int* by)
{
  int px = predx, py = predy;
  orc_clip_mv(g, &px, &py);
  int l = px - (srange << 2), t = py - (srange << 2);
  int r = px + (srange << 2), b = py + (srange << 2);
  orc_clip_mv(g, &l, &t);
  orc_clip_mv(g, &r, &b);
  *lx = l >> 2; *ty = t >> 2; *rx = r >> 2; *by = b >> 2;
}

/* search state: TLibEncoder/TEncSearch.h IntTZSearchStruct */
typedef struct {
  const Pel* org; int so;
  const Pel* ref; int rs;
  int w, h, subshift, bi;
  uint32_t lambda_cost; int predx, predy;
  uint32_t best_sad; int best_x, best_y;
  uint32_t best_dist, best_round; int point_nr;
  uint32_t n_sads;
  int lx, ty, rx, by;
} tz_state;

/* TLibEncoder/TEncSearch.cpp:312-349  xTZSearchHelp: SAD (rows subsampled when FEN && rows>8)
 * + MV rate at cost scale 2; strict '<' keeps the first visited minimum. */
static void tz_help(tz_state* s, int x, int y, int point_nr, uint32_t dist)
{
  uint32_t sad = orc_sad(s->org, s->so, s->ref + y * s->rs + x, s->rs, s->w, s->h, s->subshift, s->bi);
  sad += orc_mv_cost(s->lambda_cost, x, y, 2, s->predx, s->predy);
  s->n_sads++;
  if (sad < s->best_sad) {
    s->best_sad = sad; s->best_x = x; s->best_y = y;
    s->best_dist = dist; s->best_round = 0; s->point_nr = point_nr;
  }
}

/* TLibEncoder/TEncSearch.cpp:351-476  xTZ2PointSearch: the two untested neighbours of the
 * current best given which of the 8 compass points (1..8, row-major around 0) it was. */
static void tz_two_point(tz_state* s)
{
  int x = s->best_x, y = s->best_y;
  int up = (y - 1) >= s->ty, dn = (y + 1) <= s->by, lf = (x - 1) >= s->lx, rt = (x + 1) <= s->rx;
  switch (s->point_nr) {
    case 1: if (lf) tz_help(s, x - 1, y, 0, 2); if (up) tz_help(s, x, y - 1, 0, 2); break;
    case 2: if (up) { if (lf) tz_help(s, x - 1, y - 1, 0, 2); if (rt) tz_help(s, x + 1, y - 1, 0, 2); } break;
    case 3: if (up) tz_help(s, x, y - 1, 0, 2); if (rt) tz_help(s, x + 1, y, 0, 2); break;
    case 4: if (lf) { if (dn) tz_help(s, x - 1, y + 1, 0, 2); if (up) tz_help(s, x - 1, y - 1, 0, 2); } break;
    case 5: if (rt) { if (up) tz_help(s, x + 1, y - 1, 0, 2); if (dn) tz_help(s, x + 1, y + 1, 0, 2); } break;
    case 6: if (lf) tz_help(s, x - 1, y, 0, 2); if (dn) tz_help(s, x, y + 1, 0, 2); break;
    case 7: if (dn) { if (lf) tz_help(s, x - 1, y + 1, 0, 2); if (rt) tz_help(s, x + 1, y + 1, 0, 2); } break;
    case 8: if (rt) tz_help(s, x + 1, y, 0, 2); if (dn) tz_help(s, x, y + 1, 0, 2); break;
    default: abort();
  }
}

/* TLibEncoder/TEncSearch.cpp:535-707  xTZ8PointDiamondSearch */
static void tz_diamond(tz_state* s, int sx, int sy, int d)
{
  int top = sy - d, bot = sy + d, lef = sx - d, rig = sx + d;
  s->best_round += 1;
  if (d == 1) {
    if (top >= s->ty) tz_help(s, sx, top, 2, d);
    if (lef >= s->lx) tz_help(s, lef, sy, 4, d);
    if (rig <= s->rx) tz_help(s, rig, sy, 5, d);
    if (bot <= s->by) tz_help(s, sx, bot, 7, d);
    return;
  }
  int inside = top >= s->ty && lef >= s->lx && rig <= s->rx && bot <= s->by;
  if (d <= 8) {
    int h2 = d >> 1;
    int top2 = sy - h2, bot2 = sy + h2, lef2 = sx - h2, rig2 = sx + h2;
    if (inside) {
      tz_help(s, sx, top, 2, d);
      tz_help(s, lef2, top2, 1, h2);
      tz_help(s, rig2, top2, 3, h2);
      tz_help(s, lef, sy, 4, d);
      tz_help(s, rig, sy, 5, d);
      tz_help(s, lef2, bot2, 6, h2);
      tz_help(s, rig2, bot2, 8, h2);
      tz_help(s, sx, bot, 7, d);
    } else {
      if (top >= s->ty) tz_help(s, sx, top, 2, d);
      if (top2 >= s->ty) {
        if (lef2 >= s->lx) tz_help(s, lef2, top2, 1, h2);
        if (rig2 <= s->rx) tz_help(s, rig2, top2, 3, h2);
      }
      if (lef >= s->lx) tz_help(s, lef, sy, 4, d);
      if (rig <= s->rx) tz_help(s, rig, sy, 5, d);
      if (bot2 <= s->by) {
        if (lef2 >= s->lx) tz_help(s, lef2, bot2, 6, h2);
        if (rig2 <= s->rx) tz_help(s, rig2, bot2, 8, h2);
      }
      if (bot <= s->by) tz_help(s, sx, bot, 7, d);
    }
    return;
  }
  /* d > 8: 16 points, all tagged point 0 */
  int q = d >> 2;
  if (inside) {
    tz_help(s, sx, top, 0, d);
    tz_help(s, lef, sy, 0, d);
    tz_help(s, rig, sy, 0, d);
    tz_help(s, sx, bot, 0, d);
    for (int i = 1; i < 4; i++) {
      int yt = top + q * i, yb = bot - q * i, xl = sx - q * i, xr = sx + q * i;
      tz_help(s, xl, yt, 0, d);
      tz_help(s, xr, yt, 0, d);
      tz_help(s, xl, yb, 0, d);
      tz_help(s, xr, yb, 0, d);
    }
  } else {
    if (top >= s->ty) tz_help(s, sx, top, 0, d);
    if (lef >= s->lx) tz_help(s, lef, sy, 0, d);
    if (rig <= s->rx) tz_help(s, rig, sy, 0, d);
    if (bot <= s->by) tz_help(s, sx, bot, 0, d);
    for (int i = 1; i < 4; i++) {
      int yt = top + q * i, yb = bot - q * i, xl = sx - q * i, xr = sx + q * i;
      if (yt >= s->ty) {
        if (xl >= s->lx) tz_help(s, xl, yt, 0, d);
        if (xr <= s->rx) tz_help(s, xr, yt, 0, d);
      }
      if (yb <= s->by) {
        if (xl >= s->lx) tz_help(s, xl, yb, 0, d);
        if (xr <= s->rx) tz_help(s, xr, yb, 0, d);
      }
    }
  }
}

/* TLibEncoder/TEncSearch.cpp:4302-4474  xTZSearch with TZ_SEARCH_CONFIGURATION (:293-309):
 * iRaster 5, zero-vector test, diamond first search stopping 3 rounds after the best
 * (FASTME_SMOOTHER_MV, CommonDef.h:174), raster when best distance > 5, star refinement
 * (diamond, no early stop).  (startx_q,starty_q) is rcMv on entry = the AMVP predictor in
 * quarter pels; it is clipped and >>2 (:4311-4312).  (predx,predy) is the rate predictor. */
void orc_tz_search(const orc_cu_geom* g, const Pel* org, int so, const Pel* ref, int rs, int w, int h,
                   int lx, int ty, int rx, int by, int srange, int fen, int bi,
                   uint32_t lambda_cost, int predx, int predy, int startx_q, int starty_q,
                   orc_me_result* out)
{
  const int raster = 5;
  tz_state s;
  memset(&s, 0, sizeof(s));
  s.org = org; s.so = so; s.ref = ref; s.rs = rs; s.w = w; s.h = h; s.bi = bi;
  s.subshift = (fen && h > 8) ? 1 : 0;                 /* :324-330 */
  s.lambda_cost = lambda_cost; s.predx = predx; s.predy = predy;
  s.lx = lx; s.ty = ty; s.rx = rx; s.by = by;
  s.best_sad = ORC_MAX_UINT;

  int mx = startx_q, my = starty_q;
  orc_clip_mv(g, &mx, &my);
  mx >>= 2; my >>= 2;
  tz_help(&s, mx, my, 0, 0);                           /* :4320 predictor as start     */
  tz_help(&s, 0, 0, 0, 0);                             /* :4336-4339 zero vector       */

  int sx = s.best_x, sy = s.best_y;
  for (int d = 1; d <= srange; d *= 2) {               /* :4346-4361 first search      */
    tz_diamond(&s, sx, sy, d);
    if (s.best_round >= 3) break;
  }
  if (s.best_dist == 1) {                              /* :4382-4386                   */
    s.best_dist = 0;
    tz_two_point(&s);
  }
  if ((int)s.best_dist > raster) {                     /* :4389-4400 raster            */
    s.best_dist = raster;
    for (int y = ty; y <= by; y += raster)
      for (int x = lx; x <= rx; x += raster) tz_help(&s, x, y, 0, raster);
  }
  while (s.best_dist > 0) {                            /* :4435-4468 star refinement   */
    sx = s.best_x; sy = s.best_y;
    s.best_dist = 0; s.point_nr = 0;
    for (int d = 1; d < srange + 1; d *= 2) tz_diamond(&s, sx, sy, d);
    if (s.best_dist == 1) {
      s.best_dist = 0;
      if (s.point_nr != 0) tz_two_point(&s);
    }
  }
  out->mvx = s.best_x; out->mvy = s.best_y;
  out->sad = s.best_sad - orc_mv_cost(lambda_cost, s.best_x, s.best_y, 2, predx, predy);
  out->n_sads = s.n_sads;
}

/* TLibEncoder/TEncSearch.cpp:4227-4283  xPatternSearch: exhaustive raster over the window,
 * y outer / x inner, strict '<'. */
void orc_pattern_search(const Pel* org, int so, const Pel* ref, int rs, int w, int h,
                        int lx, int ty, int rx, int by, int fen, int bi,
                        uint32_t lambda_cost, int predx, int predy, orc_me_result* out)
{
  int subshift = (fen && h > 8) ? 1 : 0;               /* :4245-4251 */
  uint32_t best = ORC_MAX_UINT; int bx = 0, byy = 0; uint32_t n = 0;
  for (int y = ty; y <= by; y++)
    for (int x = lx; x <= rx; x++) {
      uint32_t sad = orc_sad(org, so, ref + y * rs + x, rs, w, h, subshift, bi);
      sad += orc_mv_cost(lambda_cost, x, y, 2, predx, predy);
      n++;
      if (sad < best) { best = sad; bx = x; byy = y; }
    }
  out->mvx = bx; out->mvy = byy;
  out->sad = best - orc_mv_cost(lambda_cost, bx, byy, 2, predx, predy);
  out->n_sads = n;
}

/* ------------------------------------------------------------------ fractional search */

#define FB_STRIDE 80                  /* TComPrediction.cpp:85  extWidth  */
#define FB_ROWS   (64 + 1)            /* :86  extHeight                    */
#define FBT_ROWS  (64 + 1 + 7)        /* :90                               */

typedef struct {
  Pel tmp[4][FB_STRIDE * FBT_ROWS];   /* m_filteredBlockTmp[4]            */
  Pel blk[4][4][FB_STRIDE * FB_ROWS]; /* m_filteredBlock[yFrac][xFrac]    */
} frac_bufs;

/* TLibEncoder/TEncSearch.cpp:5982-6014  xExtDIFUpSamplingH; roi = reference at integer MV */
static void ext_dif_up_h(frac_bufs* f, const Pel* roi, int rs, int w, int h, int bd)
{
  const int fs = 8, hf = 4;
  const Pel* src = roi - hf * rs - 1;
  orc_filter_hor_luma(src, rs, f->tmp[0], FB_STRIDE, w + 1, h + fs, 0, 0, bd);
  orc_filter_hor_luma(src, rs, f->tmp[2], FB_STRIDE, w + 1, h + fs, 2, 0, bd);
  orc_filter_ver_luma(f->tmp[0] + hf * FB_STRIDE + 1,       FB_STRIDE, f->blk[0][0], FB_STRIDE, w + 0, h + 0, 0, 0, 1, bd);
  orc_filter_ver_luma(f->tmp[0] + (hf - 1) * FB_STRIDE + 1, FB_STRIDE, f->blk[2][0], FB_STRIDE, w + 0, h + 1, 2, 0, 1, bd);
  orc_filter_ver_luma(f->tmp[2] + hf * FB_STRIDE,           FB_STRIDE, f->blk[0][2], FB_STRIDE, w + 1, h + 0, 0, 0, 1, bd);
  orc_filter_ver_luma(f->tmp[2] + (hf - 1) * FB_STRIDE,     FB_STRIDE, f->blk[2][2], FB_STRIDE, w + 1, h + 1, 2, 0, 1, bd);
}

/* TLibEncoder/TEncSearch.cpp:6023-6175  xExtDIFUpSamplingQ; (hx,hy) = chosen half-pel offset */
static void ext_dif_up_q(frac_bufs* f, const Pel* roi, int rs, int w, int h, int hx, int hy, int bd)
{
  const int fs = 8, hf = 4, S = FB_STRIDE;
  int extH = (hy == 0) ? h + fs : h + fs - 1;
  const Pel* src = roi - hf * rs - 1;
  if (hy > 0) src += rs;
  if (hx >= 0) src += 1;
  orc_filter_hor_luma(src, rs, f->tmp[1], S, w, extH, 1, 0, bd);
  src = roi - hf * rs - 1;
  if (hy > 0) src += rs;
  if (hx > 0) src += 1;
  orc_filter_hor_luma(src, rs, f->tmp[3], S, w, extH, 3, 0, bd);

  const Pel* ip;
  ip = f->tmp[1] + (hf - 1) * S; if (hy == 0) ip += S;                       /* @1,1 */
  orc_filter_ver_luma(ip, S, f->blk[1][1], S, w, h, 1, 0, 1, bd);
  ip = f->tmp[1] + (hf - 1) * S;                                             /* @3,1 */
  orc_filter_ver_luma(ip, S, f->blk[3][1], S, w, h, 3, 0, 1, bd);
  if (hy != 0) {
    ip = f->tmp[1] + (hf - 1) * S;                                           /* @2,1 */
    orc_filter_ver_luma(ip, S, f->blk[2][1], S, w, h, 2, 0, 1, bd);
    ip = f->tmp[3] + (hf - 1) * S;                                           /* @2,3 */
    orc_filter_ver_luma(ip, S, f->blk[2][3], S, w, h, 2, 0, 1, bd);
  } else {
    ip = f->tmp[1] + hf * S;                                                 /* @0,1 */
    orc_filter_ver_luma(ip, S, f->blk[0][1], S, w, h, 0, 0, 1, bd);
    ip = f->tmp[3] + hf * S;                                                 /* @0,3 */
    orc_filter_ver_luma(ip, S, f->blk[0][3], S, w, h, 0, 0, 1, bd);
  }
  if (hx != 0) {
    ip = f->tmp[2] + (hf - 1) * S; if (hx > 0) ip += 1; if (hy >= 0) ip += S; /* @1,2 */
    orc_filter_ver_luma(ip, S, f->blk[1][2], S, w, h, 1, 0, 1, bd);
    ip = f->tmp[2] + (hf - 1) * S; if (hx > 0) ip += 1; if (hy > 0) ip += S;  /* @3,2 */
    orc_filter_ver_luma(ip, S, f->blk[3][2], S, w, h, 3, 0, 1, bd);
  } else {
    ip = f->tmp[0] + (hf - 1) * S + 1; if (hy >= 0) ip += S;                  /* @1,0 */
    orc_filter_ver_luma(ip, S, f->blk[1][0], S, w, h, 1, 0, 1, bd);
    ip = f->tmp[0] + (hf - 1) * S + 1; if (hy > 0) ip += S;                   /* @3,0 */
    orc_filter_ver_luma(ip, S, f->blk[3][0], S, w, h, 3, 0, 1, bd);
  }
  ip = f->tmp[3] + (hf - 1) * S; if (hy == 0) ip += S;                        /* @1,3 */
  orc_filter_ver_luma(ip, S, f->blk[1][3], S, w, h, 1, 0, 1, bd);
  ip = f->tmp[3] + (hf - 1) * S;                                              /* @3,3 */
  orc_filter_ver_luma(ip, S, f->blk[3][3], S, w, h, 3, 0, 1, bd);
}

/* TLibEncoder/TEncSearch.cpp:47-71 */
static const int k_refine_h[9][2] = { {0,0},{0,-1},{0,1},{-1,0},{1,0},{-1,-1},{1,-1},{-1,1},{1,1} };
static const int k_refine_q[9][2] = { {0,0},{0,-1},{0,1},{-1,-1},{1,-1},{-1,0},{1,0},{-1,1},{1,1} };

/* TLibEncoder/TEncSearch.cpp:711-760  xPatternRefinement.  (mvfx,mvfy) is rcMvFrac on entry
 * (the absolute MV at this precision, for the rate term) and the chosen offset on exit. */
static uint32_t pattern_refinement(const frac_bufs* f, const Pel* org, int so, int w, int h,
                                   int basex, int basey, int frac, int* mvfx, int* mvfy,
                                   int hadamard, int bi, uint32_t lambda_cost, int scale, int predx, int predy)
{
  const int (*ref)[2] = (frac == 2) ? k_refine_h : k_refine_q;
  uint32_t best = ORC_MAX_UINT; int best_i = 0;
  for (int i = 0; i < 9; i++) {
    int hor = (ref[i][0] + basex) * frac, ver = (ref[i][1] + basey) * frac;
    const Pel* p = f->blk[ver & 3][hor & 3];
    if (hor == 2 && (ver & 1) == 0) p += 1;
    if ((hor & 1) == 0 && ver == 2) p += FB_STRIDE;
    int tx = ref[i][0] + *mvfx, ty = ref[i][1] + *mvfy;
    uint32_t d = hadamard ? orc_hads(org, so, p, FB_STRIDE, w, h, bi)
                          : orc_sad(org, so, p, FB_STRIDE, w, h, 0, bi);
    d += orc_mv_cost(lambda_cost, tx, ty, scale, predx, predy);
    if (d < best) { best = d; best_i = i; }
  }
  *mvfx = ref[best_i][0]; *mvfy = ref[best_i][1];
  return best;
}

/* TLibEncoder/TEncSearch.cpp:4476-4514  xPatternSearchFracDIF.  Cost scale is 1 for the
 * half-pel pass (set by the caller, :4187) and 0 for the quarter-pel pass (:4505). */
void orc_frac_search(const Pel* org, int so, const Pel* ref, int rs, int w, int h,
                     int imvx, int imvy, int hadamard, int bi, int bd,
                     uint32_t lambda_cost, int predx, int predy, orc_frac_result* out)
{
  frac_bufs* f = (frac_bufs*)malloc(sizeof(frac_bufs));
  memset(f, 0, sizeof(*f));
  const Pel* roi = ref + imvx + imvy * rs;
  ext_dif_up_h(f, roi, rs, w, h, bd);
  int hx = imvx << 1, hy = imvy << 1;
  out->cost_half = pattern_refinement(f, org, so, w, h, 0, 0, 2, &hx, &hy, hadamard, bi, lambda_cost, 1, predx, predy);
  ext_dif_up_q(f, roi, rs, w, h, hx, hy, bd);
  int qx = ((imvx << 1) + hx) << 1, qy = ((imvy << 1) + hy) << 1;
  out->cost = pattern_refinement(f, org, so, w, h, hx << 1, hy << 1, 1, &qx, &qy, hadamard, bi, lambda_cost, 0, predx, predy);
  out->halfx = hx; out->halfy = hy; out->qtrx = qx; out->qtry = qy;
  free(f);
}
