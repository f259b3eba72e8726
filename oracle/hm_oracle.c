/*
 * hm_oracle.c -- CPU restatement of the HM-7.2 hot-path arithmetic (see hm_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: the product path never calls this file.
 * Parity pin: tests/test_oracle_vs_ref.py runs every function here against the
 * reference's own compiled code (oracle/_ref/libhmref.so) and tests/golden/ (npz files)
 * holds vectors generated from that library (tests/golden/make_golden.py).
 *
 * Citations are file:line under /root/reference/source/Lib.
 */
#include "hm_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORC_MAX_UINT 0xFFFFFFFFu
#define ORC_MAX_INT  2147483647

static inline int iabs(int v) { return v < 0 ? -v : v; }
static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

/* ======================================================================= distortion */

/* TLibCommon/TComRdCost.cpp:518-989  xGetSAD4/8/16/12/16N/32/24/64/48.
 * All width-specialised variants share one arithmetic: rows are visited with step
 * 1<<subshift, the sum is scaled back by <<subshift, then >>bitIncrement. */
uint32_t orc_sad(const Pel* org, int so, const Pel* cur, int sc, int w, int h, int subshift, int bi)
{
  int step = 1 << subshift;
  uint32_t sum = 0;
  for (int rows = h; rows != 0; rows -= step) {
    for (int n = 0; n < w; n++) sum += (uint32_t)iabs(org[n] - cur[n]);
    org += so * step;
    cur += sc * step;
  }
  sum <<= subshift;
  return sum >> bi;
}

/* TLibCommon/TComRdCost.cpp:490-516  xGetSAD (generic width; ignores iSubShift). */
uint32_t orc_sad_generic(const Pel* org, int so, const Pel* cur, int sc, int w, int h, int bi)
{
  uint32_t sum = 0;
  for (int y = 0; y < h; y++) {
    for (int n = 0; n < w; n++) sum += (uint32_t)iabs(org[n] - cur[n]);
    org += so;
    cur += sc;
  }
  return sum >> bi;
}

/* TLibCommon/TComRdCost.cpp:1314-1656  xGetSSE*: each squared difference is shifted
 * by 2*bitIncrement BEFORE accumulation (:1336-1337). */
uint32_t orc_sse(const Pel* org, int so, const Pel* cur, int sc, int w, int h, int bi)
{
  uint32_t sum = 0;
  uint32_t shift = (uint32_t)bi << 1;
  for (int y = 0; y < h; y++) {
    for (int n = 0; n < w; n++) {
      int t = org[n] - cur[n];
      sum += (uint32_t)((t * t) >> shift);
    }
    org += so;
    cur += sc;
  }
  return sum;
}

/* TLibCommon/TComRdCost.cpp:1663-1682  xCalcHADs2x2 */
uint32_t orc_had2x2(const Pel* org, const Pel* cur, int so, int sc)
{
  int d0 = org[0] - cur[0], d1 = org[1] - cur[1];
  int d2 = org[so] - cur[sc], d3 = org[so + 1] - cur[sc + 1];
  int m0 = d0 + d2, m1 = d1 + d3, m2 = d0 - d2, m3 = d1 - d3;
  return (uint32_t)(iabs(m0 + m1) + iabs(m0 - m1) + iabs(m2 + m3) + iabs(m2 - m3));
}

/* 1-D unnormalised Hadamard butterfly, in place, n a power of two.  The reference's
 * hand-unrolled butterflies (TComRdCost.cpp:1700-1770, 1795-1860) compute the full 2-D
 * Hadamard transform; the SATD is the sum of |.| over all outputs and therefore does
 * not depend on the output ordering, only on the per-tile rounding below. */
static void hadamard1d(int* v, int n, int stride)
{
  for (int len = 1; len < n; len <<= 1)
    for (int i = 0; i < n; i += len << 1)
      for (int j = i; j < i + len; j++) {
        int a = v[j * stride], b = v[(j + len) * stride];
        v[j * stride] = a + b;
        v[(j + len) * stride] = a - b;
      }
}

/* TLibCommon/TComRdCost.cpp:1684-1776  xCalcHADs4x4: (sum|H4 D H4'| + 1) >> 1 */
uint32_t orc_had4x4(const Pel* org, const Pel* cur, int so, int sc)
{
  int d[16];
  for (int y = 0; y < 4; y++)
    for (int x = 0; x < 4; x++) d[y * 4 + x] = org[y * so + x] - cur[y * sc + x];
  for (int y = 0; y < 4; y++) hadamard1d(d + y * 4, 4, 1);
  for (int x = 0; x < 4; x++) hadamard1d(d + x, 4, 4);
  int satd = 0;
  for (int k = 0; k < 16; k++) satd += iabs(d[k]);
  return (uint32_t)((satd + 1) >> 1);
}

/* TLibCommon/TComRdCost.cpp:1778-1872  xCalcHADs8x8: (sum|H8 D H8'| + 2) >> 2 */
uint32_t orc_had8x8(const Pel* org, const Pel* cur, int so, int sc)
{
  int d[64];
  for (int y = 0; y < 8; y++)
    for (int x = 0; x < 8; x++) d[y * 8 + x] = org[y * so + x] - cur[y * sc + x];
  for (int y = 0; y < 8; y++) hadamard1d(d + y * 8, 8, 1);
  for (int x = 0; x < 8; x++) hadamard1d(d + x, 8, 8);
  int sad = 0;
  for (int k = 0; k < 64; k++) sad += iabs(d[k]);
  return (uint32_t)((sad + 2) >> 2);
}

/* TLibCommon/TComRdCost.cpp:2186-2287  xGetHADs (iStep == 1, NS_HAD 0):
 * 8x8 tiles when rows%8==0 && cols%8==0, else 4x4, else 2x2; total >> bitIncrement.
 * xGetHADs4 (:2122-2148) and xGetHADs8 (:2150-2184) are the same tiling for w=4 / w=8. */
uint32_t orc_hads(const Pel* org, int so, const Pel* cur, int sc, int w, int h, int bi)
{
  uint32_t sum = 0;
  if ((h % 8 == 0) && (w % 8 == 0)) {
    for (int y = 0; y < h; y += 8)
      for (int x = 0; x < w; x += 8) sum += orc_had8x8(org + y * so + x, cur + y * sc + x, so, sc);
  } else if ((h % 4 == 0) && (w % 4 == 0)) {
    for (int y = 0; y < h; y += 4)
      for (int x = 0; x < w; x += 4) sum += orc_had4x4(org + y * so + x, cur + y * sc + x, so, sc);
  } else if ((h % 2 == 0) && (w % 2 == 0)) {
    for (int y = 0; y < h; y += 2)
      for (int x = 0; x < w; x += 2) sum += orc_had2x2(org + y * so + x, cur + y * sc + x, so, sc);
  } else {
    abort(); /* :2281 assert(false) */
  }
  return sum >> bi;
}

/* TLibCommon/TComRdCost.cpp:404-447  calcHAD (intra rough mode search).  Same tiling as
 * xGetHADs; the final 2x2 branch calls the 8x8 kernel in the reference (:435-439) and is
 * unreachable for legal block sizes, so it is rejected here. */
uint32_t orc_calc_had(const Pel* p0, int s0, const Pel* p1, int s1, int w, int h, int bi)
{
  uint32_t sum = 0;
  if ((w % 8 == 0) && (h % 8 == 0)) {
    for (int y = 0; y < h; y += 8)
      for (int x = 0; x < w; x += 8) sum += orc_had8x8(p0 + y * s0 + x, p1 + y * s1 + x, s0, s1);
  } else if ((w % 4 == 0) && (h % 4 == 0)) {
    for (int y = 0; y < h; y += 4)
      for (int x = 0; x < w; x += 4) sum += orc_had4x4(p0 + y * s0 + x, p1 + y * s1 + x, s0, s1);
  } else {
    abort();
  }
  return sum >> bi;
}

/* TLibCommon/TComRdCost.cpp:286-296, 449-478  setDistParam(w,h,eDFunc) + getDistPart.
 * The function-table slot is eDFunc + g_aucConvertToBit[w] + 1; every slot of one family
 * has the same arithmetic and getDistPart always runs with iSubShift = 0, iStep = 1. */
uint32_t orc_get_dist_part(const Pel* cur, int sc, const Pel* org, int so, int w, int h, int dfunc, int bi)
{
  switch (dfunc) {
    case 1:  return orc_sse(org, so, cur, sc, w, h, bi);           /* DF_SSE  */
    case 8:  return orc_sad_generic(org, so, cur, sc, w, h, bi);   /* DF_SAD  (== xGetSADw, shift 0) */
    case 22: return orc_hads(org, so, cur, sc, w, h, bi);          /* DF_HADS */
    default: abort();
  }
}

/* TLibCommon/TComRdCost.cpp:270-284  xGetComponentBits (signed exp-Golomb length) */
uint32_t orc_mv_component_bits(int v)
{
  uint32_t len = 1;
  uint32_t t = (v <= 0) ? (uint32_t)((-v << 1) + 1) : (uint32_t)(v << 1);
  while (t != 1) { t >>= 1; len += 2; }
  return len;
}

/* TLibCommon/TComRdCost.h:203-213  getBits (FIX203 branch) */
uint32_t orc_mv_bits(int x, int y, int scale, int predx, int predy)
{
  return orc_mv_component_bits((x << scale) - predx) + orc_mv_component_bits((y << scale) - predy);
}

/* TLibCommon/TComRdCost.h:194-201  getCost(x,y) = m_uiCost * getBits(x,y) >> 16 in UInt */
uint32_t orc_mv_cost(uint32_t lambda_cost, int x, int y, int scale, int predx, int predy)
{
  return (uint32_t)(lambda_cost * orc_mv_bits(x, y, scale, predx, predy)) >> 16;
}

/* TLibCommon/TComRdCost.cpp:167-173  setLambda: m_uiLambdaMotionSAD */
uint32_t orc_lambda_motion_sad(double lambda)
{
  return (uint32_t)floor(65536.0 * sqrt(lambda));
}

/* ======================================================================= interpolation */

/* TLibCommon/TComInterpolationFilter.cpp:55-73 */
static const int16_t k_luma_filter[4][8] = {
  {  0, 0,   0, 64,  0,   0, 0,  0 },
  { -1, 4, -10, 58, 17,  -5, 1,  0 },
  { -1, 4, -11, 40, 40, -11, 4, -1 },
  {  0, 1,  -5, 17, 58, -10, 4, -1 }
};
static const int16_t k_chroma_filter[8][4] = {
  {  0, 64,  0,  0 }, { -2, 58, 10, -2 }, { -4, 54, 16, -2 }, { -6, 46, 28, -4 },
  { -4, 36, 36, -4 }, { -4, 28, 46, -6 }, { -2, 16, 54, -4 }, { -2, 10, 58, -2 }
};

#define IF_INTERNAL_PREC 14
#define IF_FILTER_PREC 6
#define IF_INTERNAL_OFFS (1 << (IF_INTERNAL_PREC - 1))

/* TLibCommon/TComInterpolationFilter.cpp:91-145  filterCopy */
void orc_filter_copy(const Pel* src, int ss, Pel* dst, int ds, int w, int h, int isFirst, int isLast, int bd)
{
  if (isFirst == isLast) {
    for (int r = 0; r < h; r++, src += ss, dst += ds)
      for (int c = 0; c < w; c++) dst[c] = src[c];
  } else if (isFirst) {
    int shift = IF_INTERNAL_PREC - bd;
    for (int r = 0; r < h; r++, src += ss, dst += ds)
      for (int c = 0; c < w; c++) {
        int16_t val = (int16_t)(src[c] << shift);
        dst[c] = (int16_t)(val - (int16_t)IF_INTERNAL_OFFS);
      }
  } else {
    int shift = IF_INTERNAL_PREC - bd;
    int16_t offset = (int16_t)IF_INTERNAL_OFFS;
    offset = (int16_t)(offset + (shift ? (1 << (shift - 1)) : 0));
    int16_t maxVal = (int16_t)((1 << bd) - 1), minVal = 0;
    for (int r = 0; r < h; r++, src += ss, dst += ds)
      for (int c = 0; c < w; c++) {
        int16_t val = src[c];
        val = (int16_t)((val + offset) >> shift);
        if (val < minVal) val = minVal;
        if (val > maxVal) val = maxVal;
        dst[c] = val;
      }
  }
}

/* TLibCommon/TComInterpolationFilter.cpp:163-244  filter<N,isVertical,isFirst,isLast>.
 * Note the (Short) truncation of (sum+offset)>>shift BEFORE the clip (:232-237). */
void orc_filter(int ntaps, int isVert, int isFirst, int isLast, const Pel* src, int ss, Pel* dst, int ds,
                int w, int h, const int16_t* coeff, int bd)
{
  int cs = isVert ? ss : 1;
  src -= (ntaps / 2 - 1) * cs;
  int headRoom = IF_INTERNAL_PREC - bd;
  int shift = IF_FILTER_PREC;
  int offset;
  int16_t maxVal;
  if (isLast) {
    shift += isFirst ? 0 : headRoom;
    offset = 1 << (shift - 1);
    offset += isFirst ? 0 : (IF_INTERNAL_OFFS << IF_FILTER_PREC);
    maxVal = (int16_t)((1 << bd) - 1);
  } else {
    shift -= isFirst ? headRoom : 0;
    offset = isFirst ? -(IF_INTERNAL_OFFS * (1 << shift)) : 0;
    maxVal = 0;
  }
  for (int r = 0; r < h; r++, src += ss, dst += ds)
    for (int c = 0; c < w; c++) {
      int sum = 0;
      for (int t = 0; t < ntaps; t++) sum += src[c + t * cs] * coeff[t];
      int16_t val = (int16_t)((sum + offset) >> shift);
      if (isLast) {
        val = (val < 0) ? 0 : val;
        val = (val > maxVal) ? maxVal : val;
      }
      dst[c] = val;
    }
}

/* TLibCommon/TComInterpolationFilter.cpp:325-337  filterHorLuma */
void orc_filter_hor_luma(const Pel* src, int ss, Pel* dst, int ds, int w, int h, int frac, int isLast, int bd)
{
  if (frac == 0) orc_filter_copy(src, ss, dst, ds, w, h, 1, isLast, bd);
  else orc_filter(8, 0, 1, isLast, src, ss, dst, ds, w, h, k_luma_filter[frac], bd);
}
/* TLibCommon/TComInterpolationFilter.cpp:352-364  filterVerLuma */
void orc_filter_ver_luma(const Pel* src, int ss, Pel* dst, int ds, int w, int h, int frac, int isFirst, int isLast, int bd)
{
  if (frac == 0) orc_filter_copy(src, ss, dst, ds, w, h, isFirst, isLast, bd);
  else orc_filter(8, 1, isFirst, isLast, src, ss, dst, ds, w, h, k_luma_filter[frac], bd);
}
/* TLibCommon/TComInterpolationFilter.cpp:378-390  filterHorChroma */
void orc_filter_hor_chroma(const Pel* src, int ss, Pel* dst, int ds, int w, int h, int frac, int isLast, int bd)
{
  if (frac == 0) orc_filter_copy(src, ss, dst, ds, w, h, 1, isLast, bd);
  else orc_filter(4, 0, 1, isLast, src, ss, dst, ds, w, h, k_chroma_filter[frac], bd);
}
/* TLibCommon/TComInterpolationFilter.cpp:405-417  filterVerChroma */
void orc_filter_ver_chroma(const Pel* src, int ss, Pel* dst, int ds, int w, int h, int frac, int isFirst, int isLast, int bd)
{
  if (frac == 0) orc_filter_copy(src, ss, dst, ds, w, h, isFirst, isLast, bd);
  else orc_filter(4, 1, isFirst, isLast, src, ss, dst, ds, w, h, k_chroma_filter[frac], bd);
}

/* ======================================================================= motion compensation */

#define ORC_TMP_STRIDE 80            /* TComPrediction.cpp:85  g_uiMaxCUWidth + 16 */
#define ORC_TMP_ROWS   (64 + 1 + 7 + 8)

/* TLibCommon/TComPrediction.cpp:554-586  xPredInterLumaBlk.  `ref` points at the PU's
 * co-located pel in the padded reference plane; mv in quarter pels (already clipped). */
void orc_pred_inter_luma_blk(const Pel* ref, int rs, int mvx, int mvy, int w, int h, Pel* dst, int ds, int bi_flag, int bd)
{
  ref += (mvx >> 2) + (mvy >> 2) * rs;
  int xFrac = mvx & 3, yFrac = mvy & 3;
  if (yFrac == 0) {
    orc_filter_hor_luma(ref, rs, dst, ds, w, h, xFrac, !bi_flag, bd);
  } else if (xFrac == 0) {
    orc_filter_ver_luma(ref, rs, dst, ds, w, h, yFrac, 1, !bi_flag, bd);
  } else {
    Pel tmp[ORC_TMP_STRIDE * ORC_TMP_ROWS];
    orc_filter_hor_luma(ref - 3 * rs, rs, tmp, ORC_TMP_STRIDE, w, h + 7, xFrac, 0, bd);
    orc_filter_ver_luma(tmp + 3 * ORC_TMP_STRIDE, ORC_TMP_STRIDE, dst, ds, w, h, yFrac, 0, !bi_flag, bd);
  }
}

/* TLibCommon/TComPrediction.cpp:600-645  xPredInterChromaBlk for ONE chroma plane.
 * w,h are the LUMA PU size; mv is the luma quarter-pel MV (chroma units 1/8). */
void orc_pred_inter_chroma_blk(const Pel* ref, int rs, int mvx, int mvy, int w, int h, Pel* dst, int ds, int bi_flag, int bd)
{
  ref += (mvx >> 3) + (mvy >> 3) * rs;
  int xFrac = mvx & 7, yFrac = mvy & 7;
  int cw = w >> 1, ch = h >> 1;
  if (yFrac == 0) {
    orc_filter_hor_chroma(ref, rs, dst, ds, cw, ch, xFrac, !bi_flag, bd);
  } else if (xFrac == 0) {
    orc_filter_ver_chroma(ref, rs, dst, ds, cw, ch, yFrac, 1, !bi_flag, bd);
  } else {
    Pel tmp[ORC_TMP_STRIDE * ORC_TMP_ROWS];
    orc_filter_hor_chroma(ref - rs, rs, tmp, ORC_TMP_STRIDE, cw, ch + 3, xFrac, 0, bd);
    orc_filter_ver_chroma(tmp + ORC_TMP_STRIDE, ORC_TMP_STRIDE, dst, ds, cw, ch, yFrac, 0, !bi_flag, bd);
  }
}

/* TLibCommon/TComYuv.cpp:520-581  addAvg (one plane) */
void orc_add_avg(const Pel* s0, int st0, const Pel* s1, int st1, Pel* dst, int ds, int w, int h, int bd)
{
  int shiftNum = IF_INTERNAL_PREC + 1 - bd;
  int offset = (1 << (shiftNum - 1)) + 2 * IF_INTERNAL_OFFS;
  int maxv = (1 << bd) - 1;
  for (int y = 0; y < h; y++, s0 += st0, s1 += st1, dst += ds)
    for (int x = 0; x < w; x++) dst[x] = (Pel)clip3(0, maxv, (s0[x] + s1[x] + offset) >> shiftNum);
}

/* TLibCommon/TComYuv.cpp:462-485  subtractLuma / subtractChroma */
void orc_subtract(const Pel* s0, int st0, const Pel* s1, int st1, Pel* dst, int ds, int w, int h)
{
  for (int y = 0; y < h; y++, s0 += st0, s1 += st1, dst += ds)
    for (int x = 0; x < w; x++) dst[x] = (Pel)(s0[x] - s1[x]);
}

/* TLibCommon/TComYuv.cpp:407-429  addClipLuma / addClipChroma: Clip(a+b) to [0, 2^bd-1] */
void orc_add_clip(const Pel* s0, int st0, const Pel* s1, int st1, Pel* dst, int ds, int w, int h, int bd)
{
  int maxv = (1 << bd) - 1;
  for (int y = 0; y < h; y++, s0 += st0, s1 += st1, dst += ds)
    for (int x = 0; x < w; x++) dst[x] = (Pel)clip3(0, maxv, s0[x] + s1[x]);
}

/* TLibCommon/TComYuv.cpp:583-633  removeHighFreq (DISABLING_CLIP_FOR_BIPREDME: no clip) */
void orc_remove_high_freq(Pel* dst, int ds, const Pel* src, int ss, int w, int h)
{
  for (int y = 0; y < h; y++, src += ss, dst += ds)
    for (int x = 0; x < w; x++) dst[x] = (Pel)((dst[x] << 1) - src[x]);
}

/* TLibCommon/TComPicYuv.cpp:259-286  xExtendPicCompBorder; `pic` points at pel (0,0) */
void orc_extend_border(Pel* pic, int stride, int w, int h, int mx, int my)
{
  Pel* pi = pic;
  for (int y = 0; y < h; y++, pi += stride)
    for (int x = 0; x < mx; x++) {
      pi[-mx + x] = pi[0];
      pi[w + x] = pi[w - 1];
    }
  pi -= (stride + mx);
  for (int y = 0; y < my; y++) memcpy(pi + (y + 1) * stride, pi, sizeof(Pel) * (size_t)(w + (mx << 1)));
  pi -= (h - 1) * stride;
  for (int y = 0; y < my; y++) memcpy(pi - (y + 1) * stride, pi, sizeof(Pel) * (size_t)(w + (mx << 1)));
}
