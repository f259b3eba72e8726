/*
 * hm_oracle_intra.c -- CPU restatement of the intra 35-mode rough search.
 *
 * TEST INFRASTRUCTURE ONLY (see hm_oracle.h).  Follows, in the reference's own loop order:
 *   TComPattern::initAdiPattern        TLibCommon/TComPattern.cpp:262-307   ([1 2 1] smoothing of the reference line)
 *   TComPattern::getPredictorPtr       TLibCommon/TComPattern.cpp:577-605   (which modes read the smoothed line)
 *   TComPrediction::predIntraLumaAng   TLibCommon/TComPrediction.cpp:337-366
 *   TComPrediction::xPredIntraAng      TLibCommon/TComPrediction.cpp:186-335
 *   TComPrediction::predIntraGetPredValDC  :127-165,  xPredIntraPlanar :689-731,  xDCPredFiltering :1010-1031
 *   TEncSearch::estIntraPredQT         TLibEncoder/TEncSearch.cpp:2530-2537 (predIntraLumaAng + calcHAD per mode)
 *
 * Reference samples are exchanged as ONE LINE of 4N+1 Pels in the order initAdiPattern itself walks them
 * (TComPattern.cpp:277-288): the left column from bottom-left (2N samples, bottom to top), the top-left corner,
 * then the row above from left to above-right (2N samples).  The reference keeps the same samples as column 0 /
 * row 0 of a (2N+1)x(2N+1) Int array; this file rebuilds that array so that the prediction code below reads its
 * neighbours exactly like the reference does (pSrc[k - srcStride], pSrc[k*srcStride - 1]).
 */
#include <stdlib.h>
#include <string.h>
#include "hm_oracle.h"

#define ORC_MAX_N 64

/* TComPattern.cpp:262-307: smoothed copy of the line; the two end samples are copied */
void orc_intra_filter_line(const Pel* line, int n, Pel* out)
{
  const int len = 4 * n + 1;
  out[0] = line[0];
  out[len - 1] = line[len - 1];
  for (int i = 1; i < len - 1; i++) out[i] = (Pel)((line[i - 1] + 2 * line[i] + line[i + 1] + 2) >> 2);
}

/* line -> the reference's 2-D layout (only column 0 and row 0 are ever read by the predictors) */
static void line_to_adi(const Pel* line, int n, int* adi)
{
  const int sw = 2 * n + 1;
  memset(adi, 0, sizeof(int) * (size_t)sw * sw);
  int l = 0;
  for (int i = 0; i < 2 * n; i++) adi[sw * (2 * n - i)] = line[l++];
  adi[0] = line[l++];
  for (int i = 0; i < 2 * n; i++) adi[1 + i] = line[l++];
}

static int clip_pel(int v, int bd) { const int mx = (1 << bd) - 1; return v < 0 ? 0 : (v > mx ? mx : v); }

/* predIntraGetPredValDC, TComPrediction.cpp:127-165 */
static Pel dc_val(const int* src, int ss, int w, int h, int above, int left)
{
  int sum = 0;
  if (above) for (int i = 0; i < w; i++) sum += src[i - ss];
  if (left) for (int i = 0; i < h; i++) sum += src[i * ss - 1];
  if (above && left) return (Pel)((sum + w) / (w + h));
  if (above) return (Pel)((sum + w / 2) / w);
  if (left) return (Pel)((sum + h / 2) / h);
  return (Pel)src[-1];
}

/* xPredIntraPlanar, TComPrediction.cpp:689-731 */
static void pred_planar(const int* src, int ss, Pel* dst, int ds, int n, int log2n)
{
  int left[ORC_MAX_N + 1], top[ORC_MAX_N + 1], bottom[ORC_MAX_N], right[ORC_MAX_N];
  const int shift1 = log2n, shift2 = log2n + 1;
  for (int k = 0; k < n + 1; k++) { top[k] = src[k - ss]; left[k] = src[k * ss - 1]; }
  const int bl = left[n], tr = top[n];
  for (int k = 0; k < n; k++) {
    bottom[k] = bl - top[k];
    right[k] = tr - left[k];
    top[k] <<= shift1;
    left[k] <<= shift1;
  }
  for (int k = 0; k < n; k++) {
    int hor = left[k] + n;
    for (int l = 0; l < n; l++) {
      hor += right[k];
      top[l] += bottom[l];
      dst[k * ds + l] = (Pel)((hor + top[l]) >> shift2);
    }
  }
}

/* xPredIntraAng, TComPrediction.cpp:186-335 (luma: bFilter = true) */
static void pred_ang(const int* src, int ss, Pel* dst, int ds, int n, int mode, int above, int left, int filter, int bd)
{
  static const int ang_table[9] = {0, 2, 5, 9, 13, 17, 21, 26, 32};
  static const int inv_table[9] = {0, 4096, 1638, 910, 630, 482, 390, 315, 256};
  const int mode_dc = mode < 2, mode_hor = !mode_dc && mode < 18, mode_ver = !mode_dc && !mode_hor;
  int ang = mode_ver ? mode - 26 : (mode_hor ? -(mode - 10) : 0);
  int abs_ang = abs(ang);
  const int sign = ang < 0 ? -1 : 1;
  const int inv = inv_table[abs_ang];
  abs_ang = ang_table[abs_ang];
  ang = sign * abs_ang;
  if (mode_dc) {
    const Pel dc = dc_val(src, ss, n, n, above, left);
    for (int k = 0; k < n; k++) for (int l = 0; l < n; l++) dst[k * ds + l] = dc;
    return;
  }
  Pel ref_above[2 * ORC_MAX_N + 1], ref_left[2 * ORC_MAX_N + 1];
  Pel *ref_main, *ref_side;
  if (ang < 0) {
    for (int k = 0; k < n + 1; k++) ref_above[k + n - 1] = (Pel)src[k - ss - 1];
    for (int k = 0; k < n + 1; k++) ref_left[k + n - 1] = (Pel)src[(k - 1) * ss - 1];
    ref_main = (mode_ver ? ref_above : ref_left) + (n - 1);
    ref_side = (mode_ver ? ref_left : ref_above) + (n - 1);
    int inv_sum = 128;
    for (int k = -1; k > (n * ang) >> 5; k--) {
      inv_sum += inv;
      ref_main[k] = ref_side[inv_sum >> 8];
    }
  } else {
    for (int k = 0; k < 2 * n + 1; k++) ref_above[k] = (Pel)src[k - ss - 1];
    for (int k = 0; k < 2 * n + 1; k++) ref_left[k] = (Pel)src[(k - 1) * ss - 1];
    ref_main = mode_ver ? ref_above : ref_left;
    ref_side = mode_ver ? ref_left : ref_above;
  }
  if (ang == 0) {
    for (int k = 0; k < n; k++) for (int l = 0; l < n; l++) dst[k * ds + l] = ref_main[l + 1];
    if (filter)
      for (int k = 0; k < n; k++) dst[k * ds] = (Pel)clip_pel(dst[k * ds] + ((ref_side[k + 1] - ref_side[0]) >> 1), bd);
  } else {
    int pos = 0;
    for (int k = 0; k < n; k++) {
      pos += ang;
      const int di = pos >> 5, df = pos & 31;
      if (df) {
        for (int l = 0; l < n; l++) {
          const int idx = l + di + 1;
          dst[k * ds + l] = (Pel)(((32 - df) * ref_main[idx] + df * ref_main[idx + 1] + 16) >> 5);
        }
      } else {
        for (int l = 0; l < n; l++) dst[k * ds + l] = ref_main[l + di + 1];
      }
    }
  }
  if (mode_hor)
    for (int k = 0; k < n - 1; k++)
      for (int l = k + 1; l < n; l++) {
        const Pel t = dst[k * ds + l];
        dst[k * ds + l] = dst[l * ds + k];
        dst[l * ds + k] = t;
      }
}

/* xDCPredFiltering, TComPrediction.cpp:1010-1031 */
static void dc_filtering(const int* src, int ss, Pel* dst, int ds, int n)
{
  dst[0] = (Pel)((src[-ss] + src[-1] + 2 * dst[0] + 2) >> 2);
  for (int x = 1; x < n; x++) dst[x] = (Pel)((src[x - ss] + 3 * dst[x] + 2) >> 2);
  for (int y = 1; y < n; y++) dst[y * ds] = (Pel)((src[y * ss - 1] + 3 * dst[y * ds] + 2) >> 2);
}

/* does `mode` read the smoothed line?  getPredictorPtr, TComPattern.cpp:577-605 with m_aucIntraFilter :49-56 */
int orc_intra_mode_filtered(int mode, int log2n)
{
  static const int thr[5] = {10, 7, 1, 0, 10};
  if (mode == 1) return 0;
  const int dh = abs(mode - 10), dv = abs(mode - 26);
  return (dh < dv ? dh : dv) > thr[log2n - 2];
}

/* predIntraLumaAng, TComPrediction.cpp:337-366: one mode of one N x N luma block */
void orc_intra_pred_luma(const Pel* line, int log2n, int mode, int above, int left, int bd, Pel* dst, int ds)
{
  const int n = 1 << log2n, sw = 2 * n + 1;
  Pel filt[4 * ORC_MAX_N + 1];
  int* adi = (int*)malloc(sizeof(int) * (size_t)sw * sw);
  const Pel* use = line;
  if (orc_intra_mode_filtered(mode, log2n)) { orc_intra_filter_line(line, n, filt); use = filt; }
  line_to_adi(use, n, adi);
  const int* src = adi + sw + 1;
  if (mode == 0) pred_planar(src, sw, dst, ds, n, log2n);
  else {
    pred_ang(src, sw, dst, ds, n, mode, above, left, 1, bd);
    if (mode == 1 && above && left) dc_filtering(src, sw, dst, ds, n);
  }
  free(adi);
}

/* the rough-search loop, TEncSearch.cpp:2530-2537: SATD of every mode's prediction against the original block.
 * preds (optional): 35 blocks of N x N, mode-major */
void orc_intra_rough(const Pel* line, const Pel* org, int so, int log2n, int above, int left, int bd, uint32_t sad[35], Pel* preds)
{
  const int n = 1 << log2n;
  Pel* tmp = (Pel*)malloc(sizeof(Pel) * (size_t)n * n);
  for (int mode = 0; mode < 35; mode++) {
    Pel* p = preds ? preds + (size_t)mode * n * n : tmp;
    orc_intra_pred_luma(line, log2n, mode, above, left, bd, p, n);
    sad[mode] = orc_calc_had(org, so, p, n, n, n, bd - 8);
  }
  free(tmp);
}
