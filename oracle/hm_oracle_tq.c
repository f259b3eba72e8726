/*
 * hm_oracle_tq.c -- CPU restatement of HM-7.2 transforms, quantisation and de-quantisation.
 * TEST INFRASTRUCTURE ONLY (see hm_oracle.h).  Citations: /root/reference/source/Lib.
 */
#include "hm_oracle.h"
#include <stdlib.h>
#include <string.h>

static inline int iabs(int v) { return v < 0 ? -v : v; }
static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

/* The HEVC core transform matrix (TLibCommon/TComRom.cpp:303-377 g_aiT4/8/16/32) is the
 * published integer approximation of 64*sqrt(2)*cos((2j+1)k*pi/2N): the 32-point matrix is
 * T32[k][j] = +-C[(2j+1)k folded into a quarter period] with the 33 magnitudes below, and the
 * 16/8/4-point matrices are its even-row sub-sampled left halves.  Generated here instead of
 * tabulated; tests pin every entry against the reference's tables through libhmref.so. */
static const int16_t k_cos_mag[33] = {
  64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
  61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0
};

static int dct_coef32(int k, int j)
{
  int m = ((2 * j + 1) * k) & 127;          /* angle m*pi/64, period 128 */
  if (m <= 32) return k_cos_mag[m];
  if (m <= 64) return -k_cos_mag[64 - m];
  if (m <= 96) return -k_cos_mag[m - 64];
  return k_cos_mag[128 - m];
}

void orc_dct_matrix(int n, int16_t* out)
{
  int step = 32 / n;
  for (int k = 0; k < n; k++)
    for (int j = 0; j < n; j++) out[k * n + j] = (int16_t)dct_coef32(k * step, j);
}

/* TLibCommon/TComTrQuant.cpp:417-441 (4), 490-526 (8), 569-613 (16), 667-720 (32)
 * partialButterflyN: recursive even/odd decomposition.  dst[k*line + j] =
 * (sum_n T[k][n]*src[j*N+n] + add) >> shift stored to short WITHOUT clipping.  Written with
 * the E/O stages the reference uses, for any N in {4,8,16,32}. */
void orc_partial_butterfly(int n, const int16_t* src, int16_t* dst, int shift, int line)
{
  int16_t T[32 * 32];
  orc_dct_matrix(n, T);
  int add = 1 << (shift - 1);
  int half = n / 2;
  for (int j = 0; j < line; j++) {
    int E[16], O[16];
    for (int k = 0; k < half; k++) {
      E[k] = src[k] + src[n - 1 - k];
      O[k] = src[k] - src[n - 1 - k];
    }
    /* odd rows: dot(T[2k+1][0..half-1], O) */
    for (int k = 1; k < n; k += 2) {
      int s = 0;
      for (int m = 0; m < half; m++) s += T[k * n + m] * O[m];
      dst[k * line] = (int16_t)((s + add) >> shift);
    }
    /* even rows recurse on E (EE/EO, EEE/EEO, ...) */
    int cur[16], len = half, rstep = 2;
    memcpy(cur, E, sizeof(int) * (size_t)half);
    while (len > 1) {
      int h2 = len / 2, EE[8], EO[8];
      for (int k = 0; k < h2; k++) {
        EE[k] = cur[k] + cur[len - 1 - k];
        EO[k] = cur[k] - cur[len - 1 - k];
      }
      for (int k = rstep; k < n; k += 2 * rstep) {
        int s = 0;
        for (int m = 0; m < h2; m++) s += T[k * n + m] * EO[m];
        dst[k * line] = (int16_t)((s + add) >> shift);
      }
      memcpy(cur, EE, sizeof(int) * (size_t)h2);
      len = h2; rstep *= 2;
    }
    dst[0] = (int16_t)((T[0] * cur[0] + add) >> shift);
    src += n;
    dst++;
  }
}

/* TLibCommon/TComTrQuant.cpp:462-488 (4), 528-567 (8), 615-665 (16), 722-793 (32)
 * partialButterflyInverseN: dst[j*N + n] = Clip3(-32768, 32767, (sum_k T[k][n]*src[k*line+j] + add) >> shift).
 * The reference builds the sum from O/EO/EEO/... partial sums; integer addition is exact and
 * the magnitudes stay below 2^31, so the plain column dot product is the same number. */
void orc_partial_butterfly_inverse(int n, const int16_t* src, int16_t* dst, int shift, int line)
{
  int16_t T[32 * 32];
  orc_dct_matrix(n, T);
  int add = 1 << (shift - 1);
  for (int j = 0; j < line; j++) {
    for (int c = 0; c < n; c++) {
      int s = 0;
      for (int k = 0; k < n; k++) s += T[k * n + c] * src[k * line];
      dst[c] = (int16_t)clip3(-32768, 32767, (s + add) >> shift);
    }
    src++;
    dst += n;
  }
}

/* TLibCommon/TComTrQuant.cpp:443-460  fastForwardDst */
void orc_fast_forward_dst(const int16_t* block, int16_t* coeff, int shift)
{
  int rnd = 1 << (shift - 1);
  for (int i = 0; i < 4; i++) {
    int b0 = block[4 * i], b1 = block[4 * i + 1], b2 = block[4 * i + 2], b3 = block[4 * i + 3];
    int c0 = b0 + b3, c1 = b1 + b3, c2 = b0 - b1, c3 = 74 * b2;
    coeff[i]      = (int16_t)((29 * c0 + 55 * c1 + c3 + rnd) >> shift);
    coeff[4 + i]  = (int16_t)((74 * (b0 + b1 - b3) + rnd) >> shift);
    coeff[8 + i]  = (int16_t)((29 * c2 + 55 * c0 - c3 + rnd) >> shift);
    coeff[12 + i] = (int16_t)((55 * c2 - 29 * c1 + c3 + rnd) >> shift);
  }
}

/* TLibCommon/TComTrQuant.cpp:462-479  fastInverseDst */
void orc_fast_inverse_dst(const int16_t* tmp, int16_t* block, int shift)
{
  int rnd = 1 << (shift - 1);
  for (int i = 0; i < 4; i++) {
    int t0 = tmp[i], t1 = tmp[4 + i], t2 = tmp[8 + i], t3 = tmp[12 + i];
    int c0 = t0 + t2, c1 = t2 + t3, c2 = t0 - t3, c3 = 74 * t1;
    block[4 * i + 0] = (int16_t)clip3(-32768, 32767, (29 * c0 + 55 * c1 + c3 + rnd) >> shift);
    block[4 * i + 1] = (int16_t)clip3(-32768, 32767, (55 * c2 - 29 * c1 + c3 + rnd) >> shift);
    block[4 * i + 2] = (int16_t)clip3(-32768, 32767, (74 * (t0 - t2 + t3) + rnd) >> shift);
    block[4 * i + 3] = (int16_t)clip3(-32768, 32767, (55 * c0 + 29 * c2 - c3 + rnd) >> shift);
  }
}

static int log2i(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

/* TLibCommon/TComTrQuant.cpp:803-885  xTrMxN (square sizes; NSQT compiled out).
 * shift_1st = log2(N) - 1 + bitIncrement, shift_2nd = log2(N) + 6; 4x4 uses the DST in both
 * passes when uiMode != REG_DCT (INTRA_TRANS_SIMP). */
void orc_xTrMxN(const int16_t* block, int16_t* coeff, int w, int h, int use_dst, int bi)
{
  int s1 = log2i(w) - 1 + bi, s2 = log2i(h) + 6;
  int16_t tmp[32 * 32];
  if (w == 4 && h == 4 && use_dst) {
    orc_fast_forward_dst(block, tmp, s1);
    orc_fast_forward_dst(tmp, coeff, s2);
  } else {
    orc_partial_butterfly(w, block, tmp, s1, h);
    orc_partial_butterfly(h, tmp, coeff, s2, w);
  }
}

/* TLibCommon/TComTrQuant.cpp:892-972  xITrMxN: shift_1st = 7, shift_2nd = 12 - bitIncrement */
void orc_xITrMxN(const int16_t* coeff, int16_t* block, int w, int h, int use_dst, int bi)
{
  int s1 = 7, s2 = 12 - bi;
  int16_t tmp[32 * 32];
  if (w == 4 && h == 4 && use_dst) {
    orc_fast_inverse_dst(coeff, tmp, s1);
    orc_fast_inverse_dst(tmp, block, s2);
  } else {
    orc_partial_butterfly_inverse(h, coeff, tmp, s1, w);
    orc_partial_butterfly_inverse(w, tmp, block, s2, h);
  }
}

/* TLibCommon/TComTrQuant.cpp:1542-1573  xT: gather strided residual, transform, widen to Int */
void orc_xT(int use_dst, const Pel* resi, int stride, int32_t* coeff, int w, int h, int bi)
{
  int16_t block[32 * 32], c16[32 * 32];
  for (int j = 0; j < h; j++) memcpy(block + j * w, resi + j * stride, (size_t)w * sizeof(int16_t));
  orc_xTrMxN(block, c16, w, h, use_dst, bi);
  for (int j = 0; j < w * h; j++) coeff[j] = c16[j];
}

/* TLibCommon/TComTrQuant.cpp:1583-1615  xIT: (short) cast of the Int coefficients (:1602) */
void orc_xIT(int use_dst, const int32_t* coeff, Pel* resi, int stride, int w, int h, int bi)
{
  int16_t block[32 * 32], c16[32 * 32];
  for (int j = 0; j < w * h; j++) c16[j] = (int16_t)coeff[j];
  orc_xITrMxN(c16, block, w, h, use_dst, bi);
  for (int j = 0; j < h; j++) memcpy(resi + j * stride, block + j * w, (size_t)w * sizeof(int16_t));
}

/* TLibCommon/TComTrQuant.cpp:1622-1660  xTransformSkip (note psCoeff[j*height + k], :1641) */
void orc_transform_skip(const Pel* resi, int stride, int32_t* coeff, int w, int h, int bd)
{
  int shift = 15 - bd - log2i(w);
  if (shift >= 0) {
    for (int j = 0; j < h; j++)
      for (int k = 0; k < w; k++) coeff[j * h + k] = resi[j * stride + k] * (1 << shift);
  } else {
    int ts = -shift, off = 1 << (ts - 1);
    for (int j = 0; j < h; j++)
      for (int k = 0; k < w; k++) coeff[j * h + k] = (resi[j * stride + k] + off) >> ts;
  }
}

/* TLibCommon/TComTrQuant.cpp:1668-1704  xITransformSkip */
void orc_itransform_skip(const int32_t* coeff, Pel* resi, int stride, int w, int h, int bd)
{
  int shift = 15 - bd - log2i(w);
  if (shift > 0) {
    int off = 1 << (shift - 1);
    for (int j = 0; j < h; j++)
      for (int k = 0; k < w; k++) resi[j * stride + k] = (Pel)((coeff[j * w + k] + off) >> shift);
  } else {
    int ts = -shift;
    for (int j = 0; j < h; j++)
      for (int k = 0; k < w; k++) resi[j * stride + k] = (Pel)(coeff[j * w + k] * (1 << ts));
  }
}

/* TLibCommon/TComRom.cpp:380-386 g_aucChromaScale (CHROMA_QP_EXTENSION) is the chroma QP
 * mapping table of the H.265 text: identity below 30, compressed 30..43, qp-6 above. */
static int chroma_scale(int qp)
{
  static const uint8_t mid[14] = { 29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37 };
  if (qp < 30) return qp;
  if (qp < 44) return mid[qp - 30];
  return qp - 6;
}

/* TLibCommon/TComTrQuant.cpp:192-222  setQPforQuant + TComTrQuant.h:91-97 QpParam::setQpParam */
void orc_set_qp(int qpy, int is_luma, int qp_bd_offset, int chroma_qp_offset, int* per, int* rem)
{
  int q;
  if (is_luma) {
    q = qpy + qp_bd_offset;
  } else {
    q = clip3(-qp_bd_offset, 57, qpy + chroma_qp_offset);
    if (q < 0) q = q + qp_bd_offset;
    else q = chroma_scale(q) + qp_bd_offset;
  }
  *per = q / 6;
  *rem = q % 6;
}

/* TLibCommon/TComRom.cpp:564-690  initSigLastScan: up-right diagonal inside 4x4 coefficient
 * groups with the groups themselves visited along up-right diagonals (DIAG), and the
 * group-wise horizontal / vertical scans (REMOVAL_8x2_2x8_CG). */
static void diag_scan_square(int n, int* ys, int* xs)
{
  int pos = 0;
  for (int line = 0; pos < n * n; line++) {
    int prim = line, scnd = 0;
    while (prim >= n) { scnd++; prim--; }
    while (prim >= 0 && scnd < n) { ys[pos] = prim; xs[pos] = scnd; pos++; scnd++; prim--; }
  }
}

void orc_scan(int scan_idx, int log2size, uint32_t* out)
{
  int n = 1 << log2size, cnt = 0;
  if (scan_idx == 0) {
    int iy[16], ix[16];
    diag_scan_square(4, iy, ix);
    if (n == 4) {
      for (int i = 0; i < 16; i++) out[i] = (uint32_t)(iy[i] * 4 + ix[i]);
      return;
    }
    int nb = n >> 2, gy[64], gx[64];
    diag_scan_square(nb, gy, gx);
    for (int b = 0; b < nb * nb; b++) {
      int offs = 4 * (gx[b] + gy[b] * n);
      for (int i = 0; i < 16; i++) out[16 * b + i] = (uint32_t)(iy[i] * n + ix[i] + offs);
    }
  } else if (scan_idx == 1) {
    int nb = n >> 2;
    for (int by = 0; by < nb; by++)
      for (int bx = 0; bx < nb; bx++)
        for (int y = 0; y < 4; y++)
          for (int x = 0; x < 4; x++) out[cnt++] = (uint32_t)((by * 4 + y) * n + bx * 4 + x);
  } else {
    int nb = n >> 2;
    for (int bx = 0; bx < nb; bx++)
      for (int by = 0; by < nb; by++)
        for (int x = 0; x < 4; x++)
          for (int y = 0; y < 4; y++) out[cnt++] = (uint32_t)((by * 4 + y) * n + bx * 4 + x);
  }
}

static const int k_quant_scales[6] = { 26214, 23302, 20560, 18396, 16384, 14564 };   /* TComRom.cpp:293-296 */
static const int k_inv_quant_scales[6] = { 40, 45, 51, 57, 64, 72 };                 /* TComRom.cpp:298-301 */

/* TLibCommon/TComTrQuant.cpp:977-1100  signBitHidingHDQ */
static void sign_bit_hiding_hdq(int32_t* q, const int32_t* coef, const uint32_t* scan, const int* deltaU, int w, int h)
{
  int lastCG = -1;
  for (int subSet = (w * h - 1) >> 4; subSet >= 0; subSet--) {
    int subPos = subSet << 4;
    int firstNZ = 16, lastNZ = -1, absSum = 0, n;
    for (n = 15; n >= 0; --n) if (q[scan[n + subPos]]) { lastNZ = n; break; }
    for (n = 0; n < 16; n++) if (q[scan[n + subPos]]) { firstNZ = n; break; }
    for (n = firstNZ; n <= lastNZ; n++) absSum += q[scan[n + subPos]];
    if (lastNZ >= 0 && lastCG == -1) lastCG = 1;
    if (lastNZ - firstNZ >= 4) {                                   /* SBH_THRESHOLD */
      uint32_t signbit = (q[scan[subPos + firstNZ]] > 0) ? 0u : 1u;
      if (signbit != (uint32_t)(absSum & 1)) {
        int minCostInc = 2147483647, minPos = -1, finalChange = 0, curCost = 2147483647, curChange = 0;
        for (n = (lastCG == 1 ? lastNZ : 15); n >= 0; --n) {
          uint32_t blkPos = scan[n + subPos];
          if (q[blkPos] != 0) {
            if (deltaU[blkPos] > 0) { curCost = -deltaU[blkPos]; curChange = 1; }
            else {
              if (n == firstNZ && iabs(q[blkPos]) == 1) curCost = 2147483647;
              else { curCost = deltaU[blkPos]; curChange = -1; }
            }
          } else {
            if (n < firstNZ) {
              uint32_t thisSign = (coef[blkPos] >= 0) ? 0u : 1u;
              if (thisSign != signbit) curCost = 2147483647;
              else { curCost = -deltaU[blkPos]; curChange = 1; }
            } else { curCost = -deltaU[blkPos]; curChange = 1; }
          }
          if (curCost < minCostInc) { minCostInc = curCost; finalChange = curChange; minPos = (int)blkPos; }
        }
        if (q[minPos] == 32767 || q[minPos] == -32768) finalChange = -1;
        if (coef[minPos] >= 0) q[minPos] += finalChange; else q[minPos] -= finalChange;
      }
    }
    if (lastCG == 1) lastCG = 0;
  }
}

/* TLibCommon/TComTrQuant.cpp:1102-1270  xQuant, non-RDOQ branch, flat scaling list
 * (m_quantCoef = g_quantScales[rem], :2904-2920), ADAPTIVE_QP_SELECTION on: qbits/add use the
 * slice base QP's `per` (:1226-1231).  abs_sum accumulates like uiAcSum (caller zeroes it). */
void orc_quant(const int32_t* coef, int32_t* qcoef, int32_t* arl, int w, int h,
               const orc_quant_param* p, const uint32_t* scan, uint32_t* abs_sum)
{
  int log2 = log2i(w);
  int tshift = 15 - p->bd - log2;
  int qscale = k_quant_scales[p->qp_rem];
  int qbits = 14 + p->base_per + tshift;
  int add = (p->is_intra_slice ? 171 : 85) << (qbits - 9);
  int qbitsC = 14 + p->base_per + tshift - 7;
  int addC = 1 << (qbitsC - 1);
  int qbits8 = qbits - 8;
  int deltaU[32 * 32];
  uint32_t acsum = *abs_sum;
  for (int n = 0; n < w * h; n++) {
    int level = coef[n];
    int sign = (level < 0) ? -1 : 1;
    int64_t tmp = (int64_t)iabs(level) * qscale;
    if (p->use_arl && arl) arl[n] = (int32_t)((tmp + addC) >> qbitsC);
    level = (int)((tmp + add) >> qbits);
    deltaU[n] = (int)((tmp - (int64_t)(int32_t)((uint32_t)level << qbits)) >> qbits8);
    acsum += (uint32_t)level;
    level *= sign;
    qcoef[n] = clip3(-32768, 32767, level);
  }
  *abs_sum = acsum;
  if (p->sign_hide && acsum >= 2) sign_bit_hiding_hdq(qcoef, coef, scan, deltaU, w, h);
}

/* TLibCommon/TComTrQuant.cpp:1272-1355  xDeQuant, flat branch (:1343-1353): the product
 * clipQCoef*scale is evaluated in 32-bit Int and wraps for large levels at high QP. */
void orc_dequant(const int32_t* qcoef, int32_t* coef, int w, int h, int per, int rem, int bd)
{
  int log2 = log2i(w);
  int tshift = 15 - bd - log2;
  int shift = 20 - 14 - tshift;
  int add = 1 << (shift - 1);
  int scale = k_inv_quant_scales[rem] << per;
  for (int n = 0; n < w * h; n++) {
    int c = clip3(-32768, 32767, qcoef[n]);
    int v = (int)((uint32_t)c * (uint32_t)scale + (uint32_t)add) >> shift;
    coef[n] = clip3(-32768, 32767, v);
  }
}
