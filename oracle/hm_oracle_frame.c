/*
 * hm_oracle_frame.c -- frame-level drivers over the restated HM-7.2 functions: the same work the
 * device frame pre-pass / batch entry points do, looped in C so that the CPU baseline is not
 * dominated by Python call overhead.  TEST INFRASTRUCTURE ONLY (see hm_oracle.h).
 * Citations: /root/reference/source/Lib.
 */
#include "hm_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <stddef.h>

/* PU partition census of one CTU (TEncCu.cpp:505-686 part sizes tested per CU; AMP only for
 * CU >= 16, TEncTop.cpp:523-534; inter NxN never tested).  Order documented in
 * include/thevc_cuda.h (tvc_me_census): depth, CU raster, part shape. */
void orc_census(int16_t* out /* ORC_CENSUS * 6: x, y, w, h, cu_x, cu_y */)
{
  int n = 0;
  for (int depth = 0; depth < 4; depth++) {
    int s = 64 >> depth;
    for (int cy = 0; cy < 64; cy += s)
      for (int cx = 0; cx < 64; cx += s) {
        int q = s / 4, hs = s / 2;
        int parts[13][4] = {{0, 0, s, s}, {0, 0, s, hs}, {0, hs, s, hs}, {0, 0, hs, s}, {hs, 0, hs, s},
                            {0, 0, s, q}, {0, q, s, s - q}, {0, 0, s, s - q}, {0, s - q, s, q},
                            {0, 0, q, s}, {q, 0, s - q, s}, {0, 0, s - q, s}, {s - q, 0, q, s}};
        int np = s >= 16 ? 13 : 5;
        for (int k = 0; k < np; k++) {
          out[6 * n + 0] = (int16_t)(cx + parts[k][0]); out[6 * n + 1] = (int16_t)(cy + parts[k][1]);
          out[6 * n + 2] = (int16_t)parts[k][2]; out[6 * n + 3] = (int16_t)parts[k][3];
          out[6 * n + 4] = (int16_t)cx; out[6 * n + 5] = (int16_t)cy;
          n++;
        }
      }
  }
}

/* xMotionEstimation's integer + fractional stages (TEncSearch.cpp:4120-4207) for every census PU
 * of one CTU against num_refs references.  cur / refs[r] point at pel (0,0) of padded luma planes
 * of equal stride.  pred_qpel: 2 ints per reference.  Outputs: num_refs * ORC_CENSUS entries. */
void orc_me_frame_ctu(const Pel* cur, const Pel* const* refs, int num_refs, int stride, int pic_w, int pic_h,
                      int ctu_x, int ctu_y, const int32_t* pred_qpel, uint32_t lambda_cost, int srange, int fen,
                      int hadamard, int do_frac, int bd, orc_me_result* int_out, orc_frac_result* frac_out)
{
  int16_t census[ORC_CENSUS * 6];
  orc_census(census);
  const int bi = bd - 8;
  for (int r = 0; r < num_refs; r++) {
    int predx = pred_qpel[2 * r], predy = pred_qpel[2 * r + 1];
    for (int k = 0; k < ORC_CENSUS; k++) {
      const int16_t* c = census + 6 * k;
      int x = ctu_x + c[0], y = ctu_y + c[1], w = c[2], h = c[3];
      orc_me_result* io = int_out + (size_t)r * ORC_CENSUS + k;
      orc_frac_result* fo = frac_out ? frac_out + (size_t)r * ORC_CENSUS + k : 0;
      memset(io, 0, sizeof(*io));
      if (fo) memset(fo, 0, sizeof(*fo));
      if (x + w > pic_w || y + h > pic_h) continue;
      orc_cu_geom g = {pic_w, pic_h, ctu_x + c[4], ctu_y + c[5], 64};
      int lx, ty, rx, by;
      orc_set_search_range(&g, predx, predy, srange, &lx, &ty, &rx, &by);
      const Pel* o = cur + (ptrdiff_t)y * stride + x;
      const Pel* rf = refs[r] + (ptrdiff_t)y * stride + x;
      orc_tz_search(&g, o, stride, rf, stride, w, h, lx, ty, rx, by, srange, fen, bi, lambda_cost, predx, predy,
                    predx, predy, io);
      if (fo && do_frac)
        orc_frac_search(o, stride, rf, stride, w, h, io->mvx, io->mvy, hadamard, bi, bd, lambda_cost, predx, predy, fo);
    }
  }
}

/* TComPrediction::motionCompensation for a list of uni-/bi-predicted PUs (TComPrediction.cpp:410-658):
 * pu = {x, y, w, h, ref0, mvx0, mvy0, ref1, mvx1, mvy1} (tvc_pu layout), refs[slot][plane] point at
 * pel (0,0); dst[plane] likewise. */
void orc_mc_batch(const Pel* const* ref_planes /* [slot*3 + plane] */, int stride_y, int stride_c, Pel* const* dst,
                  int n, const int32_t* pus, int bd)
{
  Pel* t0 = (Pel*)malloc(64 * 64 * sizeof(Pel));
  Pel* t1 = (Pel*)malloc(64 * 64 * sizeof(Pel));
  for (int i = 0; i < n; i++) {
    const int32_t* p = pus + 10 * i;
    int x = p[0], y = p[1], w = p[2], h = p[3];
    int use0 = p[4] >= 0, use1 = p[7] >= 0, bi = use0 && use1;
    for (int pl = 0; pl < 3; pl++) {
      int sh = pl ? 1 : 0, st = pl ? stride_c : stride_y;
      int cw = w >> sh, chh = h >> sh;
      Pel* d = dst[pl] + (ptrdiff_t)(y >> sh) * st + (x >> sh);
      for (int l = 0; l < 2; l++) {
        if (!(l ? use1 : use0)) continue;
        const Pel* rp = ref_planes[p[l ? 7 : 4] * 3 + pl] + (ptrdiff_t)(y >> sh) * st + (x >> sh);
        int mvx = p[l ? 8 : 5], mvy = p[l ? 9 : 6];
        Pel* out = bi ? (l ? t1 : t0) : d;
        int os = bi ? 64 : st;
        if (pl == 0) orc_pred_inter_luma_blk(rp, st, mvx, mvy, w, h, out, os, bi, bd);
        else orc_pred_inter_chroma_blk(rp, st, mvx, mvy, w, h, out, os, bi, bd);
      }
      if (bi) orc_add_avg(t0, 64, t1, 64, d, st, cw, chh, bd);
    }
  }
  free(t0); free(t1);
}

/* transformNxN (non-RDOQ quantiser) and invtransformNxN + reconstruction over a TU list
 * (TComTrQuant.cpp:1373-1529): tu = {plane, x, y, log2, flags, scan_idx, per, rem, base_per, coef_offset}
 * (tvc_tu layout).  flags: 1 DST, 2 transform skip.  levels may be written (forward) or read (inverse). */
void orc_fwd_tq_batch(const Pel* const* resi /* [plane] */, int stride_y, int stride_c, int n, const int32_t* tus,
                      int is_intra_slice, int sign_hide, int bd, int32_t* levels, uint32_t* abs_sum)
{
  int32_t coef[32 * 32];
  uint32_t scan[32 * 32];
  for (int i = 0; i < n; i++) {
    const int32_t* t = tus + 10 * i;
    int pl = t[0], N = 1 << t[3], st = pl ? stride_c : stride_y;
    const Pel* r = resi[pl] + (ptrdiff_t)t[2] * st + t[1];
    if (t[4] & 2) orc_transform_skip(r, st, coef, N, N, bd);
    else orc_xT((t[4] & 1) != 0, r, st, coef, N, N, bd - 8);
    orc_quant_param qp = {t[6], t[7], t[8], is_intra_slice, sign_hide, 0, bd};
    orc_scan(t[5], t[3], scan);
    uint32_t s = 0;
    orc_quant(coef, levels + t[9], 0, N, N, &qp, scan, &s);
    if (abs_sum) abs_sum[i] = s;
  }
}

/* transformNxN with the RDOQ quantiser (xT + xRateDistOptQuant) over a TU list: inter TUs (diagonal scan), luma
 * at transform depth 0 (root cbf) and chroma cbf context 5, one bit-estimate table and Lagrangian for the batch
 * (the bench workload; per-TU contexts are exercised by the RDOQ parity tests). */
void orc_fwd_rdoq_batch(const Pel* const* resi, int stride_y, int stride_c, int n, const int32_t* tus, int sign_hide, int bd,
                        const orc_est_bits* est, double lambda_luma, double lambda_chroma, int32_t* levels, uint32_t* abs_sum)
{
  int32_t coef[32 * 32];
  uint32_t scan[32 * 32];
  for (int i = 0; i < n; i++) {
    const int32_t* t = tus + 10 * i;
    int pl = t[0], N = 1 << t[3], st = pl ? stride_c : stride_y;
    const Pel* r = resi[pl] + (ptrdiff_t)t[2] * st + t[1];
    orc_xT(0, r, st, coef, N, N, bd - 8);
    orc_rdoq_param p = {t[3], pl == 0, 0, t[6], t[7], bd, pl == 0 ? -1 : 5, sign_hide, 0, pl == 0 ? lambda_luma : lambda_chroma};
    orc_scan(0, t[3], scan);
    uint32_t s = 0;
    orc_rdoq(coef, levels + t[9], 0, &p, est, scan, &s);
    if (abs_sum) abs_sum[i] = s;
  }
}

void orc_inv_tq_batch(Pel* const* resi, const Pel* const* pred, Pel* const* recon, int stride_y, int stride_c, int n,
                      const int32_t* tus, int bd, const int32_t* levels)
{
  int32_t coef[32 * 32];
  for (int i = 0; i < n; i++) {
    const int32_t* t = tus + 10 * i;
    int pl = t[0], N = 1 << t[3], st = pl ? stride_c : stride_y;
    ptrdiff_t off = (ptrdiff_t)t[2] * st + t[1];
    orc_dequant(levels + t[9], coef, N, N, t[6], t[7], bd);
    if (t[4] & 2) orc_itransform_skip(coef, resi[pl] + off, st, N, N, bd);
    else orc_xIT((t[4] & 1) != 0, coef, resi[pl] + off, st, N, N, bd - 8);
    if (pred && recon) orc_add_clip(pred[pl] + off, st, resi[pl] + off, st, recon[pl] + off, st, N, N, bd);
  }
}
