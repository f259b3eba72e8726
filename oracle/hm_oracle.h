/*
 * hm_oracle.h -- CPU restatement (plain C99) of the HM-7.2 hot-path arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle for the CUDA path in
 * thevc_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product library
 * (libthevc_cuda.so) never links, loads or calls anything in this directory.
 *
 * Every function cites the reference file:line (relative to
 * /root/reference/source/Lib) whose arithmetic it restates.  The restatement is
 * pinned against the reference's own compiled functions (oracle/_ref/libhmref.so,
 * built by oracle/Makefile from the reference sources where they lie) by
 * tests/test_oracle_vs_ref.py and against the committed vectors in
 * tests/golden/ (generated from libhmref.so by tests/golden/make_golden.py).
 *
 * Conventions (same as the reference): Pel = int16, TCoeff = int32, strides in
 * elements, `bi` = g_uiBitIncrement (0 for *_main, 2 for he10),
 * `bd` = g_uiBitDepth + g_uiBitIncrement (8 or 10).
 */
#ifndef HM_ORACLE_H
#define HM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int16_t Pel;

/* ------------------------------------------------------------------ distortion */
uint32_t orc_sad(const Pel* org, int so, const Pel* cur, int sc, int w, int h, int subshift, int bi);
uint32_t orc_sad_generic(const Pel* org, int so, const Pel* cur, int sc, int w, int h, int bi);
uint32_t orc_sse(const Pel* org, int so, const Pel* cur, int sc, int w, int h, int bi);
uint32_t orc_hads(const Pel* org, int so, const Pel* cur, int sc, int w, int h, int bi);
uint32_t orc_calc_had(const Pel* p0, int s0, const Pel* p1, int s1, int w, int h, int bi);
uint32_t orc_had2x2(const Pel* org, const Pel* cur, int so, int sc);
uint32_t orc_had4x4(const Pel* org, const Pel* cur, int so, int sc);
uint32_t orc_had8x8(const Pel* org, const Pel* cur, int so, int sc);
/* TComRdCost::getDistPart dispatch: dfunc = DF_SSE(1) / DF_SAD(8) / DF_HADS(22) */
uint32_t orc_get_dist_part(const Pel* cur, int sc, const Pel* org, int so, int w, int h, int dfunc, int bi);

/* motion-vector rate */
uint32_t orc_mv_component_bits(int v);
uint32_t orc_mv_bits(int x, int y, int scale, int predx, int predy);
uint32_t orc_mv_cost(uint32_t lambda_cost, int x, int y, int scale, int predx, int predy);
uint32_t orc_lambda_motion_sad(double lambda);

/* ------------------------------------------------------------------ interpolation */
void orc_filter_copy(const Pel* src, int ss, Pel* dst, int ds, int w, int h, int isFirst, int isLast, int bd);
void orc_filter(int ntaps, int isVert, int isFirst, int isLast, const Pel* src, int ss, Pel* dst, int ds,
                int w, int h, const int16_t* coeff, int bd);
void orc_filter_hor_luma(const Pel* src, int ss, Pel* dst, int ds, int w, int h, int frac, int isLast, int bd);
void orc_filter_ver_luma(const Pel* src, int ss, Pel* dst, int ds, int w, int h, int frac, int isFirst, int isLast, int bd);
void orc_filter_hor_chroma(const Pel* src, int ss, Pel* dst, int ds, int w, int h, int frac, int isLast, int bd);
void orc_filter_ver_chroma(const Pel* src, int ss, Pel* dst, int ds, int w, int h, int frac, int isFirst, int isLast, int bd);

/* ------------------------------------------------------------------ motion compensation */
void orc_pred_inter_luma_blk(const Pel* ref, int rs, int mvx, int mvy, int w, int h, Pel* dst, int ds, int bi_flag, int bd);
void orc_pred_inter_chroma_blk(const Pel* ref, int rs, int mvx, int mvy, int w, int h, Pel* dst, int ds, int bi_flag, int bd);
void orc_add_avg(const Pel* s0, int st0, const Pel* s1, int st1, Pel* dst, int ds, int w, int h, int bd);
void orc_subtract(const Pel* s0, int st0, const Pel* s1, int st1, Pel* dst, int ds, int w, int h);
void orc_add_clip(const Pel* s0, int st0, const Pel* s1, int st1, Pel* dst, int ds, int w, int h, int bd);
void orc_remove_high_freq(Pel* dst, int ds, const Pel* src, int ss, int w, int h);
void orc_extend_border(Pel* pic, int stride, int w, int h, int mx, int my);

/* ------------------------------------------------------------------ motion estimation */
typedef struct {
  int pic_w, pic_h;   /* SPS luma size                                      */
  int cu_x, cu_y;     /* m_uiCUPelX/Y of the TComDataCU doing the search     */
  int max_cu;         /* g_uiMaxCUWidth == g_uiMaxCUHeight                   */
} orc_cu_geom;

typedef struct {
  int mvx, mvy;       /* integer-pel best                                   */
  uint32_t sad;       /* ruiSAD = best cost minus its rate term             */
  uint32_t n_sads;    /* number of xTZSearchHelp / candidate evaluations    */
} orc_me_result;

void orc_clip_mv(const orc_cu_geom* g, int* mvx, int* mvy);
void orc_set_search_range(const orc_cu_geom* g, int predx, int predy, int srange,
                          int* lx, int* ty, int* rx, int* by);
/* org: PU block; ref: pointer to the co-located pel of the PU in the padded reference plane */
void orc_pattern_search(const Pel* org, int so, const Pel* ref, int rs, int w, int h,
                        int lx, int ty, int rx, int by, int fen, int bi,
                        uint32_t lambda_cost, int predx, int predy, orc_me_result* out);
void orc_tz_search(const orc_cu_geom* g, const Pel* org, int so, const Pel* ref, int rs, int w, int h,
                   int lx, int ty, int rx, int by, int srange, int fen, int bi,
                   uint32_t lambda_cost, int predx, int predy, int startx_q, int starty_q,
                   orc_me_result* out);

typedef struct {
  int halfx, halfy;   /* rcMvHalf  (-1..1)                                  */
  int qtrx, qtry;     /* rcMvQter  (-1..1)                                  */
  uint32_t cost_half; /* ruiCost after half-pel refinement                  */
  uint32_t cost;      /* ruiCost after quarter-pel refinement               */
} orc_frac_result;

/* xPatternSearchFracDIF: ref points at the PU's co-located pel; (imvx,imvy) integer MV */
void orc_frac_search(const Pel* org, int so, const Pel* ref, int rs, int w, int h,
                     int imvx, int imvy, int hadamard, int bi, int bd,
                     uint32_t lambda_cost, int predx, int predy, orc_frac_result* out);

/* ------------------------------------------------------------------ transform / quant */
void orc_dct_matrix(int n, int16_t* out /* n*n row-major T[k][j] */);
void orc_partial_butterfly(int n, const int16_t* src, int16_t* dst, int shift, int line);
void orc_partial_butterfly_inverse(int n, const int16_t* src, int16_t* dst, int shift, int line);
void orc_fast_forward_dst(const int16_t* block, int16_t* coeff, int shift);
void orc_fast_inverse_dst(const int16_t* tmp, int16_t* block, int shift);
void orc_xTrMxN(const int16_t* block, int16_t* coeff, int w, int h, int use_dst, int bi);
void orc_xITrMxN(const int16_t* coeff, int16_t* block, int w, int h, int use_dst, int bi);
void orc_xT(int use_dst, const Pel* resi, int stride, int32_t* coeff, int w, int h, int bi);
void orc_xIT(int use_dst, const int32_t* coeff, Pel* resi, int stride, int w, int h, int bi);
void orc_transform_skip(const Pel* resi, int stride, int32_t* coeff, int w, int h, int bd);
void orc_itransform_skip(const int32_t* coeff, Pel* resi, int stride, int w, int h, int bd);

typedef struct {
  int qp_per, qp_rem;   /* m_cQP after setQPforQuant                        */
  int base_per;         /* cQpBase.m_iPer (ADAPTIVE_QP_SELECTION)           */
  int is_intra_slice;   /* I_SLICE -> 171 else 85                           */
  int sign_hide;        /* PPS sign-data-hiding flag                        */
  int use_arl;          /* m_bUseAdaptQpSelect                              */
  int bd;               /* g_uiBitDepth + g_uiBitIncrement                  */
} orc_quant_param;

void orc_set_qp(int qpy, int is_luma, int qp_bd_offset, int chroma_qp_offset, int* per, int* rem);
void orc_scan(int scan_idx /*0 diag,1 hor,2 ver*/, int log2size, uint32_t* out);
/* non-RDOQ branch of xQuant + signBitHidingHDQ, flat scaling list */
void orc_quant(const int32_t* coef, int32_t* qcoef, int32_t* arl, int w, int h,
               const orc_quant_param* qp, const uint32_t* scan, uint32_t* abs_sum);
void orc_dequant(const int32_t* qcoef, int32_t* coef, int w, int h, int per, int rem, int bd);

/* ------------------------------------------------------------------ RDOQ (hm_oracle_rdoq.c) */
/* estBitsSbacStruct, TComTrQuant.h:59-72: same members, same order, plain ints */
typedef struct {
  int32_t sig_cg[2][2];          /* significantCoeffGroupBits[NUM_SIG_CG_FLAG_CTX][2]   */
  int32_t sig[42][2];            /* significantBits[NUM_SIG_FLAG_CTX][2]                */
  int32_t last_x[32];            /* lastXBits                                           */
  int32_t last_y[32];            /* lastYBits                                           */
  int32_t greater_one[24][2];    /* m_greaterOneBits[NUM_ONE_FLAG_CTX][2]               */
  int32_t level_abs[6][2];       /* m_levelAbsBits[NUM_ABS_FLAG_CTX][2]                 */
  int32_t block_cbp[15][2];      /* blockCbpBits[3*NUM_QT_CBF_CTX][2]                   */
  int32_t block_root_cbp[4][2];  /* blockRootCbpBits[4][2]                              */
  int32_t scan_zigzag[2], scan_non_zigzag[2];
} orc_est_bits;

typedef struct {
  int log2_size;        /* 2..5                                              */
  int is_luma;          /* eTType == TEXT_LUMA                               */
  int scan_idx;         /* 0 diag, 1 hor, 2 ver (zigzag is mapped to diag)   */
  int qp_per, qp_rem;   /* m_cQP after setQPforQuant                         */
  int bd;               /* g_uiBitDepth + g_uiBitIncrement                   */
  int cbf_ctx;          /* < 0: inter luma TU at transform depth 0 (root cbf);
                           else index into blockCbpBits (chroma offset included) */
  int sign_hide;        /* PPS sign-data-hiding flag                         */
  int use_arl;          /* m_bUseAdaptQpSelect                               */
  double lambda;        /* m_dLambda (selectLambda)                          */
} orc_rdoq_param;

double orc_rdoq_err_scale(int log2_size, int qp_rem, int bd);
void orc_rdoq(const int32_t* coef, int32_t* qcoef, int32_t* arl, const orc_rdoq_param* p,
              const orc_est_bits* est, const uint32_t* scan, uint32_t* abs_sum);

/* ------------------------------------------------------------------ deblocking (hm_oracle_deblock.c) */
typedef struct { uint8_t bs, qp, flags, reserved; } orc_dbk_unit;      /* same layout as tvc_dbk_unit */
void orc_deblock_pic(Pel* Y, int sy, Pel* U, Pel* V, int sc, int width, int height, const orc_dbk_unit* ver,
                     const orc_dbk_unit* hor, int beta_off2, int tc_off2, int bd);
typedef struct { int16_t type; int16_t eo[5]; int16_t bo[32]; } orc_sao_unit;      /* same layout as tvc_sao_unit */
void orc_sao_plane(const Pel* src, Pel* dst, int stride, int w, int h, int ctu, int ctus_x, const orc_sao_unit* units, int bd);

/* ------------------------------------------------------------------ frame-level drivers (hm_oracle_frame.c) */
#define ORC_CENSUS 593
void orc_census(int16_t* out /* ORC_CENSUS * 6: x, y, w, h, cu_x, cu_y */);
void orc_me_frame_ctu(const Pel* cur, const Pel* const* refs, int num_refs, int stride, int pic_w, int pic_h,
                      int ctu_x, int ctu_y, const int32_t* pred_qpel, uint32_t lambda_cost, int srange, int fen,
                      int hadamard, int do_frac, int bd, orc_me_result* int_out, orc_frac_result* frac_out);
void orc_mc_batch(const Pel* const* ref_planes, int stride_y, int stride_c, Pel* const* dst, int n, const int32_t* pus, int bd);
void orc_fwd_tq_batch(const Pel* const* resi, int stride_y, int stride_c, int n, const int32_t* tus,
                      int is_intra_slice, int sign_hide, int bd, int32_t* levels, uint32_t* abs_sum);
void orc_fwd_rdoq_batch(const Pel* const* resi, int stride_y, int stride_c, int n, const int32_t* tus, int sign_hide, int bd,
                        const orc_est_bits* est, double lambda_luma, double lambda_chroma, int32_t* levels, uint32_t* abs_sum);
void orc_inv_tq_batch(Pel* const* resi, const Pel* const* pred, Pel* const* recon, int stride_y, int stride_c, int n,
                      const int32_t* tus, int bd, const int32_t* levels);

/* ------------------------------------------------------------------ intra rough search (hm_oracle_intra.c)
 * line: 4N+1 reference samples (left column bottom to top, corner, row above left to right), see the file header */
void orc_intra_filter_line(const Pel* line, int n, Pel* out);
int orc_intra_mode_filtered(int mode, int log2n);
void orc_intra_pred_luma(const Pel* line, int log2n, int mode, int above, int left, int bd, Pel* dst, int ds);
void orc_intra_rough(const Pel* line, const Pel* org, int so, int log2n, int above, int left, int bd, uint32_t sad[35], Pel* preds);

/* ------------------------------------------------------------------ picture hashes / PSNR sums (hm_oracle_hash.c) */
void orc_md5_plane(const Pel* plane, int w, int h, int stride, int bd, unsigned char digest[16]);
void orc_crc_plane(const Pel* plane, int w, int h, int stride, int bd, unsigned char digest[16]);
void orc_checksum_plane(const Pel* plane, int w, int h, int stride, int bd, unsigned char digest[16]);
uint64_t orc_ssd_plane(const Pel* a, int sa, const Pel* b, int sb, int w, int h);

#ifdef __cplusplus
}
#endif
#endif
