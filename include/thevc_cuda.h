/*
 * thevc_cuda.h -- C ABI of TLibCuda, the B200 (sm_100a) implementation of the data-parallel
 * inner loops of the HM-7.2 HEVC codec (fr34k8/thevc).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.  The reference
 * has no FFI of its own; its boundary is the public C++ API of four leaf classes, so every entry
 * point below cites the reference interface it replaces (file:line under
 * /root/reference/source/Lib).  The HM-side shim that binds these (replacement
 * TComRdCost / TComInterpolationFilter / TComTrQuant / TComPrediction / TEncSearch members) lives
 * in thevc_b200/host/ and is described in INTEGRATION.md.
 *
 * Conventions (identical to the reference): Pel = int16_t, TCoeff = int32_t, strides in ELEMENTS,
 * motion vectors in quarter luma pels unless a name says otherwise, pictures are planar 4:2:0 with
 * a replicated margin of max_cu+16 luma pels (TComPicYuv.cpp:71-127).
 *
 * Error behaviour: every function returns TVC_OK (0) or a TVC_ERR_* code and never aborts; the HM
 * shim turns a non-zero status into exit(EXIT_FAILURE) like the reference's own fatal paths
 * (CommonDef.h:143).  There is NO CPU fallback: without a CUDA device tvc_ctx_create fails.
 *
 * Threading: one host thread per context (the reference is single threaded, process-global
 * state).  Host-pointer entry points are synchronous: results are valid on return.  *_dev entry
 * points take device pointers, are asynchronous on the context stream and do not synchronise.
 */
#ifndef THEVC_CUDA_H
#define THEVC_CUDA_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TVC_ABI_VERSION 1

enum {
  TVC_OK = 0,
  TVC_ERR_ARG = 1,      /* invalid argument / unsupported size */
  TVC_ERR_CUDA = 2,     /* CUDA runtime error; see tvc_last_error */
  TVC_ERR_NOMEM = 3,
  TVC_ERR_STATE = 4     /* e.g. tables queried before tvc_me_prepass */
};

typedef struct tvc_ctx tvc_ctx;

typedef struct {
  int width, height;    /* luma picture size (SPS)                                       */
  int bit_depth;        /* g_uiBitDepth + g_uiBitIncrement: 8 (main) or 10 (he10)         */
  int max_cu;           /* g_uiMaxCUWidth (64)                                            */
  int num_slots;        /* number of device picture slots (cur, refs, pred, resi, recon)  */
  int device;           /* CUDA device ordinal                                            */
} tvc_config;

/* ---------------------------------------------------------------------------------- context */
int  tvc_abi_version(void);
int  tvc_ctx_create(const tvc_config* cfg, tvc_ctx** out);
void tvc_ctx_destroy(tvc_ctx* ctx);
/* use a caller-owned cudaStream_t (e.g. torch's current stream) for all launches of this ctx */
int  tvc_ctx_set_stream(tvc_ctx* ctx, void* cuda_stream);
int  tvc_sync(tvc_ctx* ctx);
const char* tvc_last_error(tvc_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
uint64_t tvc_launch_count(tvc_ctx* ctx);

/* ---------------------------------------------------------------------------------- pictures
 * A slot is a device-resident TComPicYuv: three int16 planes with margins, plus (bit_depth 8) a
 * packed u8 copy of the luma plane used by the integer-ME kernels.
 * Replaces: TComPicYuv::create / extendPicBorder / copyToPic (TComPicYuv.cpp:71-127, 241-286). */
enum { TVC_PLANE_Y = 0, TVC_PLANE_U = 1, TVC_PLANE_V = 2 };

/* y,u,v point at pel (0,0) of the host planes.  with_margin != 0 also copies the margins (the
 * host picture was already border-extended, e.g. a reference reconstruction after
 * TComSlice::setRefPicList, TComSlice.cpp:402-431).                                          */
int tvc_pic_upload(tvc_ctx* ctx, int slot, const int16_t* y, int stride_y,
                   const int16_t* u, const int16_t* v, int stride_c, int with_margin);
int tvc_pic_download(tvc_ctx* ctx, int slot, int16_t* y, int stride_y,
                     int16_t* u, int16_t* v, int stride_c, int with_margin);
/* TComPicYuv::extendPicBorder (TComPicYuv.cpp:241-286) on the device; also refreshes the u8 copy */
int tvc_pic_extend_border(tvc_ctx* ctx, int slot);
/* device pointer to pel (0,0) of a plane and its stride in elements (for *_dev users)         */
int tvc_pic_device_ptr(tvc_ctx* ctx, int slot, int plane, void** ptr, int* stride);
/* same for the packed u8 luma copy (NULL when bit_depth != 8)                                */
int tvc_pic_device_ptr_u8(tvc_ctx* ctx, int slot, void** ptr, int* stride);

/* region ops on slots, one plane: TComYuv::subtract* / addClip* / removeHighFreq
 * (TComYuv.cpp:401-518, 583-633).  dst = a - b ; dst = Clip(a + b) ; dst = 2*dst - a          */
int tvc_pic_subtract(tvc_ctx* ctx, int dst_slot, int a_slot, int b_slot, int plane, int x, int y, int w, int h);
int tvc_pic_add_clip(tvc_ctx* ctx, int dst_slot, int a_slot, int b_slot, int plane, int x, int y, int w, int h);
int tvc_pic_remove_high_freq(tvc_ctx* ctx, int dst_slot, int a_slot, int plane, int x, int y, int w, int h);

/* ---------------------------------------------------------------------------------- distortion
 * Replaces TComRdCost::xGetSAD* / xGetSSE* / xGetHADs* reached through DistParam::DistFunc,
 * setDistParam and getDistPart / calcHAD (TComRdCost.cpp:286-478, 490-989, 1314-1656,
 * 1663-1872, 2122-2287).                                                                     */
enum { TVC_DIST_SAD = 0, TVC_DIST_SSE = 1, TVC_DIST_HADS = 2 };

/* Drop-in for one DistFunc call on host buffers (copies both blocks, one launch, one result).
 * sub_shift as DistParam::iSubShift (honoured for SAD only); result already >> bitIncrement.   */
int tvc_dist_block(tvc_ctx* ctx, int kind, const int16_t* org, int stride_org,
                   const int16_t* cur, int stride_cur, int w, int h, int sub_shift, uint32_t* out);

typedef struct {
  int32_t kind;                                   /* TVC_DIST_*                                */
  int32_t org_slot, org_plane, org_x, org_y;      /* block origin in a picture slot            */
  int32_t cur_slot, cur_plane, cur_x, cur_y;      /* may reach into the margin                 */
  int32_t w, h, sub_shift;
} tvc_dist_job;

/* n independent distortions over device-resident pictures; jobs and out are host arrays       */
int tvc_dist_batch(tvc_ctx* ctx, int n, const tvc_dist_job* jobs, uint32_t* out);
/* same with device-resident jobs/out, asynchronous                                            */
int tvc_dist_batch_dev(tvc_ctx* ctx, int n, const tvc_dist_job* jobs_dev, uint32_t* out_dev);

/* ---------------------------------------------------------------------------------- interpolation
 * Drop-in for TComInterpolationFilter::filterHorLuma / filterVerLuma / filterHorChroma /
 * filterVerChroma (TComInterpolationFilter.h:73-76, .cpp:325-415) on host buffers.  src points at
 * the first output sample's source position exactly as in the reference (taps reach
 * (N/2-1) elements before it and N/2 after).                                                  */
int tvc_filter_hor_luma(tvc_ctx* ctx, const int16_t* src, int src_stride, int16_t* dst, int dst_stride,
                        int w, int h, int frac, int is_last);
int tvc_filter_ver_luma(tvc_ctx* ctx, const int16_t* src, int src_stride, int16_t* dst, int dst_stride,
                        int w, int h, int frac, int is_first, int is_last);
int tvc_filter_hor_chroma(tvc_ctx* ctx, const int16_t* src, int src_stride, int16_t* dst, int dst_stride,
                          int w, int h, int frac, int is_last);
int tvc_filter_ver_chroma(tvc_ctx* ctx, const int16_t* src, int src_stride, int16_t* dst, int dst_stride,
                          int w, int h, int frac, int is_first, int is_last);

/* ---------------------------------------------------------------------------------- motion compensation
 * Replaces TComPrediction::motionCompensation -> xPredInterUni/Bi -> xPredInterLumaBlk /
 * xPredInterChromaBlk -> xWeightedAverage/TComYuv::addAvg (TComPrediction.cpp:410-658,
 * TComYuv.cpp:520-581) for a list of PUs of one picture.  MVs must already be clipped
 * (TComDataCU::clipMv) as xPredInterUni does (TComPrediction.cpp:485).                         */
typedef struct {
  int32_t x, y, w, h;          /* luma PU rectangle in the picture                              */
  int32_t ref_slot0, mvx0, mvy0;   /* list-0 reference slot (-1: unused) and quarter-pel MV     */
  int32_t ref_slot1, mvx1, mvy1;   /* list-1 reference slot (-1: unused)                        */
} tvc_pu;

/* writes Y, U and V predictions of every PU into dst_slot                                     */
int tvc_mc_batch(tvc_ctx* ctx, int dst_slot, int n, const tvc_pu* pus);
int tvc_mc_batch_dev(tvc_ctx* ctx, int dst_slot, int n, const tvc_pu* pus_dev);

/* Prediction + distortion of candidate motions (merge / AMVP candidate evaluation).  For every entry: the luma
 * prediction motionCompensation would write (uni-prediction: clipped pels; both lists: 14-bit intermediates + TComYuv::addAvg;
 * MVs already clipped by TComDataCU::clipMv, identical-motion candidates already reduced to list 0 as xCheckIdenticalMotion does,
 * TComPrediction.cpp:392-552) and its distortion against the ORIGINAL block of cur_slot at the same rectangle:
 *   kind TVC_DIST_HADS / TVC_DIST_SAD = TEncSearch::xGetInterPredictionError (TEncSearch.cpp:3059-3081: setDistParam with
 *   HadamardME on / off, iSubShift 0) as xMergeEstimation calls it per merge candidate (:3096-3149);
 *   kind TVC_DIST_SAD                 = the xPredInterLumaBlk + getDistPart(DF_SAD) of TEncSearch::xGetTemplateCost (:4057-4118)
 *                                       per AMVP candidate.
 * The candidate lists, the bit counts and the comparisons (merge index bits, m_auiMVPIdxCost, strict "<" in candidate order) stay
 * host work.  dist[i] is the reference's UInt (already >> bitIncrement).                                                        */
int tvc_pred_cost_batch(tvc_ctx* ctx, int cur_slot, int kind, int n, const tvc_pu* pus, uint32_t* dist);
int tvc_pred_cost_batch_dev(tvc_ctx* ctx, int cur_slot, int kind, int n, const tvc_pu* pus_dev, uint32_t* dist_dev);

/* The look-up form of the same evaluation (what makes it pay inside the CU loop, where the reference asks one PU at a time): for
 * one (CTU, reference, clipped MV) the distortion of EVERY PU of the CTU with that motion, as three inclusive 2-D prefix sums over
 * the CTU: SAD of the 4x4 blocks (17x17), xCalcHADs4x4 of the 4x4 tiles (17x17), xCalcHADs8x8 of the 8x8 tiles (9x9); entry (r, c)
 * = sum over block / tile rows < r and columns < c, row 0 / column 0 are zero.  For a PU at (px, py, w, h) inside the CTU:
 *   SAD  = S(I_sad4; px/4, py/4, w/4, h/4) >> bitIncrement                 (xGetSAD*, iSubShift 0)
 *   HADs = S(I_had8; px/8, py/8, w/8, h/8) >> bitIncrement when w and h are multiples of 8, else S(I_had4; ...) >> bitIncrement
 * with S(I; x, y, w, h) = I[y+h][x+w] - I[y][x+w] - I[y+h][x] + I[y][x].  Uni-prediction (clipped pels); the host keeps the grids
 * of a picture in a table keyed by (CTU, reference, MV) and serves xGetTemplateCost / xMergeEstimation from it.                 */
typedef struct {
  int32_t ref_slot;
  int32_t x0, y0;              /* CTU origin in the picture (multiples of 64)                                      */
  int32_t mvx, mvy;            /* quarter-pel, after TComDataCU::clipMv                                            */
} tvc_grid_job;
#define TVC_GRID_WORDS (289 + 289 + 81)     /* I_sad4[17][17], I_had4[17][17], I_had8[9][9]                        */
int tvc_ctu_cost_grids(tvc_ctx* ctx, int cur_slot, int n, const tvc_grid_job* jobs, uint32_t* grids /* n * TVC_GRID_WORDS */);
int tvc_ctu_cost_grids_dev(tvc_ctx* ctx, int cur_slot, int n, const tvc_grid_job* jobs_dev, uint32_t* grids_dev);

/* One PU, one reference list, into caller buffers: TComPrediction::xPredInterUni = xPredInterLumaBlk +
 * xPredInterChromaBlk (TComPrediction.cpp:483-490, 554-645).  (x, y, w, h) luma rectangle, MV already
 * clipped, dst_* point at the PU's first sample inside a TComYuv.  bi != 0 keeps the 14-bit intermediate
 * that TComYuv::addAvg averages later (xPredInterBi, :492-552).                                  */
int tvc_mc_block(tvc_ctx* ctx, int ref_slot, int x, int y, int w, int h, int mvx, int mvy, int bi,
                 int16_t* dst_y, int stride_y, int16_t* dst_u, int16_t* dst_v, int stride_c);

/* ---------------------------------------------------------------------------------- integer ME
 * (1) Frame pre-pass: SAD tables.  For every CTU of cur_slot and every reference in ref_slots,
 * all 129x129 integer candidates around a per-CTU window centre, at 4x4-block granularity, even
 * and odd rows kept apart (the FEN sub-sampled SAD of PUs taller than 8 uses even rows only,
 * TEncSearch.cpp:324-330; PU rows are CTU-aligned multiples of 4).  Any PU's SAD at any candidate
 * is a sum of its blocks' entries, so decisions are those of xGetSAD* (TComRdCost.cpp:518-989).
 * 8-bit pictures only (u8 SIMD path).
 *
 * Table layout (uint16): T[ref][ctu][cand 129*129][by 16][q 4][parity 2][bx 4]; cand = (dy+R)*129+(dx+R),
 * R = TVC_ME_RANGE, dx/dy relative to the CTU's centre; 1 KB per candidate, 17.04 MB per (ref, CTU).  */
#define TVC_ME_RANGE 64
#define TVC_ME_CAND  (2 * TVC_ME_RANGE + 1)

typedef struct { int32_t cx, cy; } tvc_me_center;   /* integer-pel window centre of a CTU       */

/* centres: num_refs * num_ctus entries (ref-major) or NULL for all-zero.  Centres are clamped so
 * that every read stays inside the padded reference plane.                                     */
int tvc_me_prepass(tvc_ctx* ctx, int cur_slot, int num_refs, const int* ref_slots,
                   const tvc_me_center* centers);
/* allocate the SAD tables for num_refs references ahead of time (17.04 MB per CTU and reference; tvc_me_prepass grows
 * them on demand otherwise, and re-allocating tens of GB costs 0.2-0.6 s each time the reference count of a GOP grows) */
int tvc_me_reserve(tvc_ctx* ctx, int num_refs);
/* 1 when the census entry points below (tvc_me_frame, tvc_me_ctu with cfg->use_tables) read SAD tables from HBM and therefore
 * need tvc_me_prepass for the picture first (TVC_ME_FUSED=0, the round-1 form); 0 when they compute every SAD on demand from a
 * search window staged in shared memory (default; k_me_group) and no table is ever written.                              */
int tvc_me_uses_tables(tvc_ctx* ctx);
/* choose the form per context: 1 group search, 0 SAD tables, -1 back to the environment's choice */
int tvc_me_set_fused(tvc_ctx* ctx, int on);
/* bytes of table storage tvc_me_prepass needs for num_refs references (allocated lazily)       */
size_t tvc_me_table_bytes(tvc_ctx* ctx, int num_refs);
/* device pointer to the tables of the last pre-pass and the (clamped) centres actually used    */
int tvc_me_tables_dev(tvc_ctx* ctx, void** tables, tvc_me_center** centers_dev);
/* read one PU's SAD (already <<sub_shift, >>bitIncrement) at `n` candidates from the tables;
 * test / host-shim access path (TComRdCost DistFunc semantics for integer ME).                 */
int tvc_me_table_lookup(tvc_ctx* ctx, int ref_index, int pu_x, int pu_y, int pu_w, int pu_h, int fen,
                        int n, const int16_t* cand_xy /* 2n: integer mv x,y */, uint32_t* out);

/* (2) Search.  Replaces TEncSearch::xMotionEstimation's integer stage: xSetSearchRange is done by
 * the caller (it needs TComDataCU::clipMv), xPatternSearch (TEncSearch.cpp:4227-4283) or
 * xPatternSearchFast -> xTZSearch (:4285-4474) run on the device with the reference's visiting
 * order and strict '<' tie rule.  SADs come from the pre-pass tables when the candidate is
 * covered (use_tables), else from the pictures.                                               */
enum { TVC_ME_FULL = 0, TVC_ME_TZ = 1 };

typedef struct {
  int32_t ref_index;           /* index into the pre-pass ref_slots (tables) ...                */
  int32_t ref_slot;            /* ... and the slot itself (on-demand SAD path)                  */
  int32_t x, y, w, h;          /* luma PU rectangle                                             */
  int32_t mode;                /* TVC_ME_FULL / TVC_ME_TZ                                       */
  int32_t fen;                 /* getUseFastEnc(): sub-sample rows when h > 8                   */
  int32_t search_range;        /* m_iSearchRange (TZ distance bound)                            */
  int32_t lx, ty, rx, by;      /* window from xSetSearchRange, integer pels                     */
  int32_t predx, predy;        /* rate predictor (quarter pels), TComRdCost::setPredictor       */
  int32_t startx, starty;      /* TZ start point, integer pels (clipMv(pred) >> 2)              */
  uint32_t lambda_cost;        /* m_uiCost after getMotionCost(1,0): floor(65536*sqrt(lambda))  */
} tvc_me_job;

typedef struct {
  int32_t mvx, mvy;            /* best integer MV                                               */
  uint32_t sad;                /* ruiSAD: best cost minus its rate term                         */
  uint32_t n_sads;             /* candidates evaluated                                          */
} tvc_me_result;

int tvc_me_search_batch(tvc_ctx* ctx, int cur_slot, int use_tables, int n, const tvc_me_job* jobs,
                        tvc_me_result* out);
int tvc_me_search_batch_dev(tvc_ctx* ctx, int cur_slot, int use_tables, int n, const tvc_me_job* jobs_dev,
                            tvc_me_result* out_dev);

/* ---------------------------------------------------------------------------------- fractional ME
 * Replaces TEncSearch::xPatternSearchFracDIF = xExtDIFUpSamplingH + xPatternRefinement(2) +
 * xExtDIFUpSamplingQ + xPatternRefinement(1) (TEncSearch.cpp:4476-4514, 5982-6175, 711-760):
 * 8-tap interpolation of the half/quarter planes around the integer MV and 9+9 Hadamard-SATD
 * (or SAD) evaluations with the MV rate at cost scale 1 then 0.                               */
typedef struct {
  int32_t ref_slot;
  int32_t x, y, w, h;
  int32_t imvx, imvy;          /* integer MV from the integer stage                             */
  int32_t predx, predy;        /* rate predictor, quarter pels                                  */
  uint32_t lambda_cost;
  int32_t hadamard;            /* getUseHADME()                                                 */
} tvc_frac_job;

typedef struct {
  int32_t halfx, halfy;        /* rcMvHalf in {-1,0,1}                                          */
  int32_t qtrx, qtry;          /* rcMvQter in {-1,0,1}                                          */
  uint32_t cost_half, cost;    /* ruiCost after each refinement                                 */
} tvc_frac_result;

int tvc_me_frac_batch(tvc_ctx* ctx, int cur_slot, int n, const tvc_frac_job* jobs, tvc_frac_result* out);
int tvc_me_frac_batch_dev(tvc_ctx* ctx, int cur_slot, int n, const tvc_frac_job* jobs_dev, tvc_frac_result* out_dev);

/* Bi-prediction refinement of one PU: TEncSearch::xMotionEstimation with bBi (TEncSearch.cpp:4120-4207).  The search target is
 * not the picture but the block 2 * org - pred(other list) the host built (TComYuv::removeHighFreq, TComYuv.cpp:583-633; values
 * -255 .. 510 for 8-bit content): it is copied into `target_slot` at the PU's position, then xPatternSearch (exhaustive raster of
 * the +-bipredSearchRange window, :4227-4283; job->mode must be TVC_ME_FULL) and xPatternSearchFracDIF (:4476-4514) run against
 * job->ref_slot in ONE call (one copy up, three kernels, one copy back).  Synchronous.                                          */
int tvc_me_bipred(tvc_ctx* ctx, int target_slot, const int16_t* target, int target_stride, const tvc_me_job* job, int hadamard,
                  tvc_me_result* int_out, tvc_frac_result* frac_out);

/* ---------------------------------------------------------------------------------- frame-level ME pre-pass
 * The TEncCu frame pre-pass named by the north star: for one picture and up to 8 references it
 * runs -- for EVERY PU of the HM partition census of every CTU
 * (593 PU instances per CTU = 13 part shapes x 21 CUs at depths 0-2 + 5 x 64 CUs at
 * depth 3; census order below) and every reference -- xMotionEstimation's integer stage
 * (xSetSearchRange + xTZSearch, TEncSearch.cpp:4209-4225, 4302-4474) and fractional stage
 * (xPatternSearchFracDIF, :4476-4514), all on the device.  The AMVP predictor is a sequential
 * output of the CU loop, so the pre-pass takes one predictor GUESS per
 * (reference, CTU), in quarter pels; it is used as rate predictor, search-window centre (after
 * TComDataCU::clipMv with the PU's own CU origin, TComDataCU.cpp:3505-3517) and TZ start of every
 * PU of that CTU, and (clipped with the CTU origin, >> 2) as the centre of the search window the
 * group kernel stages (default form: every SAD is computed on demand from that window, nothing is
 * written to HBM but the results; with TVC_ME_FUSED=0 / tvc_me_set_fused(ctx, 0) the round-1 form
 * runs tvc_me_prepass first and the searches read its SAD tables).  A host whose real predictor
 * differs asks again with tvc_me_ctu (one CTU, one reference) or tvc_me_search_batch (single PUs).
 *
 * Census order inside a CTU (index 0..592): depth 0,1,2,3; CUs of a depth in raster order; parts of
 * a CU: 2Nx2N, 2NxN[0,1], Nx2N[0,1], then (CU >= 16 only) 2NxnU[0,1], 2NxnD[0,1], nLx2N[0,1],
 * nRx2N[0,1].  tvc_me_census writes the rectangles.  Result index = (ref * num_ctus + ctu) * 593 + k.
 * PUs that do not lie inside the picture (bottom/right partial CTUs) get n_sads == 0.           */
#define TVC_ME_CENSUS 593
typedef struct { int16_t x, y, w, h, cu_x, cu_y; } tvc_census_pu;   /* relative to the CTU origin    */
int tvc_me_census(tvc_census_pu* out /* TVC_ME_CENSUS entries */);

typedef struct {
  int32_t search_range;        /* m_iSearchRange (64 in every cfg)                               */
  int32_t fen;                 /* FastEncoderDecision                                            */
  int32_t hadamard;            /* HadamardME                                                     */
  int32_t use_tables;          /* 1: integer search reads the SAD tables; 0: on-demand SADs only  */
  int32_t do_frac;             /* 0: integer stage only                                          */
  uint32_t lambda_cost;        /* floor(65536*sqrt(lambda))                                      */
} tvc_me_frame_cfg;

/* pred_qpel: num_refs * num_ctus predictor guesses (quarter pels, ref-major) or NULL for zero.
 * int_out / frac_out: host arrays of num_refs * num_ctus * TVC_ME_CENSUS entries (either may be NULL). */
int tvc_me_frame(tvc_ctx* ctx, int cur_slot, int num_refs, const int* ref_slots, const tvc_me_center* pred_qpel,
                 const tvc_me_frame_cfg* cfg, tvc_me_result* int_out, tvc_frac_result* frac_out);
/* the same, both stages, with each job's result in 16 bytes instead of 40 (what TEncSearch::xMotionEstimation consumes: the integer
 * vector with its ruiSAD, the two refinement offsets and the final ruiCost).  A census PU outside the picture has sad == cost ==
 * 0xFFFFFFFF.  cfg->do_frac must be set.  At 1080p x 4 references: 19 MB back instead of 48 MB.                                   */
typedef struct { int16_t mvx, mvy; int8_t halfx, halfy, qtrx, qtry; uint32_t sad; uint32_t cost; } tvc_me_packed;
int tvc_me_frame_packed(tvc_ctx* ctx, int cur_slot, int num_refs, const int* ref_slots, const tvc_me_center* pred_qpel,
                        const tvc_me_frame_cfg* cfg, tvc_me_packed* out);
/* asynchronous, results stay on the device (pred_qpel is still a HOST array: it is tiny)         */
int tvc_me_frame_dev(tvc_ctx* ctx, int cur_slot, int num_refs, const int* ref_slots, const tvc_me_center* pred_qpel,
                     const tvc_me_frame_cfg* cfg, tvc_me_result** int_dev, tvc_frac_result** frac_dev);

/* The same two stages for ONE (CTU, reference) group with an explicit predictor: what a host does when its CU loop
 * reaches a CTU and learns the real AMVP predictor (the frame pre-pass can only guess it).  Measured on HM's own
 * 1080p LDP runs, 98.6 % of the integer searches of a (picture, CTU, reference) group use the group's first
 * predictor, so one call serves the whole group's TEncSearch::xMotionEstimation calls by look-up.  ref_index < 0 or
 * cfg->use_tables == 0: SADs straight from the pictures.  Synchronous; int_out / frac_out: TVC_ME_CENSUS entries. */
int tvc_me_ctu(tvc_ctx* ctx, int cur_slot, int ref_index, int ref_slot, int ctu, tvc_me_center pred_qpel,
               const tvc_me_frame_cfg* cfg, tvc_me_result* int_out, tvc_frac_result* frac_out);

/* The same group, asynchronously: queued on a side stream behind everything the context has queued so far, result kept under
 * `ticket` (0 .. TVC_ME_CTU_TICKETS - 1) until tvc_me_ctu_fetch waits for it and copies it out.  A host whose CU loop just learned
 * the first predictor of (CTU k, reference r) asks for (CTU k + 1, r) with the same predictor while it codes CTU k: in HM's own
 * 1080p runs the first predictor of a CTU equals its left neighbour's in 96.6 % of the groups (the guess of the frame pre-pass, the
 * previous picture's predictor, holds in 40-75 %), so the next CTU's 593 results are already there when the CU loop arrives.  Re-using
 * a ticket whose result was never fetched drops that result.                                                                     */
#define TVC_ME_CTU_TICKETS 8
int tvc_me_ctu_async(tvc_ctx* ctx, int ticket, int cur_slot, int ref_index, int ref_slot, int ctu, tvc_me_center pred_qpel,
                     const tvc_me_frame_cfg* cfg);
int tvc_me_ctu_fetch(tvc_ctx* ctx, int ticket, tvc_me_result* int_out, tvc_frac_result* frac_out);

/* work counters of the last tvc_me_frame[_dev] call (for roofline accounting).  Default (group search): stats[0] = candidate SADs
 * the searches evaluated (the reference's own count: sum of n_sads), stats[1] = sample differences of those (w x (h >> iSubShift)
 * per candidate), stats[2] = census jobs whose fractional search was served at CU level (k_me_frac_cu).  TVC_ME_FUSED=0 (SAD tables): stats[0] = 16-byte table granules the reference-visible candidates of
 * k_me_search required, stats[1] = candidates served by the shared raster stage, stats[2] = raster candidates walked by
 * k_me_raster.  Synchronises.                                                                                                */
int tvc_me_frame_stats(tvc_ctx* ctx, uint64_t stats[3]);

/* ---------------------------------------------------------------------------------- transform / quant
 * Replaces TComTrQuant::transformNxN = xT (-> xTrMxN -> partialButterfly4/8/16/32 /
 * fastForwardDst) or xTransformSkip, then xQuant's non-RDOQ branch with signBitHidingHDQ; and
 * invtransformNxN = xDeQuant + xIT / xITransformSkip (TComTrQuant.cpp:417-972, 977-1355,
 * 1373-1704).  Flat scaling lists (ScalingList 0 in every cfg).  RDOQ (xRateDistOptQuant) stays on
 * the host: tvc_fwd_transform_batch returns the Int coefficients it consumes.                  */
enum {
  TVC_TU_DST = 1,              /* 4x4 intra luma: uiMode != REG_DCT                              */
  TVC_TU_SKIP = 2,             /* transform skip (4x4)                                           */
  TVC_TU_BYPASS = 4            /* cu_transquant_bypass: copy                                     */
};

typedef struct {
  int32_t plane;               /* TVC_PLANE_*                                                    */
  int32_t x, y;                /* TU origin in samples of that plane                             */
  int32_t log2_size;           /* 2..5                                                           */
  int32_t flags;               /* TVC_TU_*                                                       */
  int32_t scan_idx;            /* 0 diag, 1 hor, 2 ver (sign-data hiding walks the coding scan)  */
  int32_t qp_per, qp_rem;      /* m_cQP after setQPforQuant                                      */
  int32_t base_per;            /* per of the slice base QP (ADAPTIVE_QP_SELECTION)               */
  int32_t coef_offset;         /* element offset of this TU's w*h raster in the coefficient buffer */
} tvc_tu;

typedef struct {
  int32_t is_intra_slice;      /* rounding 171 (I) / 85 (P,B) >> 9                               */
  int32_t sign_hide;           /* PPS sign_data_hiding                                           */
  int32_t use_arl;             /* m_bUseAdaptQpSelect: also write ARL coefficients               */
} tvc_quant_cfg;

/* Host-pointer entry points copy directly from / into page-locked caller buffers (cudaHostAlloc /
 * cudaHostRegister) and stage pageable ones.  Elements of a coefficient / level buffer that no TU of the
 * list covers are unspecified on return.
 * residual plane (resi_slot) -> forward transform -> Int coefficients (TCoeff raster per TU)   */
int tvc_fwd_transform_batch(tvc_ctx* ctx, int resi_slot, int n, const tvc_tu* tus, int32_t* coef, size_t coef_elems);
/* residual -> transform -> quant (+sign hiding): levels, optional ARL, per-TU uiAbsSum          */
int tvc_fwd_tq_batch(tvc_ctx* ctx, int resi_slot, int n, const tvc_tu* tus, const tvc_quant_cfg* qc,
                     int32_t* levels, int32_t* arl /* or NULL */, size_t coef_elems, uint32_t* abs_sum);
/* levels -> dequant -> inverse transform -> residual written into resi_slot; if pred_slot >= 0
 * also recon_slot = Clip(pred + resi) (TComYuv::addClip)                                        */
int tvc_inv_tq_batch(tvc_ctx* ctx, int resi_slot, int pred_slot, int recon_slot, int n, const tvc_tu* tus,
                     const int32_t* levels, size_t coef_elems);
/* device-resident variants (tus_dev / coefficient buffers on the device), asynchronous.  The TU
 * list must be grouped by ascending log2_size (the host-pointer entry points require the same);
 * counts[4] = number of 4x4, 8x8, 16x16, 32x32 TUs (host array).                               */
int tvc_fwd_tq_batch_dev(tvc_ctx* ctx, int resi_slot, int n, const tvc_tu* tus_dev, const int32_t* counts,
                         const tvc_quant_cfg* qc, int32_t* levels_dev, int32_t* arl_dev, uint32_t* abs_sum_dev);
int tvc_inv_tq_batch_dev(tvc_ctx* ctx, int resi_slot, int pred_slot, int recon_slot, int n, const tvc_tu* tus_dev,
                         const int32_t* counts, const int32_t* levels_dev);

/* drop-ins on host blocks for the leaf members (one TU per call)                                */
int tvc_xT(tvc_ctx* ctx, int use_dst, const int16_t* resi, int stride, int32_t* coef, int w, int h);
int tvc_xIT(tvc_ctx* ctx, int use_dst, const int32_t* coef, int16_t* resi, int stride, int w, int h);
int tvc_xDeQuant(tvc_ctx* ctx, const int32_t* qcoef, int32_t* coef, int w, int h, int per, int rem);

/* ---------------------------------------------------------------------------------- RDOQ
 * Replaces TComTrQuant::xRateDistOptQuant (TComTrQuant.cpp:1719-2305) with its helpers xGetCodedLevel,
 * xGetICRateCost, xGetICRate, xGetRateLast, xGetRateSigCoef, xGetRateSigCoeffGroup (:2446-2698),
 * getSigCtxInc, calcPatternSigCtx, getSigCoeffGroupCtxInc (:2315-2428, 2707-2743) and the flat-list
 * error scale of setErrScaleCoeff (:2794-2817), batched over TUs.  The Lagrangian costs are IEEE
 * doubles evaluated in the reference's order without FMA contraction, so levels, uiAbsSum and ARL
 * coefficients are the reference's.  The CABAC bit estimates are an INPUT: the host's
 * TEncEntropy::estimateBit fills them from its live context models exactly as today.           */
typedef struct {                 /* estBitsSbacStruct, TComTrQuant.h:59-72: same members, same order, so
                                    HM passes m_pcEstBitsSbac with a cast                            */
  int32_t sig_cg[2][2];          /* significantCoeffGroupBits */
  int32_t sig[42][2];            /* significantBits           */
  int32_t last_x[32];            /* lastXBits                 */
  int32_t last_y[32];            /* lastYBits                 */
  int32_t greater_one[24][2];    /* m_greaterOneBits          */
  int32_t level_abs[6][2];       /* m_levelAbsBits            */
  int32_t block_cbp[15][2];      /* blockCbpBits              */
  int32_t block_root_cbp[4][2];  /* blockRootCbpBits          */
  int32_t scan_zigzag[2], scan_non_zigzag[2];
} tvc_est_bits;

typedef struct {
  int32_t log2_size;           /* 2..5                                                           */
  int32_t is_luma;             /* eTType == TEXT_LUMA                                            */
  int32_t scan_idx;            /* 0 diag (also for SCAN_ZIGZAG), 1 hor, 2 ver; 1/2 only for log2 <= 3 */
  int32_t qp_per, qp_rem;      /* m_cQP after setQPforQuant                                      */
  int32_t cbf_ctx;             /* < 0: inter luma TU at transform depth 0 (blockRootCbpBits[0]);
                                  else the blockCbpBits row: luma (trDepth==0 ? 1 : 0), chroma 5 + trDepth */
  int32_t est_index;           /* which tvc_est_bits of the call's table array                   */
  int32_t coef_offset;         /* element offset of this TU's raster in coef / levels / arl      */
  double lambda;               /* m_dLambda after selectLambda                                   */
} tvc_rdoq_tu;

/* coef: xT output (tvc_fwd_transform_batch layout); levels: signed TCoeff out; arl: written only when
 * qc->use_arl (else untouched, may be NULL); abs_sum[i] = uiAbsSum of TU i.  qc->is_intra_slice is unused. */
int tvc_rdoq_batch(tvc_ctx* ctx, int n, const tvc_rdoq_tu* tus, int n_est, const tvc_est_bits* est,
                   const tvc_quant_cfg* qc, const int32_t* coef, int32_t* levels, int32_t* arl, size_t coef_elems,
                   uint32_t* abs_sum);
/* device-resident, asynchronous on the context stream */
int tvc_rdoq_batch_dev(tvc_ctx* ctx, int n, const tvc_rdoq_tu* tus_dev, int n_est, const tvc_est_bits* est_dev,
                       const tvc_quant_cfg* qc, const int32_t* coef_dev, int32_t* levels_dev, int32_t* arl_dev,
                       size_t coef_elems, uint32_t* abs_sum_dev);
/* residual plane -> forward transform (device coefficients, no host copy) for tvc_rdoq_batch_dev */
int tvc_fwd_transform_batch_dev(tvc_ctx* ctx, int resi_slot, int n, const tvc_tu* tus_dev, const int32_t* counts,
                                int32_t* coef_dev);
/* transformNxN with RDOQ on: residual plane -> xT -> xRateDistOptQuant, coefficients stay on the device.
 * tus[i] and rdoq_tus[i] describe the same TU (same log2_size and coef_offset); host pointers.  */
int tvc_fwd_rdoq_batch(tvc_ctx* ctx, int resi_slot, int n, const tvc_tu* tus, const tvc_rdoq_tu* rdoq_tus, int n_est,
                       const tvc_est_bits* est, const tvc_quant_cfg* qc, int32_t* levels, int32_t* arl, size_t coef_elems,
                       uint32_t* abs_sum);
/* the whole residual round trip of a TU list in one call: transformNxN with RDOQ on, then invtransformNxN of the same levels
 * and the reconstruction (what TEncSearch::xEstimateResidualQT / xEncodeResidualQT do per TU, batched): the levels and
 * uiAbsSum go to the host (CABAC needs them), the inverse path reads the levels where RDOQ left them on the device.
 * inv_resi_slot receives the reconstructed residual, recon_slot = Clip(pred_slot + residual).                     */
int tvc_fwd_rdoq_recon_batch(tvc_ctx* ctx, int resi_slot, int inv_resi_slot, int pred_slot, int recon_slot, int n, const tvc_tu* tus,
                             const tvc_rdoq_tu* rdoq_tus, int n_est, const tvc_est_bits* est, const tvc_quant_cfg* qc, int32_t* levels,
                             size_t coef_elems, uint32_t* abs_sum);
/* the same with the levels returned as int16 (what CABAC codes fits 16 bits: TComTrQuant.cpp:1298 clips to [-32768, 32767]): half
 * the bytes of a picture's largest device-to-host transfer.  A level beyond 16 bits is an error (TVC_ERR_ARG); no ARL output. */
int tvc_fwd_rdoq_recon_batch16(tvc_ctx* ctx, int resi_slot, int inv_resi_slot, int pred_slot, int recon_slot, int n, const tvc_tu* tus,
                               const tvc_rdoq_tu* rdoq_tus, int n_est, const tvc_est_bits* est, const tvc_quant_cfg* qc,
                               int16_t* levels16, size_t coef_elems, uint32_t* abs_sum);
/* drop-in for one xRateDistOptQuant call on host blocks (w x w, raster)                          */
int tvc_xRateDistOptQuant(tvc_ctx* ctx, const int32_t* coef, int32_t* qcoef, int32_t* arl, int w, int is_luma,
                          int scan_idx, int qp_per, int qp_rem, int cbf_ctx, int sign_hide, int use_arl,
                          double lambda, const tvc_est_bits* est, uint32_t* abs_sum);

/* ---------------------------------------------------------------------------------- deblocking filter
 * SURVEY.md 8(f)-1.  Replaces the sample work of TComLoopFilter::loopFilterPic (TComLoopFilter.cpp:153-191):
 * xEdgeFilterLuma / xEdgeFilterChroma (:571-797) with xPelFilterLuma / xPelFilterChroma / xUseStrongFiltering /
 * xCalcDP / xCalcDQ (:799-921) and the tc / beta tables (:56-64), on a device-resident picture, in place: all vertical
 * edges, then all horizontal edges (the reference's two passes; inside a pass edges lie 8 pels apart and touch at most
 * 3 + 4 pels on either side, so their order does not matter).  The boundary strengths stay host work (they come out of
 * the CU tree: xSetLoopfilterParam, xSetEdgefilterTU/PU, xGetBoundaryStrengthSingle, :266-569): the host hands over
 * one record per 4-pel edge unit on the 8x8 grid.  Chroma edges (4:2:0) are derived from the same records: units with
 * bs > 1 on the 16-pel luma grid, QP through g_aucChromaScale (TComRom.cpp:380-386).                               */
typedef struct {
  uint8_t bs;                  /* m_aapucBS at the unit: 0 none, 1, 2                                              */
  uint8_t qp;                  /* (QP_P + QP_Q + 1) >> 1 of the two CUs (:627)                                     */
  uint8_t flags;               /* bit 0: part P not filtered, bit 1: part Q not filtered (IPCM / lossless, :651-658) */
  uint8_t reserved;
} tvc_dbk_unit;
/* ver: vertical edges, entry [(y >> 2) * ((width + 7) >> 3) + (x >> 3)] = edge at luma column x (multiple of 8), rows
 * y .. y+3; hor: horizontal edges, entry [(y >> 3) * ((width + 3) >> 2) + (x >> 2)] = edge at luma row y (multiple of
 * 8), columns x .. x+3.  Host arrays; either may be NULL (pass skipped).                                          */
int tvc_deblock_pic(tvc_ctx* ctx, int slot, const tvc_dbk_unit* ver, const tvc_dbk_unit* hor, int beta_offset_div2,
                    int tc_offset_div2);

/* ---------------------------------------------------------------------------------- sample adaptive offset (apply)
 * SURVEY.md 8(f)-1, second half.  Replaces the sample work of TComSampleAdaptiveOffset::processSaoUnitAll ->
 * processSaoCu -> processSaoCuOrg (TComSampleAdaptiveOffset.cpp:781-1003, 1072-1236) for one colour component of a
 * single-slice picture.  The reference filters in place CTU by CTU and keeps the unfiltered left column / upper row
 * in line buffers (m_pTmpL1/2, m_pTmpU1/2), i.e. every edge class is taken on the DEBLOCKED picture: here the
 * deblocked picture is read from src_slot and the result written to dst_slot (a different slot), one thread per
 * sample.  The SAO decision (offsets, types, merges) stays host work: one record per CTU, merges resolved.        */
typedef struct {
  int16_t type;                /* -1 off, 0..3 SAO_EO_0..3 (0 deg, 90 deg, 135 deg, 45 deg), 4 SAO_BO                */
  int16_t eo[5];               /* m_iOffsetEo[edgeType 0..4] (already << m_uiSaoBitIncrease); eo[2] is 0          */
  int16_t bo[32];              /* offset of band k = sample >> (bit_depth - 5) (the reference's offset[k + 1])     */
} tvc_sao_unit;
/* units: one per CTU in raster order (host array); samples of CTUs with type < 0 are copied                        */
int tvc_sao_plane(tvc_ctx* ctx, int src_slot, int dst_slot, int plane, const tvc_sao_unit* units);

/* ---------------------------------------------------------------------------------- picture hash / PSNR sums
 * what the decoded-picture-hash SEI and the PSNR report need of a reconstruction, taken where the picture
 * already lies.  tvc_pic_hash = calcMD5 / calcCRC / calcChecksum (TLibCommon/TComPicYuvMD5.cpp:119-200; call sites
 * TLibEncoder/TEncGOP.cpp:1150-1172, TLibDecoder/TDecGop.cpp:340-370): digest is the reference's unsigned char [3][16] (Y, Cb, Cr;
 * MD5 16 bytes, CRC 2, checksum 4, the rest zero).  method = the SEI's hash type = --SEIpictureDigest.
 * tvc_pic_ssd = the three UInt64 sums of squared differences of TEncGOP::xCalculateAddPSNR (TLibEncoder/TEncGOP.cpp:1582-1641)
 * between two slots (original, reconstruction); the log10 stays with the caller.                                              */
enum { TVC_HASH_MD5 = 1, TVC_HASH_CRC = 2, TVC_HASH_CHECKSUM = 3 };
int tvc_pic_hash(tvc_ctx* ctx, int slot, int method, uint8_t* digest /* [3][16] */);
int tvc_pic_ssd(tvc_ctx* ctx, int slot_a, int slot_b, uint64_t* ssd /* [3] */);

/* ---------------------------------------------------------------------------------- intra 35-mode rough search
 * SURVEY.md 8(f)-2.  Replaces the per-mode body of the rough search in TEncSearch::estIntraPredQT
 * (TLibEncoder/TEncSearch.cpp:2530-2537): TComPrediction::predIntraLumaAng (TLibCommon/TComPrediction.cpp:337-366 ->
 * xPredIntraPlanar :689-731, xPredIntraAng :186-335, predIntraGetPredValDC :127-165, xDCPredFiltering :1010-1031) on
 * the reference samples TComPattern::initAdiPattern prepared (TLibCommon/TComPattern.cpp:213-307, including its
 * [1 2 1] smoothing; which modes read the smoothed samples: getPredictorPtr :577-605) followed by TComRdCost::calcHAD
 * (TLibCommon/TComRdCost.cpp:404-447), for all 35 modes of an N x N luma PU at once.  The neighbour substitution
 * (fillReferenceSamples, TComPattern.cpp:368-552) depends on the CU tree and stays host work; so do the mode bits
 * (xModeBitsIntra: CABAC state) and the candidate list (xUpdateCandList): the host adds bits * sqrt(lambda) to the
 * returned SATDs in its own loop, in mode order, exactly as before.
 * Reference samples travel as ONE LINE of 4N+1 Pels per PU, in the order initAdiPattern walks them (:277-288):
 * left column from the bottom-left sample upwards (2N), the top-left corner, the row above from left to above-right
 * (2N) -- i.e. column 0 (bottom to top) and row 0 of the reference's (2N+1)x(2N+1) m_piYuvExt array, UNFILTERED.   */
typedef struct {
  int32_t log2_size;           /* 2..6: N = 4..64 (64: the 2Nx2N PU of a 64x64 CU)                                 */
  int32_t line_offset;         /* element offset of the PU's 4N+1 reference samples in `lines`                     */
  int32_t org_offset;          /* element offset of the PU's original block in `org`                               */
  int32_t org_stride;          /* elements                                                                         */
  int32_t above, left;         /* bAbove / bLeft as predIntraLumaAng receives them (initAdiPattern sets both)      */
} tvc_intra_job;
#define TVC_INTRA_MODES 35
/* sad[35 * i + mode] = calcHAD(org, prediction of `mode`) of job i (already >> bitIncrement); host arrays          */
int tvc_intra_rough_batch(tvc_ctx* ctx, int n, const tvc_intra_job* jobs, const int16_t* lines, size_t line_elems,
                          const int16_t* org, size_t org_elems, uint32_t* sad);
/* the same on device-resident arrays, asynchronous on the context stream; preds_dev (optional): the 35 predictions of
 * every job, job i mode m at element pred_offset[i] + m * N * N (pred_offset_dev: n element offsets, device)       */
int tvc_intra_rough_batch_dev(tvc_ctx* ctx, int n, const tvc_intra_job* jobs_dev, const int16_t* lines_dev,
                              const int16_t* org_dev, uint32_t* sad_dev, int16_t* preds_dev, const int64_t* pred_offset_dev);
/* drop-in for ONE estIntraPredQT rough search: 35 SATDs of one PU; preds (optional, host): 35 x N x N predictions,
 * mode-major = what predIntraLumaAng writes for each mode                                                          */
int tvc_intra_rough(tvc_ctx* ctx, int log2_size, const int16_t* line, const int16_t* org, int org_stride, int above,
                    int left, uint32_t sad[TVC_INTRA_MODES], int16_t* preds);

/* ---------------------------------------------------------------------------------- per-phase device timing
 * CUDA events recorded on the context stream around every kernel group, so that bench.py can report
 * each kernel's duration measured live inside the timed region (not under a profiler).         */
enum {
  TVC_PH_ME_TABLES = 0, TVC_PH_ME_SEARCH = 1, TVC_PH_ME_FRAC = 2, TVC_PH_MC = 3, TVC_PH_FWD_TQ = 4,
  TVC_PH_INV_TQ = 5, TVC_PH_OTHER = 6, TVC_PH_ME_RASTER = 7, TVC_PH_RDOQ = 8, TVC_PH_DEBLOCK = 9, TVC_PH_INTRA = 10, TVC_PH_COUNT = 11
};
int tvc_prof_enable(tvc_ctx* ctx, int on);
/* synchronises the stream, adds the elapsed time of every recorded pair to per-phase sums and returns
 * the sums (milliseconds) and the number of timed kernel groups per phase; reset != 0 clears them */
int tvc_prof_read(tvc_ctx* ctx, double* ms_sum, uint64_t* groups, int reset);

/* ---------------------------------------------------------------------------------- diagnostics
 * integer-pipe micro-benchmarks used for the ME roofline denominator; returns the
 * measured rate in giga-instructions/s for the whole GPU                                        */
enum { TVC_UB_VABSDIFF4 = 0, TVC_UB_IADD3 = 1, TVC_UB_IMAD = 2, TVC_UB_LDS128 = 3, TVC_UB_DP2A = 4,
       TVC_UB_HBM_WRITE = 5 /* pure 16-byte coalesced stores over 8 GiB: the result is GB/s, the write-only HBM roofline */ };
int tvc_ubench(tvc_ctx* ctx, int which, double* ginstr_per_s);

#ifdef __cplusplus
}
#endif
#endif /* THEVC_CUDA_H */
