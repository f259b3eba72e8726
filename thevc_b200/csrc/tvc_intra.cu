// tvc_intra.cu -- intra 35-mode rough search: predIntraLumaAng + calcHAD for all modes of a PU.
//
// Reference: TEncSearch::estIntraPredQT TLibEncoder/TEncSearch.cpp:2530-2537; TComPrediction::predIntraLumaAng
// TLibCommon/TComPrediction.cpp:337-366 (xPredIntraAng :186-335, xPredIntraPlanar :689-731, predIntraGetPredValDC
// :127-165, xDCPredFiltering :1010-1031); TComPattern::initAdiPattern smoothing TLibCommon/TComPattern.cpp:262-307,
// getPredictorPtr :577-605; TComRdCost::calcHAD TLibCommon/TComRdCost.cpp:404-447 (xCalcHADs8x8 :1776-1872, 4x4 :1689-1774).
//
// Mapping: one CTA per PU.  The 4N+1 reference samples (unfiltered + smoothed) and the original block sit in shared
// memory.  A work item is (mode, Hadamard tile); T = 8 (4 for 4x4 PUs) adjacent lanes own one tile, one tile row per
// lane: the lane evaluates its T prediction samples in closed form (no mode needs its neighbours' predictions), takes
// the horizontal butterflies in registers and the vertical ones with T-lane shuffles; tile sums go to one shared
// counter per mode.  The 2-D Hadamard sum of absolute values does not depend on butterfly order, only the per-tile
// rounding ((s+2)>>2 for 8x8, (s+1)>>1 for 4x4) and the final >> bitIncrement are the reference's.
#include "tvc_internal.cuh"

namespace tvc {

constexpr int kIntraThreads = 128;
constexpr int kIntraMaxN = 64;

constexpr int kIntraExtLen = 3 * kIntraMaxN + 2;
struct IntraSmem {
  int16_t line[2][4 * kIntraMaxN + 4];   // [0] as given, [1] smoothed; centre (corner) at index 2N
  int16_t org[kIntraMaxN * kIntraMaxN];
  // per angular mode (2..34) the reference's refMain array (xPredIntraAng :228-260) in the vertical family's frame: entry N + i is
  // main(i), i = -N .. 2N; negative i (negative angles only) are the projected side samples.  Built once per PU, so that a
  // prediction sample is two loads and one multiply-add instead of the projection arithmetic.
  int16_t ext[33][kIntraExtLen];
  uint32_t sad[TVC_INTRA_MODES];
  int dc;
};

__constant__ int8_t c_ang_table[9] = {0, 2, 5, 9, 13, 17, 21, 26, 32};
__constant__ int16_t c_inv_ang_table[9] = {0, 4096, 1638, 910, 630, 482, 390, 315, 256};
__constant__ uint8_t c_intra_filter_thr[5] = {10, 7, 1, 0, 10};     // TComPattern::m_aucIntraFilter

// everything of a mode that does not depend on the sample position
struct ModeParam {
  const int16_t* R;     // reference line the mode reads, pointing at the corner sample
  int kind;             // 0 planar, 1 DC, 2 angular
  int ver;              // angular: vertical family (modes 18..34); horizontal modes are the transpose
  int ang, inv;
};

__device__ __forceinline__ ModeParam mode_param(const int16_t* line_given, const int16_t* line_smoothed, int mode, int n, int log2n)
{
  ModeParam p;
  const int dh = abs(mode - 10), dv = abs(mode - 26);
  const bool filt = mode != 1 && min(dh, dv) > (int)c_intra_filter_thr[log2n - 2];
  p.R = (filt ? line_smoothed : line_given) + 2 * n;
  p.kind = mode < 2 ? mode : 2;
  p.ver = mode >= 18;
  const int idx = p.ver ? mode - 26 : -(mode - 10);
  const int a = (int)c_ang_table[abs(idx)];
  p.ang = idx < 0 ? -a : a;
  p.inv = (int)c_inv_ang_table[abs(idx)];
  return p;
}

// prediction sample (x, y) of an N x N block; R[+1+i] = row above, R[-1-j] = left column, R[0] = corner
__device__ __forceinline__ int pred_sample(const ModeParam& p, int n, int log2n, int x, int y, int dc, bool dc_edges, int max_pel)
{
  const int16_t* R = p.R;
  if (p.kind == 0)
    return ((n - 1 - x) * R[-1 - y] + (x + 1) * R[1 + n] + (n - 1 - y) * R[1 + x] + (y + 1) * R[-1 - n] + n) >> (log2n + 1);
  if (p.kind == 1) {
    if (dc_edges) {
      if (x == 0 && y == 0) return (R[1] + R[-1] + 2 * dc + 2) >> 2;
      if (y == 0) return (R[1 + x] + 3 * dc + 2) >> 2;
      if (x == 0) return (R[-1 - y] + 3 * dc + 2) >> 2;
    }
    return dc;
  }
  const int k = p.ver ? y : x, l = p.ver ? x : y;        // row / column in the vertical family's frame
  const int sm = p.ver ? 1 : -1;                         // main(i) = R[sm * i], side(j) = R[-sm * j]
  auto main_ref = [&](int i) -> int {
    if (i >= 0) return R[sm * i];
    return R[-sm * ((128 + (-i) * p.inv) >> 8)];         // projected side samples (invAngleSum walk)
  };
  if (p.ang == 0) {
    int v = R[sm * (l + 1)];
    if (l == 0) { v += (R[-sm * (k + 1)] - R[0]) >> 1; v = v < 0 ? 0 : (v > max_pel ? max_pel : v); }
    return v;
  }
  const int pos = (k + 1) * p.ang, di = pos >> 5, df = pos & 31, i0 = l + di + 1;
  if (df == 0) return main_ref(i0);
  return ((32 - df) * main_ref(i0) + df * main_ref(i0 + 1) + 16) >> 5;
}

// angular prediction sample from the extended line E (pointing at main(0)) of the mode
__device__ __forceinline__ int pred_sample_ext(const ModeParam& p, const int16_t* __restrict__ E, const int16_t* __restrict__ R, int x, int y,
                                               int max_pel)
{
  const int k = p.ver ? y : x, l = p.ver ? x : y;
  if (p.ang == 0) {
    int v = E[l + 1];
    if (l == 0) { const int sm = p.ver ? 1 : -1; v += (R[-sm * (k + 1)] - R[0]) >> 1; v = v < 0 ? 0 : (v > max_pel ? max_pel : v); }
    return v;
  }
  const int pos = (k + 1) * p.ang, di = pos >> 5, df = pos & 31, i0 = l + di + 1;
  return df ? ((32 - df) * E[i0] + df * E[i0 + 1] + 16) >> 5 : E[i0];
}

template <int T>
__device__ __forceinline__ uint32_t tile_had(int* d, int lane)
{
  // horizontal
#pragma unroll
  for (int s = 1; s < T; s <<= 1)
#pragma unroll
    for (int i = 0; i < T; i++)
      if (!(i & s)) { const int a = d[i], b = d[i | s]; d[i] = a + b; d[i | s] = a - b; }
  // vertical, across the T lanes of the tile
#pragma unroll
  for (int s = 1; s < T; s <<= 1) {
    const bool hi = lane & s;
#pragma unroll
    for (int i = 0; i < T; i++) {
      const int o = __shfl_xor_sync(0xffffffffu, d[i], s);
      d[i] = hi ? o - d[i] : d[i] + o;
    }
  }
  uint32_t sum = 0;
#pragma unroll
  for (int i = 0; i < T; i++) sum += (uint32_t)abs(d[i]);
#pragma unroll
  for (int s = 1; s < T; s <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
  return T == 8 ? (sum + 2) >> 2 : (sum + 1) >> 1;
}

template <int T>
__device__ __forceinline__ void rough_items(IntraSmem& S, int n, int log2n, bool above, bool left, int bd, int16_t* __restrict__ preds)
{
  const int tid = threadIdx.x, lane = tid & 31;
  const int row = lane % T, group = tid / T, groups = kIntraThreads / T;
  const int tiles_x = n / T, tiles = tiles_x * tiles_x, items = TVC_INTRA_MODES * tiles;
  const int max_pel = (1 << bd) - 1;
  const bool dc_edges = above && left;
  for (int base = 0; base < items; base += groups) {      // uniform trip count: every lane takes part in the shuffles
    const int item = base + group;
    const bool valid = item < items;
    const int it = valid ? item : 0;
    const int mode = it / tiles, tile = it - mode * tiles;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int y = ty * T + row, x0 = tx * T;
    const ModeParam mp = mode_param(S.line[0], S.line[1], mode, n, log2n);
    int d[T];
#pragma unroll
    const int16_t* E = S.ext[mode >= 2 ? mode - 2 : 0] + n;
    for (int i = 0; i < T; i++) {
      const int pv = mode >= 2 ? pred_sample_ext(mp, E, mp.R, x0 + i, y, max_pel) : pred_sample(mp, n, log2n, x0 + i, y, S.dc, dc_edges, max_pel);
      if (preds && valid) preds[(size_t)mode * n * n + y * n + x0 + i] = (int16_t)pv;
      d[i] = (int)S.org[y * n + x0 + i] - pv;
    }
    const uint32_t s = tile_had<T>(d, lane);
    if (valid && row == 0) atomicAdd(&S.sad[mode], s);
  }
}

__global__ void __launch_bounds__(kIntraThreads)
k_intra_rough(int n_jobs, const tvc_intra_job* __restrict__ jobs, const int16_t* __restrict__ lines, const int16_t* __restrict__ org,
              uint32_t* __restrict__ sad, int16_t* __restrict__ preds, const int64_t* __restrict__ pred_offset, int bd)
{
  __shared__ IntraSmem S;
  const int job = blockIdx.x, tid = threadIdx.x;
  if (job >= n_jobs) return;
  const tvc_intra_job j = jobs[job];
  if (j.log2_size <= 3) return;                    // served by k_intra_rough_small (warp per PU)
  const int log2n = j.log2_size, n = 1 << log2n, len = 4 * n + 1;
  const int16_t* ln = lines + j.line_offset;
  for (int i = tid; i < len; i += kIntraThreads) {
    const int c = ln[i];
    S.line[0][i] = (int16_t)c;
    // initAdiPattern :290-296: the two end samples are copied, the rest is (a + 2b + c + 2) >> 2
    S.line[1][i] = (i == 0 || i == len - 1) ? (int16_t)c : (int16_t)((ln[i - 1] + 2 * c + ln[i + 1] + 2) >> 2);
  }
  const int16_t* ob = org + j.org_offset;
  for (int i = tid; i < n * n; i += kIntraThreads) S.org[i] = ob[(i >> log2n) * j.org_stride + (i & (n - 1))];
  if (tid < TVC_INTRA_MODES) S.sad[tid] = 0;
  __syncthreads();
  for (int e = tid; e < 33 * (3 * n + 1); e += kIntraThreads) {
    const int m = e / (3 * n + 1), i = e - m * (3 * n + 1) - n;             // mode m + 2, main index i = -N .. 2N
    const ModeParam mp = mode_param(S.line[0], S.line[1], m + 2, n, log2n);
    const int sm = mp.ver ? 1 : -1;
    int v = 0;
    if (i >= 0) v = mp.R[sm * i];
    else if (mp.ang < 0 && i > ((n * mp.ang) >> 5)) v = mp.R[-sm * ((128 + (-i) * mp.inv) >> 8)];     // the entries the reference extends (:243-247)
    S.ext[m][i + n] = (int16_t)v;
  }
  if (tid < 32) {       // predIntraGetPredValDC :127-165 on the unfiltered line
    int sum = 0;
    const int16_t* R = S.line[0] + 2 * n;
    for (int i = tid; i < n; i += 32) sum += (j.above ? R[1 + i] : 0) + (j.left ? R[-1 - i] : 0);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
    if (tid == 0) {
      int dc;
      if (j.above && j.left) dc = (sum + n) / (2 * n);
      else if (j.above || j.left) dc = (sum + n / 2) / n;
      else dc = R[-1];
      S.dc = dc;
    }
  }
  __syncthreads();
  int16_t* pj = preds ? preds + pred_offset[job] : nullptr;
  if (n >= 8) rough_items<8>(S, n, log2n, j.above != 0, j.left != 0, bd, pj);
  else rough_items<4>(S, n, log2n, j.above != 0, j.left != 0, bd, pj);
  __syncthreads();
  if (tid < TVC_INTRA_MODES) sad[(size_t)job * TVC_INTRA_MODES + tid] = S.sad[tid] >> (bd - 8);
}

// ---- PUs of 4x4 and 8x8 (94 % of the intra PUs of a picture): one WARP per PU, kIntraSmallWarps PUs per CTA, no block-level barrier.
// A mode has exactly one Hadamard tile here, so the tile sums are stored, not accumulated.
constexpr int kIntraSmallWarps = 8;
struct IntraSmallSmem {
  int16_t line[2][36];
  int16_t org[64];
  uint32_t sad[TVC_INTRA_MODES + 1];
};

template <int T>
__device__ __forceinline__ void rough_small(IntraSmallSmem& S, int log2n, int dc, bool above, bool left, int bd, int16_t* __restrict__ preds)
{
  constexpr int n = T;
  const int lane = threadIdx.x & 31, row = lane % T, group = lane / T, groups = 32 / T;
  const int max_pel = (1 << bd) - 1;
  const bool dc_edges = above && left;
  for (int base = 0; base < TVC_INTRA_MODES; base += groups) {
    const int mode_raw = base + group;
    const bool valid = mode_raw < TVC_INTRA_MODES;
    const int mode = valid ? mode_raw : 0;
    const ModeParam mp = mode_param(S.line[0], S.line[1], mode, n, log2n);
    int d[T];
#pragma unroll
    for (int i = 0; i < T; i++) {
      const int pv = pred_sample(mp, n, log2n, i, row, dc, dc_edges, max_pel);
      if (preds && valid) preds[(size_t)mode * n * n + row * n + i] = (int16_t)pv;
      d[i] = (int)S.org[row * n + i] - pv;
    }
    const uint32_t s = tile_had<T>(d, lane);
    if (valid && row == 0) S.sad[mode] = s;
  }
}

__global__ void __launch_bounds__(kIntraSmallWarps * 32)
k_intra_rough_small(int n_jobs, const tvc_intra_job* __restrict__ jobs, const int16_t* __restrict__ lines, const int16_t* __restrict__ org,
                    uint32_t* __restrict__ sad, int16_t* __restrict__ preds, const int64_t* __restrict__ pred_offset, int bd)
{
  __shared__ IntraSmallSmem SS[kIntraSmallWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int job = blockIdx.x * kIntraSmallWarps + warp;
  if (job >= n_jobs) return;
  const tvc_intra_job j = jobs[job];
  if (j.log2_size > 3) return;                     // served by k_intra_rough (CTA per PU)
  IntraSmallSmem& S = SS[warp];
  const int log2n = j.log2_size, n = 1 << log2n, len = 4 * n + 1;
  const int16_t* ln = lines + j.line_offset;
  for (int i = lane; i < len; i += 32) {
    const int c = ln[i];
    S.line[0][i] = (int16_t)c;
    S.line[1][i] = (i == 0 || i == len - 1) ? (int16_t)c : (int16_t)((ln[i - 1] + 2 * c + ln[i + 1] + 2) >> 2);
  }
  const int16_t* ob = org + j.org_offset;
  for (int i = lane; i < n * n; i += 32) S.org[i] = ob[(i >> log2n) * j.org_stride + (i & (n - 1))];
  __syncwarp();
  int sum = 0;
  {
    const int16_t* R = S.line[0] + 2 * n;
    if (lane < n) sum = (j.above ? R[1 + lane] : 0) + (j.left ? R[-1 - lane] : 0);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
  }
  int dc;
  if (j.above && j.left) dc = (sum + n) / (2 * n);
  else if (j.above || j.left) dc = (sum + n / 2) / n;
  else dc = S.line[0][2 * n - 1];
  int16_t* pj = preds ? preds + pred_offset[job] : nullptr;
  if (n == 8) rough_small<8>(S, log2n, dc, j.above != 0, j.left != 0, bd, pj);
  else rough_small<4>(S, log2n, dc, j.above != 0, j.left != 0, bd, pj);
  __syncwarp();
  for (int m = lane; m < TVC_INTRA_MODES; m += 32) sad[(size_t)job * TVC_INTRA_MODES + m] = S.sad[m] >> (bd - 8);
}

static int launch_intra(tvc_ctx* c, int n, const tvc_intra_job* jobs_dev, const int16_t* lines_dev, const int16_t* org_dev,
                        uint32_t* sad_dev, int16_t* preds_dev, const int64_t* pred_offset_dev)
{
  ProfScope ps(c, TVC_PH_INTRA);
  // both kernels walk the whole list; each serves the PU sizes it is built for and leaves the others at once
  k_intra_rough<<<n, kIntraThreads, 0, c->stream>>>(n, jobs_dev, lines_dev, org_dev, sad_dev, preds_dev, pred_offset_dev, c->cfg.bit_depth);
  TVC_LAUNCH_CHECK(c);
  k_intra_rough_small<<<(n + kIntraSmallWarps - 1) / kIntraSmallWarps, kIntraSmallWarps * 32, 0, c->stream>>>(n, jobs_dev, lines_dev, org_dev, sad_dev,
                                                                                                           preds_dev, pred_offset_dev, c->cfg.bit_depth);
  TVC_LAUNCH_CHECK(c);
  return TVC_OK;
}

static int validate_intra(tvc_ctx* c, int n, const tvc_intra_job* jobs, size_t line_elems, size_t org_elems)
{
  for (int i = 0; i < n; i++) {
    const tvc_intra_job& j = jobs[i];
    const bool size_ok = j.log2_size >= 2 && j.log2_size <= 6;
    const size_t nn = size_ok ? (size_t)1 << j.log2_size : 0;
    if (!size_ok || j.line_offset < 0 || (size_t)j.line_offset + 4 * nn + 1 > line_elems || j.org_offset < 0 || j.org_stride < (int)nn ||
        (size_t)j.org_offset + (nn - 1) * (size_t)j.org_stride + nn > org_elems)
      return set_err(c, TVC_ERR_ARG, "tvc_intra_rough_batch: job %d invalid", i);
  }
  return TVC_OK;
}

}  // namespace tvc

using namespace tvc;

extern "C" {

int tvc_intra_rough_batch_dev(tvc_ctx* c, int n, const tvc_intra_job* jobs_dev, const int16_t* lines_dev, const int16_t* org_dev,
                              uint32_t* sad_dev, int16_t* preds_dev, const int64_t* pred_offset_dev)
{
  if (!c || n < 0 || (n && (!jobs_dev || !lines_dev || !org_dev || !sad_dev)) || (preds_dev && !pred_offset_dev))
    return set_err(c, TVC_ERR_ARG, "tvc_intra_rough_batch_dev: bad argument");
  if (n == 0) return TVC_OK;
  return launch_intra(c, n, jobs_dev, lines_dev, org_dev, sad_dev, preds_dev, pred_offset_dev);
}

int tvc_intra_rough_batch(tvc_ctx* c, int n, const tvc_intra_job* jobs, const int16_t* lines, size_t line_elems, const int16_t* org,
                          size_t org_elems, uint32_t* sad)
{
  if (!c || n < 0 || (n && (!jobs || !lines || !org || !sad))) return set_err(c, TVC_ERR_ARG, "tvc_intra_rough_batch: bad argument");
  if (n == 0) return TVC_OK;
  int r;
  if ((r = validate_intra(c, n, jobs, line_elems, org_elems))) return r;
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t job_b = up((size_t)n * sizeof(tvc_intra_job)), line_b = up(line_elems * 2), org_b = up(org_elems * 2);
  const size_t sad_b = (size_t)n * TVC_INTRA_MODES * 4;
  if ((r = ensure_scratch(c, c->in, job_b + line_b + org_b))) return r;
  if ((r = ensure_scratch(c, c->out, sad_b))) return r;
  char* hi = (char*)c->in.host;
  memcpy(hi, jobs, (size_t)n * sizeof(tvc_intra_job));
  memcpy(hi + job_b, lines, line_elems * 2);
  memcpy(hi + job_b + line_b, org, org_elems * 2);
  TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, hi, job_b + line_b + org_elems * 2, cudaMemcpyHostToDevice, c->stream));
  char* di = (char*)c->in.dev;
  if ((r = launch_intra(c, n, (const tvc_intra_job*)di, (const int16_t*)(di + job_b), (const int16_t*)(di + job_b + line_b),
                        (uint32_t*)c->out.dev, nullptr, nullptr)))
    return r;
  TVC_CUDA(c, cudaMemcpyAsync(c->out.host, c->out.dev, sad_b, cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(sad, c->out.host, sad_b);
  return TVC_OK;
}

int tvc_intra_rough(tvc_ctx* c, int log2_size, const int16_t* line, const int16_t* org, int org_stride, int above, int left,
                    uint32_t sad[TVC_INTRA_MODES], int16_t* preds)
{
  if (!c || log2_size < 2 || log2_size > 6 || !line || !org || !sad || org_stride < (1 << log2_size))
    return set_err(c, TVC_ERR_ARG, "tvc_intra_rough: bad argument");
  const int n = 1 << log2_size;
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  // in: [job][pred offset][line][original block, packed]; out: [35 sads][35 predictions]
  const size_t job_b = 256, line_b = up((size_t)(4 * n + 1) * 2), org_b = up((size_t)n * n * 2);
  const size_t sad_b = 256, pred_b = preds ? (size_t)TVC_INTRA_MODES * n * n * 2 : 0;
  int r;
  if ((r = ensure_scratch(c, c->in, job_b + line_b + org_b))) return r;
  if ((r = ensure_scratch(c, c->out, sad_b + pred_b))) return r;
  char* hi = (char*)c->in.host;
  tvc_intra_job j = {log2_size, 0, 0, n, above, left};
  memcpy(hi, &j, sizeof(j));
  const int64_t zero = 0;
  memcpy(hi + 128, &zero, sizeof(zero));
  memcpy(hi + job_b, line, (size_t)(4 * n + 1) * 2);
  int16_t* ho = (int16_t*)(hi + job_b + line_b);
  for (int y = 0; y < n; y++) memcpy(ho + (size_t)y * n, org + (size_t)y * org_stride, (size_t)n * 2);
  TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, hi, job_b + line_b + (size_t)n * n * 2, cudaMemcpyHostToDevice, c->stream));
  char* di = (char*)c->in.dev;
  char* dout = (char*)c->out.dev;
  if ((r = launch_intra(c, 1, (const tvc_intra_job*)di, (const int16_t*)(di + job_b), (const int16_t*)(di + job_b + line_b), (uint32_t*)dout,
                        preds ? (int16_t*)(dout + sad_b) : nullptr, (const int64_t*)(di + 128))))
    return r;
  TVC_CUDA(c, cudaMemcpyAsync(c->out.host, c->out.dev, sad_b + pred_b, cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(sad, c->out.host, TVC_INTRA_MODES * 4);
  if (preds) memcpy(preds, (char*)c->out.host + sad_b, pred_b);
  return TVC_OK;
}

}  // extern "C"
