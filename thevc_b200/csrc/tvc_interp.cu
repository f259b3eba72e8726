// tvc_interp.cu -- TComInterpolationFilter drop-ins and batched motion compensation
// (TComInterpolationFilter.cpp:325-415; TComPrediction.cpp:410-658; TComYuv.cpp:520-581).
#include "tvc_internal.cuh"
#include "tvc_interp.cuh"
#include "tvc_dist.cuh"

namespace tvc {

// generic block filter on a dense staging buffer: one thread per output sample
template <int N>
__global__ void k_filter_block(const int16_t* __restrict__ src, int ss, int16_t* __restrict__ dst, int w, int h,
                               int frac, int isVert, int isFirst, int isLast, int bd)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * h) return;
  int y = i / w, x = i - y * w;
  dst[i] = if_sample<N>(src + (size_t)y * ss + x, isVert ? ss : 1, frac, isFirst != 0, isLast != 0, bd);
}

static int filter_dropin(tvc_ctx* c, int ntaps, int isVert, const int16_t* src, int ss, int16_t* dst, int ds, int w,
                         int h, int frac, int isFirst, int isLast)
{
  if (!c || !src || !dst || w <= 0 || h <= 0 || w > 256 || h > 256 || frac < 0 || frac >= (ntaps == 8 ? 4 : 8))
    return set_err(c, TVC_ERR_ARG, "tvc_filter_*: bad argument");
  int before = frac ? (ntaps / 2 - 1) : 0, after = frac ? (ntaps / 2) : 0;
  int x0 = isVert ? 0 : -before, x1 = isVert ? w : w + after;
  int y0 = isVert ? -before : 0, y1 = isVert ? h + after : h;
  int sw = x1 - x0, sh = y1 - y0;
  size_t in_bytes = (size_t)sw * sh * 2, out_bytes = (size_t)w * h * 2;
  int r;
  if ((r = ensure_scratch(c, c->in, in_bytes))) return r;
  if ((r = ensure_scratch(c, c->out, out_bytes))) return r;
  int16_t* hp = (int16_t*)c->in.host;
  for (int y = 0; y < sh; y++) memcpy(hp + (size_t)y * sw, src + (ptrdiff_t)(y0 + y) * ss + x0, (size_t)sw * 2);
  TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, hp, in_bytes, cudaMemcpyHostToDevice, c->stream));
  const int16_t* dsrc = (const int16_t*)c->in.dev + (size_t)(-y0) * sw + (-x0);
  int n = w * h;
  if (ntaps == 8)
    k_filter_block<8><<<(n + 255) / 256, 256, 0, c->stream>>>(dsrc, sw, (int16_t*)c->out.dev, w, h, frac, isVert, isFirst, isLast, c->cfg.bit_depth);
  else
    k_filter_block<4><<<(n + 255) / 256, 256, 0, c->stream>>>(dsrc, sw, (int16_t*)c->out.dev, w, h, frac, isVert, isFirst, isLast, c->cfg.bit_depth);
  TVC_LAUNCH_CHECK(c);
  TVC_CUDA(c, cudaMemcpyAsync(c->out.host, c->out.dev, out_bytes, cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  const int16_t* op = (const int16_t*)c->out.host;
  for (int y = 0; y < h; y++) memcpy(dst + (ptrdiff_t)y * ds, op + (size_t)y * w, (size_t)w * 2);
  return TVC_OK;
}

// ------------------------------------------------------------------------------------------ MC
// One CTA per (PU, plane).  xPredInterLumaBlk / xPredInterChromaBlk (TComPrediction.cpp:554-645):
// frac in one direction -> single 1-D pass; both -> horizontal pass over h+N-1 rows into shared
// memory (14-bit intermediates) then vertical pass.  Bi-prediction keeps both lists at 14 bits
// and averages with TComYuv::addAvg's rounding (TComYuv.cpp:537-549).
constexpr int kMcMaxW = 64, kMcMaxH = 64;

template <int N>
__device__ void mc_one_list(const int16_t* __restrict__ ref, int rs, int mvx, int mvy, int w, int h, bool bi, int bd,
                            int16_t* __restrict__ tmp /* smem w*(h+N-1) */, int16_t* __restrict__ dst, int ds)
{
  constexpr int FB = (N == 8) ? 2 : 3;           // fractional bits
  constexpr int FM = (1 << FB) - 1;
  ref += (mvx >> FB) + (ptrdiff_t)(mvy >> FB) * rs;
  int xf = mvx & FM, yf = mvy & FM;
  bool last = !bi;
  int n = w * h;
  if (yf == 0) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      int y = i / w, x = i - y * w;
      dst[(size_t)y * ds + x] = if_sample<N>(ref + (ptrdiff_t)y * rs + x, 1, xf, true, last, bd);
    }
  } else if (xf == 0) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      int y = i / w, x = i - y * w;
      dst[(size_t)y * ds + x] = if_sample<N>(ref + (ptrdiff_t)y * rs + x, rs, yf, true, last, bd);
    }
  } else {
    constexpr int HB = N / 2 - 1;
    int th = h + N - 1;
    const int16_t* r0 = ref - (ptrdiff_t)HB * rs;
    for (int i = threadIdx.x; i < w * th; i += blockDim.x) {
      int y = i / w, x = i - y * w;
      tmp[i] = if_sample<N>(r0 + (ptrdiff_t)y * rs + x, 1, xf, true, false, bd);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      int y = i / w, x = i - y * w;
      dst[(size_t)y * ds + x] = if_sample<N>(tmp + (y + HB) * w + x, w, yf, false, last, bd);
    }
  }
}

__global__ void __launch_bounds__(256) k_mc_batch(PlaneTable pt, int dst_slot, int n, const tvc_pu* __restrict__ pus, int bd)
{
  __shared__ int16_t s_tmp[kMcMaxW * (kMcMaxH + 7)];
  __shared__ int16_t s_l0[kMcMaxW * kMcMaxH];
  __shared__ int16_t s_l1[kMcMaxW * kMcMaxH];
  int pl = blockIdx.y;
  tvc_pu pu = pus[blockIdx.x];
  int sh = pl ? 1 : 0;
  int w = pu.w >> sh, h = pu.h >> sh, x0 = pu.x >> sh, y0 = pu.y >> sh;
  int stride = pt.stride[pl];
  int16_t* dst = pt.org[dst_slot][pl] + (ptrdiff_t)y0 * stride + x0;
  bool use0 = pu.ref_slot0 >= 0, use1 = pu.ref_slot1 >= 0;
  bool bi = use0 && use1;
  if (!bi) {
    int slot = use0 ? pu.ref_slot0 : pu.ref_slot1;
    int mvx = use0 ? pu.mvx0 : pu.mvx1, mvy = use0 ? pu.mvy0 : pu.mvy1;
    const int16_t* ref = pt.org[slot][pl] + (ptrdiff_t)y0 * stride + x0;
    if (pl == 0) mc_one_list<8>(ref, stride, mvx, mvy, w, h, false, bd, s_tmp, dst, stride);
    else         mc_one_list<4>(ref, stride, mvx, mvy, w, h, false, bd, s_tmp, dst, stride);
    return;
  }
  const int16_t* ref0 = pt.org[pu.ref_slot0][pl] + (ptrdiff_t)y0 * stride + x0;
  const int16_t* ref1 = pt.org[pu.ref_slot1][pl] + (ptrdiff_t)y0 * stride + x0;
  if (pl == 0) mc_one_list<8>(ref0, stride, pu.mvx0, pu.mvy0, w, h, true, bd, s_tmp, s_l0, w);
  else         mc_one_list<4>(ref0, stride, pu.mvx0, pu.mvy0, w, h, true, bd, s_tmp, s_l0, w);
  __syncthreads();
  if (pl == 0) mc_one_list<8>(ref1, stride, pu.mvx1, pu.mvy1, w, h, true, bd, s_tmp, s_l1, w);
  else         mc_one_list<4>(ref1, stride, pu.mvx1, pu.mvy1, w, h, true, bd, s_tmp, s_l1, w);
  __syncthreads();
  int shiftNum = kIfPrec + 1 - bd;
  int offset = (1 << (shiftNum - 1)) + 2 * kIfOffs;
  int maxv = (1 << bd) - 1;
  for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
    int y = i / w, x = i - y * w;
    int v = ((int)s_l0[i] + (int)s_l1[i] + offset) >> shiftNum;
    v = v < 0 ? 0 : (v > maxv ? maxv : v);
    dst[(size_t)y * stride + x] = (int16_t)v;
  }
}

// luma prediction of one candidate (uni: clipped pels; both lists: 14-bit intermediates + addAvg) kept in shared memory, then its
// distortion against the original block of cur_slot: the body of xGetInterPredictionError (TEncSearch.cpp:3059-3081: motionCompensation
// + setDistParam(bHadamard) + DistFunc) and of xGetTemplateCost's prediction + SAD (TEncSearch.cpp:4057-4118).  CTA per candidate.
__global__ void __launch_bounds__(256) k_pred_cost(PlaneTable pt, int cur_slot, int kind, int n, const tvc_pu* __restrict__ pus,
                                                   uint32_t* __restrict__ out, int bd)
{
  __shared__ int16_t s_tmp[kMcMaxW * (kMcMaxH + 7)];
  __shared__ int16_t s_l0[kMcMaxW * kMcMaxH];
  __shared__ int16_t s_l1[kMcMaxW * kMcMaxH];
  __shared__ uint32_t s_sum;
  const tvc_pu pu = pus[blockIdx.x];
  const int w = pu.w, h = pu.h, stride = pt.stride[0];
  const bool use0 = pu.ref_slot0 >= 0, use1 = pu.ref_slot1 >= 0, bi = use0 && use1;
  if (threadIdx.x == 0) s_sum = 0;
  if (!bi) {
    const int slot = use0 ? pu.ref_slot0 : pu.ref_slot1;
    const int mvx = use0 ? pu.mvx0 : pu.mvx1, mvy = use0 ? pu.mvy0 : pu.mvy1;
    mc_one_list<8>(pt.org[slot][0] + (ptrdiff_t)pu.y * stride + pu.x, stride, mvx, mvy, w, h, false, bd, s_tmp, s_l0, w);
  } else {
    mc_one_list<8>(pt.org[pu.ref_slot0][0] + (ptrdiff_t)pu.y * stride + pu.x, stride, pu.mvx0, pu.mvy0, w, h, true, bd, s_tmp, s_l0, w);
    __syncthreads();
    mc_one_list<8>(pt.org[pu.ref_slot1][0] + (ptrdiff_t)pu.y * stride + pu.x, stride, pu.mvx1, pu.mvy1, w, h, true, bd, s_tmp, s_l1, w);
    __syncthreads();
    const int shiftNum = kIfPrec + 1 - bd, offset = (1 << (shiftNum - 1)) + 2 * kIfOffs, maxv = (1 << bd) - 1;
    for (int i = threadIdx.x; i < w * h; i += blockDim.x) {      // TComYuv::addAvg, TComYuv.cpp:539-549
      int v = ((int)s_l0[i] + (int)s_l1[i] + offset) >> shiftNum;
      s_l0[i] = (int16_t)(v < 0 ? 0 : (v > maxv ? maxv : v));
    }
  }
  __syncthreads();
  const int16_t* org = pt.org[cur_slot][0] + (ptrdiff_t)pu.y * stride + pu.x;
  uint32_t acc = 0;
  if (kind == TVC_DIST_SAD) {
    for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
      const int y = i / w, x = i - y * w;
      acc += (uint32_t)abs((int)org[(ptrdiff_t)y * stride + x] - (int)s_l0[i]);
    }
  } else if (!(w & 7) && !(h & 7)) {        // xGetHADs tiling (TComRdCost.cpp:2186-2287): PU sides are multiples of 4
    const int tx = w >> 3, nt = tx * (h >> 3);
    for (int t = threadIdx.x; t < nt; t += blockDim.x) {
      const int ty = t / tx, txx = t - ty * tx;
      acc += had_tile<8>(org + (ptrdiff_t)(ty * 8) * stride + txx * 8, stride, s_l0 + (ty * 8) * w + txx * 8, w);
    }
  } else {
    const int tx = w >> 2, nt = tx * (h >> 2);
    for (int t = threadIdx.x; t < nt; t += blockDim.x) {
      const int ty = t / tx, txx = t - ty * tx;
      acc += had_tile<4>(org + (ptrdiff_t)(ty * 4) * stride + txx * 4, stride, s_l0 + (ty * 4) * w + txx * 4, w);
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&s_sum, acc);
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = s_sum >> (bd - 8);
}

// ---- CTU-wide candidate cost grids (the look-up form of merge / AMVP candidate evaluation)
// One CTA per (CTU, reference, clipped MV): the 64x64 luma prediction at that MV (reference window staged into shared memory with
// coordinates clamped to the padded plane = the picture's replicated border, so every sample equals what a PU-sized prediction
// reads), then, against the original CTU, the SAD of every 4x4 block, the xCalcHADs4x4 value of every 4x4 tile and the xCalcHADs8x8
// value of every 8x8 tile, each as an inclusive 2-D prefix sum (17x17, 17x17, 9x9).  The distortion of ANY PU of the CTU with this
// motion is then four look-ups: a PU's SAD is the sum of its blocks (xGetSAD*, iSubShift 0), its SATD the sum of its tiles (xGetHADs
// tiles from the PU origin; PU origins are multiples of 8 whenever both sides are, else multiples of 4).
constexpr int kGridWin = 64 + 7;
__global__ void __launch_bounds__(256) k_ctu_cost_grids(PlaneTable pt, int cur_slot, int n, const tvc_grid_job* __restrict__ jobs,
                                                        uint32_t* __restrict__ out, int bd, int plane_w, int plane_h, int mx, int my)
{
  __shared__ int16_t s_win[kGridWin * kGridWin];
  __shared__ int16_t s_tmp[kGridWin * 64];
  __shared__ int16_t s_pred[64 * 64];
  __shared__ uint32_t s_g[3][17 * 17];
  const tvc_grid_job j = jobs[blockIdx.x];
  const int stride = pt.stride[0], tid = threadIdx.x;
  const int xf = j.mvx & 3, yf = j.mvy & 3;
  const int wx0 = j.x0 + (j.mvx >> 2) - 3, wy0 = j.y0 + (j.mvy >> 2) - 3;
  const int16_t* ref = pt.org[j.ref_slot][0];
  for (int i = tid; i < kGridWin * kGridWin; i += 256) {
    const int y = i / kGridWin, x = i - y * kGridWin;
    const int gx = min(max(wx0 + x, -mx), plane_w + mx - 1), gy = min(max(wy0 + y, -my), plane_h + my - 1);
    s_win[i] = ref[(ptrdiff_t)gy * stride + gx];
  }
  __syncthreads();
  // xPredInterLumaBlk, bi = false (TComPrediction.cpp:554-590)
  if (yf == 0) {
    for (int i = tid; i < 4096; i += 256) s_pred[i] = if_sample<8>(s_win + ((i >> 6) + 3) * kGridWin + (i & 63) + 3, 1, xf, true, true, bd);
  } else if (xf == 0) {
    for (int i = tid; i < 4096; i += 256) s_pred[i] = if_sample<8>(s_win + ((i >> 6) + 3) * kGridWin + (i & 63) + 3, kGridWin, yf, true, true, bd);
  } else {
    for (int i = tid; i < kGridWin * 64; i += 256) s_tmp[i] = if_sample<8>(s_win + (i >> 6) * kGridWin + (i & 63) + 3, 1, xf, true, false, bd);
    __syncthreads();
    for (int i = tid; i < 4096; i += 256) s_pred[i] = if_sample<8>(s_tmp + ((i >> 6) + 3) * 64 + (i & 63), 64, yf, false, true, bd);
  }
  for (int i = tid; i < 3 * 17 * 17; i += 256) (&s_g[0][0])[i] = 0;
  __syncthreads();
  const int16_t* org = pt.org[cur_slot][0] + (ptrdiff_t)j.y0 * stride + j.x0;      // rows / columns past the picture lie in the slot's margin
  {
    const int by = tid >> 4, bx = tid & 15;
    const int16_t* o = org + (ptrdiff_t)(by * 4) * stride + bx * 4;
    const int16_t* p = s_pred + (by * 4) * 64 + bx * 4;
    uint32_t sad = 0;
#pragma unroll
    for (int y = 0; y < 4; y++)
#pragma unroll
      for (int x = 0; x < 4; x++) sad += (uint32_t)abs((int)o[(ptrdiff_t)y * stride + x] - (int)p[y * 64 + x]);
    s_g[0][(by + 1) * 17 + bx + 1] = sad;
    s_g[1][(by + 1) * 17 + bx + 1] = had_tile<4>(o, stride, p, 64);
    if (tid < 64) {
      const int ty = tid >> 3, tx = tid & 7;
      s_g[2][(ty + 1) * 17 + tx + 1] = had_tile<8>(org + (ptrdiff_t)(ty * 8) * stride + tx * 8, stride, s_pred + (ty * 8) * 64 + tx * 8, 64);
    }
  }
  __syncthreads();
  // prefix sums: rows, then columns (thread per line)
  if (tid < 3 * 16) {
    const int g = tid >> 4, r = (tid & 15) + 1;
    if (g < 2 || r <= 8) { uint32_t a = 0; for (int c = 1; c <= (g < 2 ? 16 : 8); c++) { a += s_g[g][r * 17 + c]; s_g[g][r * 17 + c] = a; } }
  }
  __syncthreads();
  if (tid < 3 * 16) {
    const int g = tid >> 4, c = (tid & 15) + 1;
    if (g < 2 || c <= 8) { uint32_t a = 0; for (int r = 1; r <= (g < 2 ? 16 : 8); r++) { a += s_g[g][r * 17 + c]; s_g[g][r * 17 + c] = a; } }
  }
  __syncthreads();
  uint32_t* o = out + (size_t)blockIdx.x * TVC_GRID_WORDS;
  for (int i = tid; i < TVC_GRID_WORDS; i += 256) {
    uint32_t v;
    if (i < 289) v = s_g[0][i];
    else if (i < 578) v = s_g[1][i - 289];
    else { const int k = i - 578; v = s_g[2][(k / 9) * 17 + (k % 9)]; }
    o[i] = v;
  }
}

// one PU, one list, dense output [Y w*h][U (w/2)(h/2)][V]: the drop-in for xPredInterUni
__global__ void __launch_bounds__(256) k_mc_block(PlaneTable pt, int ref_slot, int x, int y, int w, int h, int mvx, int mvy, int bi,
                                                  int bd, int16_t* __restrict__ out)
{
  __shared__ int16_t s_tmp[kMcMaxW * (kMcMaxH + 7)];
  const int pl = blockIdx.x, sh = pl ? 1 : 0;
  const int pw = w >> sh, ph = h >> sh, x0 = x >> sh, y0 = y >> sh, stride = pt.stride[pl];
  const int16_t* ref = pt.org[ref_slot][pl] + (ptrdiff_t)y0 * stride + x0;
  int16_t* dst = out + (pl == 0 ? 0 : (pl == 1 ? w * h : w * h + pw * ph));
  if (pl == 0) mc_one_list<8>(ref, stride, mvx, mvy, pw, ph, bi != 0, bd, s_tmp, dst, pw);
  else         mc_one_list<4>(ref, stride, mvx, mvy, pw, ph, bi != 0, bd, s_tmp, dst, pw);
}

}  // namespace tvc

using namespace tvc;

extern "C" {

int tvc_filter_hor_luma(tvc_ctx* c, const int16_t* s, int ss, int16_t* d, int ds, int w, int h, int frac, int is_last)
{ return filter_dropin(c, 8, 0, s, ss, d, ds, w, h, frac, 1, is_last != 0); }
int tvc_filter_ver_luma(tvc_ctx* c, const int16_t* s, int ss, int16_t* d, int ds, int w, int h, int frac, int is_first, int is_last)
{ return filter_dropin(c, 8, 1, s, ss, d, ds, w, h, frac, is_first != 0, is_last != 0); }
int tvc_filter_hor_chroma(tvc_ctx* c, const int16_t* s, int ss, int16_t* d, int ds, int w, int h, int frac, int is_last)
{ return filter_dropin(c, 4, 0, s, ss, d, ds, w, h, frac, 1, is_last != 0); }
int tvc_filter_ver_chroma(tvc_ctx* c, const int16_t* s, int ss, int16_t* d, int ds, int w, int h, int frac, int is_first, int is_last)
{ return filter_dropin(c, 4, 1, s, ss, d, ds, w, h, frac, is_first != 0, is_last != 0); }

int tvc_mc_block(tvc_ctx* c, int ref_slot, int x, int y, int w, int h, int mvx, int mvy, int bi, int16_t* dst_y, int stride_y,
                 int16_t* dst_u, int16_t* dst_v, int stride_c)
{
  if (!c || !valid_slot(c, ref_slot) || !dst_y || !dst_u || !dst_v || w <= 0 || h <= 0 || w > 64 || h > 64 || (w & 3) || (h & 3) ||
      x < 0 || y < 0)
    return set_err(c, TVC_ERR_ARG, "tvc_mc_block: bad argument");
  const Pic& p = c->pics[ref_slot];
  const int ix = x + (mvx >> 2), iy = y + (mvy >> 2);
  if (ix - 3 < -p.mx[0] || iy - 3 < -p.my[0] || ix + w + 4 > p.w[0] + p.mx[0] || iy + h + 4 > p.h[0] + p.my[0])
    return set_err(c, TVC_ERR_ARG, "tvc_mc_block: MV reaches outside the padded picture (clipMv first)");
  const size_t elems = (size_t)w * h * 3 / 2;
  int r;
  if ((r = ensure_scratch(c, c->out, elems * 2))) return r;
  {
    ProfScope ps(c, TVC_PH_MC);
    k_mc_block<<<3, 256, 0, c->stream>>>(c->planes, ref_slot, x, y, w, h, mvx, mvy, bi, c->cfg.bit_depth, (int16_t*)c->out.dev);
    TVC_LAUNCH_CHECK(c);
  }
  TVC_CUDA(c, cudaMemcpyAsync(c->out.host, c->out.dev, elems * 2, cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  const int16_t* o = (const int16_t*)c->out.host;
  for (int r2 = 0; r2 < h; r2++) memcpy(dst_y + (ptrdiff_t)r2 * stride_y, o + (size_t)r2 * w, (size_t)w * 2);
  const int cw = w >> 1, ch = h >> 1;
  const int16_t* ou = o + (size_t)w * h;
  const int16_t* ov = ou + (size_t)cw * ch;
  for (int r2 = 0; r2 < ch; r2++) {
    memcpy(dst_u + (ptrdiff_t)r2 * stride_c, ou + (size_t)r2 * cw, (size_t)cw * 2);
    memcpy(dst_v + (ptrdiff_t)r2 * stride_c, ov + (size_t)r2 * cw, (size_t)cw * 2);
  }
  return TVC_OK;
}

int tvc_mc_batch_dev(tvc_ctx* c, int dst_slot, int n, const tvc_pu* pus_dev)
{
  if (!c || !valid_slot(c, dst_slot) || n < 0 || (n && !pus_dev)) return set_err(c, TVC_ERR_ARG, "tvc_mc_batch_dev: bad argument");
  if (n == 0) return TVC_OK;
  dim3 grd(n, 3);
  ProfScope ps(c, TVC_PH_MC);
  k_mc_batch<<<grd, 256, 0, c->stream>>>(c->planes, dst_slot, n, pus_dev, c->cfg.bit_depth);
  TVC_LAUNCH_CHECK(c);
  return TVC_OK;
}

static int validate_pus(tvc_ctx* c, int plane_slot, int n, const tvc_pu* pus, const char* who)
{
  const Pic& p = c->pics[plane_slot];
  for (int i = 0; i < n; i++) {
    const tvc_pu& u = pus[i];
    bool ok = u.w > 0 && u.h > 0 && u.w <= 64 && u.h <= 64 && !(u.w & 3) && !(u.h & 3) && u.x >= 0 && u.y >= 0 &&
              u.x + u.w <= p.w[0] + p.mx[0] && u.y + u.h <= p.h[0] + p.my[0] && (u.ref_slot0 >= 0 || u.ref_slot1 >= 0) &&
              (u.ref_slot0 < 0 || valid_slot(c, u.ref_slot0)) && (u.ref_slot1 < 0 || valid_slot(c, u.ref_slot1));
    // the 8-tap window of a clipped MV stays inside the margin (TComDataCU::clipMv, TComDataCU.cpp:3505)
    for (int l = 0; ok && l < 2; l++) {
      int slot = l ? u.ref_slot1 : u.ref_slot0;
      if (slot < 0) continue;
      int mvx = l ? u.mvx1 : u.mvx0, mvy = l ? u.mvy1 : u.mvy0;
      int ix = u.x + (mvx >> 2), iy = u.y + (mvy >> 2);
      ok = ix - 3 >= -p.mx[0] && iy - 3 >= -p.my[0] && ix + u.w + 4 <= p.w[0] + p.mx[0] && iy + u.h + 4 <= p.h[0] + p.my[0];
    }
    if (!ok) return set_err(c, TVC_ERR_ARG, "%s: PU %d invalid or MV reaches outside the padded picture", who, i);
  }
  return TVC_OK;
}

int tvc_mc_batch(tvc_ctx* c, int dst_slot, int n, const tvc_pu* pus)
{
  if (!c || !valid_slot(c, dst_slot) || n < 0 || (n && !pus)) return set_err(c, TVC_ERR_ARG, "tvc_mc_batch: bad argument");
  if (n == 0) return TVC_OK;
  int r;
  if ((r = validate_pus(c, dst_slot, n, pus, "tvc_mc_batch"))) return r;
  if ((r = ensure_scratch(c, c->in, (size_t)n * sizeof(tvc_pu)))) return r;
  memcpy(c->in.host, pus, (size_t)n * sizeof(tvc_pu));
  TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, c->in.host, (size_t)n * sizeof(tvc_pu), cudaMemcpyHostToDevice, c->stream));
  if ((r = tvc_mc_batch_dev(c, dst_slot, n, (const tvc_pu*)c->in.dev))) return r;
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  return TVC_OK;
}

int tvc_pred_cost_batch_dev(tvc_ctx* c, int cur_slot, int kind, int n, const tvc_pu* pus_dev, uint32_t* dist_dev)
{
  if (!c || !valid_slot(c, cur_slot) || n < 0 || (kind != TVC_DIST_SAD && kind != TVC_DIST_HADS) || (n && (!pus_dev || !dist_dev)))
    return set_err(c, TVC_ERR_ARG, "tvc_pred_cost_batch_dev: bad argument (kind is TVC_DIST_SAD or TVC_DIST_HADS)");
  if (n == 0) return TVC_OK;
  ProfScope ps(c, TVC_PH_MC);
  k_pred_cost<<<n, 256, 0, c->stream>>>(c->planes, cur_slot, kind, n, pus_dev, dist_dev, c->cfg.bit_depth);
  TVC_LAUNCH_CHECK(c);
  return TVC_OK;
}

int tvc_pred_cost_batch(tvc_ctx* c, int cur_slot, int kind, int n, const tvc_pu* pus, uint32_t* dist)
{
  if (!c || !valid_slot(c, cur_slot) || n < 0 || (kind != TVC_DIST_SAD && kind != TVC_DIST_HADS) || (n && (!pus || !dist)))
    return set_err(c, TVC_ERR_ARG, "tvc_pred_cost_batch: bad argument (kind is TVC_DIST_SAD or TVC_DIST_HADS)");
  if (n == 0) return TVC_OK;
  int r;
  if ((r = validate_pus(c, cur_slot, n, pus, "tvc_pred_cost_batch"))) return r;
  for (int i = 0; i < n; i++)
    if (pus[i].x + pus[i].w > c->pics[cur_slot].w[0] || pus[i].y + pus[i].h > c->pics[cur_slot].h[0])
      return set_err(c, TVC_ERR_ARG, "tvc_pred_cost_batch: PU %d lies outside the picture", i);
  if ((r = ensure_scratch(c, c->in, (size_t)n * sizeof(tvc_pu)))) return r;
  if ((r = ensure_scratch(c, c->out, (size_t)n * 4))) return r;
  memcpy(c->in.host, pus, (size_t)n * sizeof(tvc_pu));
  TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, c->in.host, (size_t)n * sizeof(tvc_pu), cudaMemcpyHostToDevice, c->stream));
  if ((r = tvc_pred_cost_batch_dev(c, cur_slot, kind, n, (const tvc_pu*)c->in.dev, (uint32_t*)c->out.dev))) return r;
  TVC_CUDA(c, cudaMemcpyAsync(c->out.host, c->out.dev, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(dist, c->out.host, (size_t)n * 4);
  return TVC_OK;
}

int tvc_ctu_cost_grids_dev(tvc_ctx* c, int cur_slot, int n, const tvc_grid_job* jobs_dev, uint32_t* grids_dev)
{
  if (!c || !valid_slot(c, cur_slot) || n < 0 || (n && (!jobs_dev || !grids_dev))) return set_err(c, TVC_ERR_ARG, "tvc_ctu_cost_grids_dev: bad argument");
  if (n == 0) return TVC_OK;
  const Pic& p = c->pics[cur_slot];
  ProfScope ps(c, TVC_PH_MC);
  k_ctu_cost_grids<<<n, 256, 0, c->stream>>>(c->planes, cur_slot, n, jobs_dev, grids_dev, c->cfg.bit_depth, p.w[0], p.h[0], p.mx[0], p.my[0]);
  TVC_LAUNCH_CHECK(c);
  return TVC_OK;
}

int tvc_ctu_cost_grids(tvc_ctx* c, int cur_slot, int n, const tvc_grid_job* jobs, uint32_t* grids)
{
  if (!c || !valid_slot(c, cur_slot) || n < 0 || (n && (!jobs || !grids))) return set_err(c, TVC_ERR_ARG, "tvc_ctu_cost_grids: bad argument");
  if (n == 0) return TVC_OK;
  const Pic& p = c->pics[cur_slot];
  for (int i = 0; i < n; i++) {
    const tvc_grid_job& j = jobs[i];
    // the CTU lies in the picture; the MV may point anywhere within +-2^14 quarter pels (reads are clamped to the padded plane)
    if (!valid_slot(c, j.ref_slot) || j.x0 < 0 || j.y0 < 0 || (j.x0 & 63) || (j.y0 & 63) || j.x0 >= p.w[0] || j.y0 >= p.h[0] ||
        j.mvx < -(1 << 14) || j.mvx > (1 << 14) || j.mvy < -(1 << 14) || j.mvy > (1 << 14))
      return set_err(c, TVC_ERR_ARG, "tvc_ctu_cost_grids: job %d invalid", i);
  }
  int r;
  const size_t out_b = (size_t)n * TVC_GRID_WORDS * 4;
  if ((r = ensure_scratch(c, c->in, (size_t)n * sizeof(tvc_grid_job)))) return r;
  if ((r = ensure_scratch(c, c->out, out_b))) return r;
  memcpy(c->in.host, jobs, (size_t)n * sizeof(tvc_grid_job));
  TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, c->in.host, (size_t)n * sizeof(tvc_grid_job), cudaMemcpyHostToDevice, c->stream));
  if ((r = tvc_ctu_cost_grids_dev(c, cur_slot, n, (const tvc_grid_job*)c->in.dev, (uint32_t*)c->out.dev))) return r;
  TVC_CUDA(c, cudaMemcpyAsync(c->out.host, c->out.dev, out_b, cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(grids, c->out.host, out_b);
  return TVC_OK;
}

}  // extern "C"
