// tvc_dist.cu -- SAD / SSE / Hadamard-SATD distortion (TComRdCost.cpp:490-989, 1314-1656, 1663-1872,
// 2122-2287).  One warp per distortion job; bit-exact with the reference's DistFunc results.
#include "tvc_internal.cuh"
#include "tvc_dist.cuh"

namespace tvc {

// whole-warp distortion of one block; result valid in every lane
__device__ uint32_t dist_warp(int kind, const int16_t* __restrict__ org, int so, const int16_t* __restrict__ cur, int sc,
                              int w, int h, int sub, int bi)
{
  const int lane = threadIdx.x & 31;
  uint32_t acc = 0;
  if (kind == TVC_DIST_SAD) {
    // rows 0, step, 2*step .. ; sum <<= sub ; >> bi        (TComRdCost.cpp:518-546 and siblings)
    int step = 1 << sub, nrows = h >> sub, total = w * nrows;
    for (int i = lane; i < total; i += 32) {
      int r = i / w, x = i - r * w;
      int d = (int)org[(r * step) * so + x] - (int)cur[(r * step) * sc + x];
      acc += (uint32_t)abs(d);
    }
    acc = warp_sum(acc);
    return (acc << sub) >> bi;
  }
  if (kind == TVC_DIST_SSE) {
    // each squared difference >> 2*bi before accumulation (TComRdCost.cpp:1336-1337)
    int total = w * h, sh = bi << 1;
    for (int i = lane; i < total; i += 32) {
      int r = i / w, x = i - r * w;
      int d = (int)org[r * so + x] - (int)cur[r * sc + x];
      acc += (uint32_t)((d * d) >> sh);
    }
    return warp_sum(acc);
  }
  // TVC_DIST_HADS: xGetHADs tiling (TComRdCost.cpp:2186-2287)
  if ((h & 7) == 0 && (w & 7) == 0) {
    int tx = w >> 3, nt = tx * (h >> 3);
    for (int t = lane; t < nt; t += 32) {
      int ty = t / tx, txx = t - ty * tx;
      acc += had_tile<8>(org + (ty * 8) * so + txx * 8, so, cur + (ty * 8) * sc + txx * 8, sc);
    }
  } else if ((h & 3) == 0 && (w & 3) == 0) {
    int tx = w >> 2, nt = tx * (h >> 2);
    for (int t = lane; t < nt; t += 32) {
      int ty = t / tx, txx = t - ty * tx;
      acc += had_tile<4>(org + (ty * 4) * so + txx * 4, so, cur + (ty * 4) * sc + txx * 4, sc);
    }
  } else {
    int tx = w >> 1, nt = tx * (h >> 1);
    for (int t = lane; t < nt; t += 32) {
      int ty = t / tx, txx = t - ty * tx;
      acc += had_tile<2>(org + (ty * 2) * so + txx * 2, so, cur + (ty * 2) * sc + txx * 2, sc);
    }
  }
  return warp_sum(acc) >> bi;
}

__global__ void __launch_bounds__(128) k_dist_batch(PlaneTable pt, int n, const tvc_dist_job* __restrict__ jobs,
                                                    uint32_t* __restrict__ out, int bi)
{
  int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= n) return;
  tvc_dist_job jb = jobs[j];
  int so = pt.stride[jb.org_plane], sc = pt.stride[jb.cur_plane];
  const int16_t* o = pt.org[jb.org_slot][jb.org_plane] + (ptrdiff_t)jb.org_y * so + jb.org_x;
  const int16_t* c = pt.org[jb.cur_slot][jb.cur_plane] + (ptrdiff_t)jb.cur_y * sc + jb.cur_x;
  uint32_t r = dist_warp(jb.kind, o, so, c, sc, jb.w, jb.h, jb.kind == TVC_DIST_SAD ? jb.sub_shift : 0, bi);
  if ((threadIdx.x & 31) == 0) out[j] = r;
}

__global__ void k_dist_ptr(int kind, const int16_t* org, int so, const int16_t* cur, int sc, int w, int h, int sub,
                           int bi, uint32_t* out)
{
  uint32_t r = dist_warp(kind, org, so, cur, sc, w, h, sub, bi);
  if (threadIdx.x == 0) *out = r;
}

static int validate_job(tvc_ctx* c, const tvc_dist_job& j)
{
  if (j.kind < 0 || j.kind > 2 || !valid_slot(c, j.org_slot) || !valid_slot(c, j.cur_slot) || j.org_plane < 0 ||
      j.org_plane > 2 || j.cur_plane < 0 || j.cur_plane > 2 || j.w <= 0 || j.h <= 0 || j.w > 128 || j.h > 128)
    return TVC_ERR_ARG;
  if (j.kind == TVC_DIST_HADS && ((j.w | j.h) & 1)) return TVC_ERR_ARG;
  if (j.kind == TVC_DIST_SAD && (j.sub_shift < 0 || j.sub_shift > 2 || (j.h & ((1 << j.sub_shift) - 1)))) return TVC_ERR_ARG;
  const Pic& po = c->pics[j.org_slot];
  const Pic& pc = c->pics[j.cur_slot];
  int a = j.org_plane, b = j.cur_plane;
  if (j.org_x < -po.mx[a] || j.org_y < -po.my[a] || j.org_x + j.w > po.w[a] + po.mx[a] || j.org_y + j.h > po.h[a] + po.my[a]) return TVC_ERR_ARG;
  if (j.cur_x < -pc.mx[b] || j.cur_y < -pc.my[b] || j.cur_x + j.w > pc.w[b] + pc.mx[b] || j.cur_y + j.h > pc.h[b] + pc.my[b]) return TVC_ERR_ARG;
  return TVC_OK;
}

}  // namespace tvc

using namespace tvc;

extern "C" {

int tvc_dist_batch_dev(tvc_ctx* c, int n, const tvc_dist_job* jobs_dev, uint32_t* out_dev)
{
  if (!c || n < 0 || !jobs_dev || !out_dev) return set_err(c, TVC_ERR_ARG, "tvc_dist_batch_dev: bad argument");
  if (n == 0) return TVC_OK;
  k_dist_batch<<<(n + 3) / 4, 128, 0, c->stream>>>(c->planes, n, jobs_dev, out_dev, c->bi);
  TVC_LAUNCH_CHECK(c);
  return TVC_OK;
}

int tvc_dist_batch(tvc_ctx* c, int n, const tvc_dist_job* jobs, uint32_t* out)
{
  if (!c || n < 0 || (n && (!jobs || !out))) return set_err(c, TVC_ERR_ARG, "tvc_dist_batch: bad argument");
  if (n == 0) return TVC_OK;
  for (int i = 0; i < n; i++)
    if (validate_job(c, jobs[i])) return set_err(c, TVC_ERR_ARG, "tvc_dist_batch: job %d invalid", i);
  int r;
  if ((r = ensure_scratch(c, c->in, (size_t)n * sizeof(tvc_dist_job)))) return r;
  if ((r = ensure_scratch(c, c->out, (size_t)n * sizeof(uint32_t)))) return r;
  memcpy(c->in.host, jobs, (size_t)n * sizeof(tvc_dist_job));
  TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, c->in.host, (size_t)n * sizeof(tvc_dist_job), cudaMemcpyHostToDevice, c->stream));
  if ((r = tvc_dist_batch_dev(c, n, (const tvc_dist_job*)c->in.dev, (uint32_t*)c->out.dev))) return r;
  TVC_CUDA(c, cudaMemcpyAsync(c->out.host, c->out.dev, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(out, c->out.host, (size_t)n * sizeof(uint32_t));
  return TVC_OK;
}

int tvc_dist_block(tvc_ctx* c, int kind, const int16_t* org, int so, const int16_t* cur, int sc, int w, int h,
                   int sub_shift, uint32_t* out)
{
  if (!c || !org || !cur || !out || kind < 0 || kind > 2 || w <= 0 || h <= 0 || w > 128 || h > 128)
    return set_err(c, TVC_ERR_ARG, "tvc_dist_block: bad argument");
  if (kind == TVC_DIST_HADS && ((w | h) & 1)) return set_err(c, TVC_ERR_ARG, "tvc_dist_block: odd Hadamard size");
  if (kind != TVC_DIST_SAD) sub_shift = 0;
  if (sub_shift < 0 || sub_shift > 2 || (h & ((1 << sub_shift) - 1))) return set_err(c, TVC_ERR_ARG, "tvc_dist_block: bad sub_shift");
  size_t blk = (size_t)w * h;
  int r;
  if ((r = ensure_scratch(c, c->in, 2 * blk * sizeof(int16_t)))) return r;
  if ((r = ensure_scratch(c, c->out, sizeof(uint32_t)))) return r;
  int16_t* hp = (int16_t*)c->in.host;
  for (int y = 0; y < h; y++) {
    memcpy(hp + (size_t)y * w, org + (ptrdiff_t)y * so, (size_t)w * 2);
    memcpy(hp + blk + (size_t)y * w, cur + (ptrdiff_t)y * sc, (size_t)w * 2);
  }
  TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, hp, 2 * blk * 2, cudaMemcpyHostToDevice, c->stream));
  const int16_t* d = (const int16_t*)c->in.dev;
  k_dist_ptr<<<1, 32, 0, c->stream>>>(kind, d, w, d + blk, w, w, h, sub_shift, c->bi, (uint32_t*)c->out.dev);
  TVC_LAUNCH_CHECK(c);
  TVC_CUDA(c, cudaMemcpyAsync(c->out.host, c->out.dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  *out = *(uint32_t*)c->out.host;
  return TVC_OK;
}

}  // extern "C"
