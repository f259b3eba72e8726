// tvc_ctx.cu -- context, device pictures (TComPicYuv mirror), region ops.
#include "tvc_internal.cuh"
#include <stdarg.h>

namespace tvc {

int set_err(tvc_ctx* c, int code, const char* fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}

int check_cuda(tvc_ctx* c, cudaError_t e, const char* what)
{
  return set_err(c, TVC_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

int ensure_scratch(tvc_ctx* c, Scratch& s, size_t bytes)
{
  if (bytes <= s.bytes) return TVC_OK;
  size_t nb = bytes + bytes / 2 + 4096;
  if (s.dev) cudaFree(s.dev);
  if (s.host) cudaFreeHost(s.host);
  s.dev = nullptr; s.host = nullptr; s.bytes = 0;
  TVC_CUDA(c, cudaMalloc(&s.dev, nb));
  TVC_CUDA(c, cudaMallocHost(&s.host, nb));
  s.bytes = nb;
  return TVC_OK;
}

bool is_pinned(const void* p)
{
  cudaPointerAttributes a;
  if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

int stage_acquire(tvc_ctx* c, Scratch& s, cudaEvent_t& ev, size_t bytes)
{
  if (!ev) TVC_CUDA(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  else TVC_CUDA(c, cudaEventSynchronize(ev));
  return ensure_scratch(c, s, bytes);
}

// int16 -> u8 copy of a whole padded luma plane (values are 0..255 for bit_depth 8)
__global__ void k_pel_to_u8(const int16_t* __restrict__ src, int sstride, uint8_t* __restrict__ dst, int dstride,
                            int wtot, int htot)
{
  int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  int y = blockIdx.y;
  if (x4 >= wtot || y >= htot) return;
  const int16_t* s = src + (size_t)y * sstride + x4;
  uint32_t v = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int p = (x4 + i < wtot) ? s[i] : 0;
    p = p < 0 ? 0 : (p > 255 ? 255 : p);
    v |= (uint32_t)p << (8 * i);
  }
  *reinterpret_cast<uint32_t*>(dst + (size_t)y * dstride + x4) = v;
}

// TComPicYuv::xExtendPicCompBorder (TComPicYuv.cpp:259-286): left/right replicate, then rows
__global__ void k_extend_lr(int16_t* org, int stride, int w, int h, int mx)
{
  int y = blockIdx.x * blockDim.y + threadIdx.y;
  if (y >= h) return;
  int16_t* row = org + (size_t)y * stride;
  int16_t l = row[0], r = row[w - 1];
  for (int x = threadIdx.x; x < mx; x += blockDim.x) {
    row[-mx + x] = l;
    row[w + x] = r;
  }
}
__global__ void k_extend_tb(int16_t* org, int stride, int w, int h, int mx, int my)
{
  int x = blockIdx.x * blockDim.x + threadIdx.x - mx;
  if (x >= w + mx) return;
  int16_t t = org[x], b = org[(size_t)(h - 1) * stride + x];
  for (int y = 1 + threadIdx.y; y <= my; y += blockDim.y) {
    org[-(ptrdiff_t)y * stride + x] = t;
    org[(size_t)(h - 1 + y) * stride + x] = b;
  }
}

// region ops: TComYuv::subtract / addClip / removeHighFreq (TComYuv.cpp:401-518, 583-633)
__global__ void k_region_op(int op, int16_t* dst, const int16_t* a, const int16_t* b, int stride,
                            int w, int h, int maxv)
{
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  size_t i = (size_t)y * stride + x;
  int v;
  if (op == 0) v = a[i] - b[i];
  else if (op == 1) { v = a[i] + b[i]; v = v < 0 ? 0 : (v > maxv ? maxv : v); }
  else v = (dst[i] << 1) - a[i];
  dst[i] = (int16_t)v;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_tmap_u8(tvc_ctx* c, CUtensorMap* m, void* base, int width_bytes, int rows, int pitch, int bx, int by)
{
  PFN_encodeTiled fn = (PFN_encodeTiled)c->encode_tiled;
  cuuint64_t gdim[2] = {(cuuint64_t)width_bytes, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)pitch};
  cuuint32_t box[2] = {(cuuint32_t)bx, (cuuint32_t)by};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(c, TVC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return TVC_OK;
}

static int refresh_u8(tvc_ctx* c, Pic& p)
{
  if (!p.buf8) return TVC_OK;
  int wtot = p.w[0] + 2 * p.mx[0], htot = p.h[0] + 2 * p.my[0];
  dim3 blk(128), grd((wtot / 4 + 127) / 128 + 1, htot);
  k_pel_to_u8<<<grd, blk, 0, c->stream>>>(p.buf[0], p.stride[0], p.buf8, p.stride8, wtot, htot);
  TVC_LAUNCH_CHECK(c);
  return TVC_OK;
}

}  // namespace tvc

using namespace tvc;

extern "C" {

int tvc_abi_version(void) { return TVC_ABI_VERSION; }

int tvc_ctx_create(const tvc_config* cfg, tvc_ctx** out)
{
  if (!cfg || !out) return TVC_ERR_ARG;
  *out = nullptr;
  if (cfg->width <= 0 || cfg->height <= 0 || (cfg->width & 1) || (cfg->height & 1)) return TVC_ERR_ARG;
  if (cfg->bit_depth != 8 && cfg->bit_depth != 10) return TVC_ERR_ARG;
  if (cfg->max_cu != 64) return TVC_ERR_ARG;
  if (cfg->num_slots < 1 || cfg->num_slots > kMaxSlots) return TVC_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev)
    return TVC_ERR_CUDA;   // no CPU fallback by design
  if (cudaSetDevice(cfg->device) != cudaSuccess) return TVC_ERR_CUDA;
  tvc_ctx* c = new tvc_ctx();
  c->cfg = *cfg;
  c->bi = cfg->bit_depth - 8;
  c->num_ctus_x = (cfg->width + cfg->max_cu - 1) / cfg->max_cu;
  c->num_ctus_y = (cfg->height + cfg->max_cu - 1) / cfg->max_cu;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return TVC_ERR_CUDA; }
  c->own_stream = true;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &c->encode_tiled, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    c->encode_tiled = nullptr;
  memset(&c->planes, 0, sizeof(c->planes));
  c->pics.resize(cfg->num_slots);
  int lm = cfg->max_cu + 16;
  for (int s = 0; s < cfg->num_slots; s++) {
    Pic& p = c->pics[s];
    for (int pl = 0; pl < 3; pl++) {
      int sh = pl ? 1 : 0;
      p.w[pl] = cfg->width >> sh; p.h[pl] = cfg->height >> sh;
      p.mx[pl] = lm >> sh; p.my[pl] = lm >> sh;
      p.stride[pl] = ((p.w[pl] + 2 * p.mx[pl]) + 63) & ~63;
      size_t elems = (size_t)p.stride[pl] * (p.h[pl] + 2 * p.my[pl]);
      if (cudaMalloc(&p.buf[pl], elems * sizeof(int16_t)) != cudaSuccess) { tvc_ctx_destroy(c); return TVC_ERR_NOMEM; }
      cudaMemsetAsync(p.buf[pl], 0, elems * sizeof(int16_t), c->stream);
      p.org[pl] = p.buf[pl] + (size_t)p.my[pl] * p.stride[pl] + p.mx[pl];
      c->planes.org[s][pl] = p.org[pl];
      c->planes.stride[pl] = p.stride[pl];
    }
    if (cfg->bit_depth == 8) {
      p.stride8 = ((p.w[0] + 2 * p.mx[0]) + 127) & ~127;
      size_t bytes = (size_t)p.stride8 * (p.h[0] + 2 * p.my[0]);
      if (cudaMalloc(&p.buf8, bytes) != cudaSuccess) { tvc_ctx_destroy(c); return TVC_ERR_NOMEM; }
      cudaMemsetAsync(p.buf8, 0, bytes, c->stream);
      p.org8 = p.buf8 + (size_t)p.my[0] * p.stride8 + p.mx[0];
      if (c->encode_tiled) {
        int rows = p.h[0] + 2 * p.my[0];
        if (make_tmap_u8(c, &p.tmap_cur, p.buf8, p.stride8, rows, p.stride8, 64, 64) == TVC_OK &&
            make_tmap_u8(c, &p.tmap_cur80, p.buf8, p.stride8, rows, p.stride8, 80, 64) == TVC_OK &&
            make_tmap_u8(c, &p.tmap_ref, p.buf8, p.stride8, rows, p.stride8, 208, 192) == TVC_OK)
          p.has_tmap = true;
      }
    }
  }
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) { tvc_ctx_destroy(c); return TVC_ERR_CUDA; }
  *out = c;
  return TVC_OK;
}

void tvc_ctx_destroy(tvc_ctx* c)
{
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  cudaDeviceSynchronize();
  for (auto& p : c->pics) {
    for (int pl = 0; pl < 3; pl++) if (p.buf[pl]) cudaFree(p.buf[pl]);
    if (p.buf8) cudaFree(p.buf8);
  }
  if (c->in.dev) cudaFree(c->in.dev);
  if (c->in.host) cudaFreeHost(c->in.host);
  if (c->out.dev) cudaFree(c->out.dev);
  if (c->out.host) cudaFreeHost(c->out.host);
  if (c->rdoq_scratch) cudaFree(c->rdoq_scratch);
  for (auto& t : c->ctu_tickets) {
    if (t.dev) cudaFree(t.dev);
    if (t.host) cudaFreeHost(t.host);
    if (t.ev) cudaEventDestroy(t.ev);
  }
  if (c->spec_stream) cudaStreamDestroy(c->spec_stream);
  if (c->spec_ev) cudaEventDestroy(c->spec_ev);
  if (c->me_tables) cudaFree(c->me_tables);
  if (c->me_centers) cudaFree(c->me_centers);
  for (tvc::Scratch* sc : {&c->me_stage, &c->fr_stage}) {
    if (sc->dev) cudaFree(sc->dev);
    if (sc->host) cudaFreeHost(sc->host);
  }
  for (auto& p : c->prof_live) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (auto e : c->prof_pool) cudaEventDestroy(e);
  if (c->fr_int_ready) cudaEventDestroy(c->fr_int_ready);
  for (auto e : c->pipe_ev) if (e) cudaEventDestroy(e);
  for (auto st : c->pipe) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  if (c->me_ev) cudaEventDestroy(c->me_ev);
  if (c->fr_ev) cudaEventDestroy(c->fr_ev);
  if (c->fr_jobs) cudaFree(c->fr_jobs);
  if (c->fr_int) cudaFree(c->fr_int);
  if (c->fr_fjobs) cudaFree(c->fr_fjobs);
  if (c->fr_frac) cudaFree(c->fr_frac);
  if (c->fr_rast) cudaFree(c->fr_rast);
  if (c->fr_sweep) cudaFree(c->fr_sweep);
  if (c->fr_packed) cudaFree(c->fr_packed);
  if (c->frac_done) cudaFree(c->frac_done);
  if (c->frac_list) cudaFree(c->frac_list);
  if (c->grp_cost) cudaFree(c->grp_cost);
  if (c->grp_order) cudaFree(c->grp_order);
  if (c->bi_buf) cudaFree(c->bi_buf);
  if (c->bi_host) cudaFreeHost(c->bi_host);
  if (c->fr_stats) cudaFree(c->fr_stats);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int tvc_ctx_set_stream(tvc_ctx* c, void* s)
{
  if (!c) return TVC_ERR_ARG;
  if (c->own_stream && c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
  c->stream = (cudaStream_t)s;
  c->own_stream = false;
  return TVC_OK;
}

int tvc_sync(tvc_ctx* c)
{
  if (!c) return TVC_ERR_ARG;
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  return TVC_OK;
}

const char* tvc_last_error(tvc_ctx* c) { return c ? c->err.c_str() : "null context"; }
uint64_t tvc_launch_count(tvc_ctx* c) { return c ? c->launches : 0; }

static int copy_plane(tvc_ctx* c, Pic& p, int pl, int16_t* host, int hstride, int with_margin, bool up)
{
  int mx = with_margin ? p.mx[pl] : 0, my = with_margin ? p.my[pl] : 0;
  int16_t* d = p.org[pl] - (ptrdiff_t)my * p.stride[pl] - mx;
  int16_t* h = host - (ptrdiff_t)my * hstride - mx;
  size_t wbytes = (size_t)(p.w[pl] + 2 * mx) * sizeof(int16_t);
  size_t rows = (size_t)(p.h[pl] + 2 * my);
  if (up) TVC_CUDA(c, cudaMemcpy2DAsync(d, (size_t)p.stride[pl] * 2, h, (size_t)hstride * 2, wbytes, rows, cudaMemcpyHostToDevice, c->stream));
  else    TVC_CUDA(c, cudaMemcpy2DAsync(h, (size_t)hstride * 2, d, (size_t)p.stride[pl] * 2, wbytes, rows, cudaMemcpyDeviceToHost, c->stream));
  return TVC_OK;
}

int tvc_pic_upload(tvc_ctx* c, int slot, const int16_t* y, int sy, const int16_t* u, const int16_t* v, int sc, int with_margin)
{
  if (!c || !valid_slot(c, slot) || !y) return set_err(c, TVC_ERR_ARG, "tvc_pic_upload: bad argument");
  Pic& p = c->pics[slot];
  int r;
  if ((r = copy_plane(c, p, 0, (int16_t*)y, sy, with_margin, true))) return r;
  if (u && (r = copy_plane(c, p, 1, (int16_t*)u, sc, with_margin, true))) return r;
  if (v && (r = copy_plane(c, p, 2, (int16_t*)v, sc, with_margin, true))) return r;
  if ((r = refresh_u8(c, p))) return r;
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  return TVC_OK;
}

int tvc_pic_download(tvc_ctx* c, int slot, int16_t* y, int sy, int16_t* u, int16_t* v, int sc, int with_margin)
{
  if (!c || !valid_slot(c, slot)) return set_err(c, TVC_ERR_ARG, "tvc_pic_download: bad argument");
  Pic& p = c->pics[slot];
  int r;
  if (y && (r = copy_plane(c, p, 0, y, sy, with_margin, false))) return r;
  if (u && (r = copy_plane(c, p, 1, u, sc, with_margin, false))) return r;
  if (v && (r = copy_plane(c, p, 2, v, sc, with_margin, false))) return r;
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  return TVC_OK;
}

int tvc_pic_extend_border(tvc_ctx* c, int slot)
{
  if (!c || !valid_slot(c, slot)) return set_err(c, TVC_ERR_ARG, "tvc_pic_extend_border: bad slot");
  Pic& p = c->pics[slot];
  for (int pl = 0; pl < 3; pl++) {
    dim3 b1(32, 8), g1((p.h[pl] + 7) / 8);
    k_extend_lr<<<g1, b1, 0, c->stream>>>(p.org[pl], p.stride[pl], p.w[pl], p.h[pl], p.mx[pl]);
    TVC_LAUNCH_CHECK(c);
    int wt = p.w[pl] + 2 * p.mx[pl];
    dim3 b2(64, 4), g2((wt + 63) / 64);
    k_extend_tb<<<g2, b2, 0, c->stream>>>(p.org[pl], p.stride[pl], p.w[pl], p.h[pl], p.mx[pl], p.my[pl]);
    TVC_LAUNCH_CHECK(c);
  }
  return refresh_u8(c, p);
}

int tvc_prof_enable(tvc_ctx* c, int on)
{
  if (!c) return TVC_ERR_ARG;
  c->prof_on = on != 0;
  return TVC_OK;
}

int tvc_prof_read(tvc_ctx* c, double* ms_sum, uint64_t* groups, int reset)
{
  if (!c) return TVC_ERR_ARG;
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  for (auto& p : c->prof_live) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { c->prof_ms[p.phase] += ms; c->prof_n[p.phase]++; }
    c->prof_pool.push_back(p.a); c->prof_pool.push_back(p.b);
  }
  c->prof_live.clear();
  for (int i = 0; i < TVC_PH_COUNT; i++) {
    if (ms_sum) ms_sum[i] = c->prof_ms[i];
    if (groups) groups[i] = c->prof_n[i];
    if (reset) { c->prof_ms[i] = 0; c->prof_n[i] = 0; }
  }
  return TVC_OK;
}

int tvc_pic_device_ptr(tvc_ctx* c, int slot, int plane, void** ptr, int* stride)
{
  if (!c || !valid_slot(c, slot) || plane < 0 || plane > 2 || !ptr) return set_err(c, TVC_ERR_ARG, "tvc_pic_device_ptr: bad argument");
  *ptr = c->pics[slot].org[plane];
  if (stride) *stride = c->pics[slot].stride[plane];
  return TVC_OK;
}

int tvc_pic_device_ptr_u8(tvc_ctx* c, int slot, void** ptr, int* stride)
{
  if (!c || !valid_slot(c, slot) || !ptr) return set_err(c, TVC_ERR_ARG, "tvc_pic_device_ptr_u8: bad argument");
  *ptr = c->pics[slot].org8;
  if (stride) *stride = c->pics[slot].stride8;
  return TVC_OK;
}

static int region_op(tvc_ctx* c, int op, int dst, int a, int b, int plane, int x, int y, int w, int h)
{
  if (!c || !valid_slot(c, dst) || !valid_slot(c, a) || (op != 2 && !valid_slot(c, b)) || plane < 0 || plane > 2)
    return set_err(c, TVC_ERR_ARG, "region op: bad slot/plane");
  Pic& p = c->pics[dst];
  if (w <= 0 || h <= 0 || x < -p.mx[plane] || y < -p.my[plane] || x + w > p.w[plane] + p.mx[plane] || y + h > p.h[plane] + p.my[plane])
    return set_err(c, TVC_ERR_ARG, "region op: rectangle outside the padded plane");
  size_t off = (size_t)((ptrdiff_t)y * p.stride[plane] + x);
  ProfScope ps(c, TVC_PH_OTHER);
  dim3 blk(32, 8), grd((w + 31) / 32, (h + 7) / 8);
  const int16_t* pb = (op != 2) ? c->pics[b].org[plane] + (ptrdiff_t)off : nullptr;
  k_region_op<<<grd, blk, 0, c->stream>>>(op, p.org[plane] + (ptrdiff_t)off, c->pics[a].org[plane] + (ptrdiff_t)off, pb,
                                          p.stride[plane], w, h, (1 << c->cfg.bit_depth) - 1);
  TVC_LAUNCH_CHECK(c);
  return TVC_OK;
}

int tvc_pic_subtract(tvc_ctx* c, int d, int a, int b, int pl, int x, int y, int w, int h) { return region_op(c, 0, d, a, b, pl, x, y, w, h); }
int tvc_pic_add_clip(tvc_ctx* c, int d, int a, int b, int pl, int x, int y, int w, int h) { return region_op(c, 1, d, a, b, pl, x, y, w, h); }
int tvc_pic_remove_high_freq(tvc_ctx* c, int d, int a, int pl, int x, int y, int w, int h) { return region_op(c, 2, d, a, -1, pl, x, y, w, h); }

}  // extern "C"
