// tvc_deblock.cu -- in-loop deblocking filter of one picture, in place on a device-resident picture slot.
//
// Replaces the sample work under TComLoopFilter::loopFilterPic (TComLoopFilter.cpp:153-191): xEdgeFilterLuma (:571-678),
// xEdgeFilterChroma (:680-797), xPelFilterLuma (:799-867), xPelFilterChroma (:869-892), xUseStrongFiltering (:901-911),
// xCalcDP / xCalcDQ (:913-921), tctable_8x8 / betatable_8x8 (:56-64).  Boundary strengths, QPs and no-filter flags come
// from the host as one record per 4-pel edge unit (include/thevc_cuda.h).
//
// Mapping: HBM-bound byte shuffling, no reuse.  One thread per LINE of an edge unit (4 lines per unit, the 4 threads of
// a unit are adjacent lanes): a vertical-edge line is 8 contiguous pels (two 8-byte loads / stores), a horizontal-edge
// line is a column, and adjacent lanes hold adjacent columns (coalesced rows).  The filter decision of a unit needs the
// second derivatives of its lines 0 and 3: two shuffles.  The same four threads then filter the unit's chroma samples
// (2 lines x Cb, Cr) when the unit lies on the 16-pel grid and bs > 1.
#include "tvc_internal.cuh"

namespace tvc {

static __constant__ uint8_t c_tc[54] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                        2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 5, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 22, 24};
static __constant__ uint8_t c_beta[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15,
                                          16, 17, 18, 20, 22, 24, 26, 28, 30, 32, 34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56, 58, 60, 62, 64};
// g_aucChromaScale, CHROMA_QP_EXTENSION table (TComRom.cpp:380-386)
static __constant__ uint8_t c_chroma_scale[58] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28,
                                                  29, 29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51};

__device__ __forceinline__ int clip3i(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

// DIR 0: vertical edges (filter across columns, lines are rows); DIR 1: horizontal edges
template <int DIR>
__global__ void __launch_bounds__(256)
k_deblock(int16_t* __restrict__ Y, int sy, int16_t* __restrict__ U, int16_t* __restrict__ V, int sc, int width, int height,
          const tvc_dbk_unit* __restrict__ units, int units_w, int units_h, int beta_off2, int tc_off2, int bd)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int line = t & 3, unit = t >> 2;
  const bool in_range = unit < units_w * units_h;
  // unit index: vertical edges -> (row of 4 lines uy, edge column ux): x = ux*8, y = uy*4; horizontal -> x = ux*4, y = uy*8
  const int ux = in_range ? unit % units_w : 0, uy = in_range ? unit / units_w : 0;
  tvc_dbk_unit u = {0, 0, 0, 0};
  if (in_range) u = units[unit];
  const int x = DIR == 0 ? ux * 8 : ux * 4 + line;
  const int y = DIR == 0 ? uy * 4 + line : uy * 8;
  const bool live = in_range && u.bs != 0 && x < width && y < height && (DIR == 0 ? x > 0 : y > 0);
  const int off = DIR == 0 ? 1 : sy;
  const int scale = 1 << (bd - 8);
  const int maxv = (1 << bd) - 1;
  const bool p_keep = (u.flags & 1) != 0, q_keep = (u.flags & 2) != 0;

  // ---- luma (xEdgeFilterLuma)
  int m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int16_t* p = Y + (ptrdiff_t)y * sy + x;         // sample m4 (first of part Q)
  if (live) {
    if (DIR == 0) {
      // x is a multiple of 8 and pel 0 of a row is 16-byte aligned: the P half starts 8 bytes before a 16-byte boundary
      const uint2 a = *reinterpret_cast<const uint2*>(p - 4), b = *reinterpret_cast<const uint2*>(p);
      m[0] = (int16_t)(a.x & 0xffff); m[1] = (int16_t)(a.x >> 16); m[2] = (int16_t)(a.y & 0xffff); m[3] = (int16_t)(a.y >> 16);
      m[4] = (int16_t)(b.x & 0xffff); m[5] = (int16_t)(b.x >> 16); m[6] = (int16_t)(b.y & 0xffff); m[7] = (int16_t)(b.y >> 16);
    } else {
#pragma unroll
      for (int k = 0; k < 8; k++) m[k] = p[(ptrdiff_t)(k - 4) * off];
    }
  }
  const int dpl = abs(m[1] - 2 * m[2] + m[3]), dql = abs(m[4] - 2 * m[5] + m[6]);       // xCalcDP / xCalcDQ of this line
  const int lane = threadIdx.x & 31, base = lane & ~3;
  const int dp0 = __shfl_sync(0xffffffffu, dpl, base), dq0 = __shfl_sync(0xffffffffu, dql, base);
  const int dp3 = __shfl_sync(0xffffffffu, dpl, base + 3), dq3 = __shfl_sync(0xffffffffu, dql, base + 3);
  // xUseStrongFiltering needs samples of lines 0 and 3 as well
  const int s_l = abs(m[0] - m[3]) + abs(m[7] - m[4]), g_l = abs(m[3] - m[4]);
  const int s0 = __shfl_sync(0xffffffffu, s_l, base), g0 = __shfl_sync(0xffffffffu, g_l, base);
  const int s3 = __shfl_sync(0xffffffffu, s_l, base + 3), g3 = __shfl_sync(0xffffffffu, g_l, base + 3);
  if (live) {
    const int qp = u.qp;
    const int idx_tc = clip3i(0, 51 + 2, qp + 2 * ((int)u.bs - 1) + (tc_off2 << 1));
    const int idx_b = clip3i(0, 51, qp + (beta_off2 << 1));
    const int tc = c_tc[idx_tc] * scale, beta = c_beta[idx_b] * scale;
    const int side = (beta + (beta >> 1)) >> 3, thr_cut = tc * 10;
    const int d0 = dp0 + dq0, d3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3, d = d0 + d3;
    if (d < beta) {
      const bool fp = dp < side, fq = dq < side;
      const bool sw = (s0 < (beta >> 3)) && (2 * d0 < (beta >> 2)) && (g0 < ((tc * 5 + 1) >> 1)) &&
                      (s3 < (beta >> 3)) && (2 * d3 < (beta >> 2)) && (g3 < ((tc * 5 + 1) >> 1));
      int o[8];
#pragma unroll
      for (int k = 0; k < 8; k++) o[k] = m[k];
      if (sw) {
        o[3] = clip3i(m[3] - 2 * tc, m[3] + 2 * tc, (m[1] + 2 * m[2] + 2 * m[3] + 2 * m[4] + m[5] + 4) >> 3);
        o[4] = clip3i(m[4] - 2 * tc, m[4] + 2 * tc, (m[2] + 2 * m[3] + 2 * m[4] + 2 * m[5] + m[6] + 4) >> 3);
        o[2] = clip3i(m[2] - 2 * tc, m[2] + 2 * tc, (m[1] + m[2] + m[3] + m[4] + 2) >> 2);
        o[5] = clip3i(m[5] - 2 * tc, m[5] + 2 * tc, (m[3] + m[4] + m[5] + m[6] + 2) >> 2);
        o[1] = clip3i(m[1] - 2 * tc, m[1] + 2 * tc, (2 * m[0] + 3 * m[1] + m[2] + m[3] + m[4] + 4) >> 3);
        o[6] = clip3i(m[6] - 2 * tc, m[6] + 2 * tc, (m[3] + m[4] + m[5] + 3 * m[6] + 2 * m[7] + 4) >> 3);
      } else {
        int delta = (9 * (m[4] - m[3]) - 3 * (m[5] - m[2]) + 8) >> 4;
        if (abs(delta) < thr_cut) {
          delta = clip3i(-tc, tc, delta);
          o[3] = clip3i(0, maxv, m[3] + delta);
          o[4] = clip3i(0, maxv, m[4] - delta);
          const int tc2 = tc >> 1;
          if (fp) o[2] = clip3i(0, maxv, m[2] + clip3i(-tc2, tc2, ((((m[1] + m[3] + 1) >> 1) - m[2] + delta) >> 1)));
          if (fq) o[5] = clip3i(0, maxv, m[5] + clip3i(-tc2, tc2, ((((m[6] + m[4] + 1) >> 1) - m[5] - delta) >> 1)));
        }
      }
      if (p_keep) { o[3] = m[3]; o[2] = m[2]; o[1] = m[1]; }
      if (q_keep) { o[4] = m[4]; o[5] = m[5]; o[6] = m[6]; }
      if (DIR == 0) {
        uint2 a, b;
        a.x = (uint32_t)(uint16_t)o[0] | ((uint32_t)(uint16_t)o[1] << 16); a.y = (uint32_t)(uint16_t)o[2] | ((uint32_t)(uint16_t)o[3] << 16);
        b.x = (uint32_t)(uint16_t)o[4] | ((uint32_t)(uint16_t)o[5] << 16); b.y = (uint32_t)(uint16_t)o[6] | ((uint32_t)(uint16_t)o[7] << 16);
        *reinterpret_cast<uint2*>(p - 4) = a;
        *reinterpret_cast<uint2*>(p) = b;
      } else {
#pragma unroll
        for (int k = 1; k < 7; k++) p[(ptrdiff_t)(k - 4) * off] = (int16_t)o[k];
      }
    }
  }

  // ---- chroma (xEdgeFilterChroma): edges on the 8-pel chroma grid = 16-pel luma grid, bs > 1 only; the unit covers 2
  // chroma lines; thread `line` takes plane line >> 1, chroma line line & 1
  const bool c_edge = DIR == 0 ? ((ux & 1) == 0) : ((uy & 1) == 0);
  if (in_range && u.bs > 1 && c_edge) {
    const int cl = line & 1;
    const int cx = DIR == 0 ? ux * 4 : ux * 2 + cl;
    const int cy = DIR == 0 ? uy * 2 + cl : uy * 4;
    if (cx < (width >> 1) && cy < (height >> 1) && (DIR == 0 ? cx > 0 : cy > 0)) {
      int16_t* c = ((line >> 1) ? V : U) + (ptrdiff_t)cy * sc + cx;
      const int coff = DIR == 0 ? 1 : sc;
      const int qpc = c_chroma_scale[clip3i(0, 51, (int)u.qp)];
      const int idx_tc = clip3i(0, 51 + 2, qpc + 2 * ((int)u.bs - 1) + (tc_off2 << 1));
      const int tc = c_tc[idx_tc] * scale;
      const int m4 = c[0], m3 = c[-coff], m5 = c[coff], m2 = c[-2 * coff];
      const int delta = clip3i(-tc, tc, ((((m4 - m3) << 2) + m2 - m5 + 4) >> 3));
      if (!p_keep) c[-coff] = (int16_t)clip3i(0, maxv, m3 + delta);
      if (!q_keep) c[0] = (int16_t)clip3i(0, maxv, m4 - delta);
    }
  }
}

// ---------------------------------------------------------------------------------------------- SAO apply
// processSaoCuOrg (TComSampleAdaptiveOffset.cpp:781-1003) as a pure function of the deblocked picture: one thread per
// sample, the CTU's record read through the read-only cache.  Picture-border samples keep their value for the classes
// whose neighbour would lie outside (iStartX / iEndX / iStartY / iEndY, :848-849, 866-867, 891-895, 929-933).
__global__ void __launch_bounds__(256)
k_sao_plane(const int16_t* __restrict__ src, int16_t* __restrict__ dst, int stride, int w, int h, int ctu, int ctus_x,
            const tvc_sao_unit* __restrict__ units, int bd)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const tvc_sao_unit* u = units + (y / ctu) * ctus_x + (x / ctu);
  const ptrdiff_t o = (ptrdiff_t)y * stride + x;
  const int c = src[o];
  const int type = u->type;
  int v = c;
  if (type == 4) {
    v = c + u->bo[c >> (bd - 5)];
  } else if (type >= 0) {
    const int dx = type == 1 ? 0 : 1;                        // neighbour a = (x - dx, y - dy), b = (x + dx, y + dy)
    const int dy = type == 0 ? 0 : (type == 3 ? -1 : 1);     // 45 deg: a = (x - 1, y + 1), b = (x + 1, y - 1)
    const bool ok = (dx == 0 || (x > 0 && x < w - 1)) && (dy == 0 || (y > 0 && y < h - 1));
    if (ok) {
      const int a = src[o - dx - (ptrdiff_t)dy * stride], b = src[o + dx + (ptrdiff_t)dy * stride];
      const int e = (c > a) - (c < a) + (c > b) - (c < b) + 2;
      v = c + u->eo[e];
    }
  }
  const int maxv = (1 << bd) - 1;
  dst[o] = (int16_t)(v < 0 ? 0 : (v > maxv ? maxv : v));
}

}  // namespace tvc

using namespace tvc;

extern "C" int tvc_sao_plane(tvc_ctx* c, int src_slot, int dst_slot, int plane, const tvc_sao_unit* units)
{
  if (!c || !valid_slot(c, src_slot) || !valid_slot(c, dst_slot) || src_slot == dst_slot || plane < 0 || plane > 2 || !units)
    return set_err(c, TVC_ERR_ARG, "tvc_sao_plane: bad argument (source and destination must be different slots)");
  const Pic& s = c->pics[src_slot];
  const Pic& d = c->pics[dst_slot];
  const int n = c->num_ctus_x * c->num_ctus_y;
  int r;
  if ((r = ensure_scratch(c, c->in, (size_t)n * sizeof(tvc_sao_unit)))) return r;
  memcpy(c->in.host, units, (size_t)n * sizeof(tvc_sao_unit));
  TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, c->in.host, (size_t)n * sizeof(tvc_sao_unit), cudaMemcpyHostToDevice, c->stream));
  const int w = s.w[plane], h = s.h[plane], ctu = c->cfg.max_cu >> (plane ? 1 : 0);
  dim3 blk(64, 4), grd((w + 63) / 64, (h + 3) / 4);
  ProfScope ps(c, TVC_PH_DEBLOCK);
  k_sao_plane<<<grd, blk, 0, c->stream>>>(s.org[plane], d.org[plane], s.stride[plane], w, h, ctu, c->num_ctus_x, (const tvc_sao_unit*)c->in.dev,
                                         c->cfg.bit_depth);
  TVC_LAUNCH_CHECK(c);
  return TVC_OK;
}

extern "C" int tvc_deblock_pic(tvc_ctx* c, int slot, const tvc_dbk_unit* ver, const tvc_dbk_unit* hor, int beta_offset_div2, int tc_offset_div2)
{
  if (!c || !valid_slot(c, slot) || beta_offset_div2 < -13 || beta_offset_div2 > 13 || tc_offset_div2 < -13 || tc_offset_div2 > 13)
    return set_err(c, TVC_ERR_ARG, "tvc_deblock_pic: bad argument");
  const Pic& p = c->pics[slot];
  const int W = p.w[0], H = p.h[0];
  const int vw = (W + 7) >> 3, vh = (H + 3) >> 2, hw = (W + 3) >> 2, hh = (H + 7) >> 3;
  const size_t nv = ver ? (size_t)vw * vh : 0, nh = hor ? (size_t)hw * hh : 0;
  if (nv + nh == 0) return TVC_OK;
  int r;
  if ((r = ensure_scratch(c, c->in, (nv + nh) * sizeof(tvc_dbk_unit)))) return r;
  if (nv) memcpy(c->in.host, ver, nv * sizeof(tvc_dbk_unit));
  if (nh) memcpy((char*)c->in.host + nv * sizeof(tvc_dbk_unit), hor, nh * sizeof(tvc_dbk_unit));
  TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, c->in.host, (nv + nh) * sizeof(tvc_dbk_unit), cudaMemcpyHostToDevice, c->stream));
  const tvc_dbk_unit* dv = (const tvc_dbk_unit*)c->in.dev;
  const tvc_dbk_unit* dh = dv + nv;
  ProfScope ps(c, TVC_PH_DEBLOCK);
  if (nv) {
    k_deblock<0><<<(unsigned)((nv * 4 + 255) / 256), 256, 0, c->stream>>>(p.org[0], p.stride[0], p.org[1], p.org[2], p.stride[1], W, H, dv, vw, vh,
                                                                         beta_offset_div2, tc_offset_div2, c->cfg.bit_depth);
    TVC_LAUNCH_CHECK(c);
  }
  if (nh) {
    k_deblock<1><<<(unsigned)((nh * 4 + 255) / 256), 256, 0, c->stream>>>(p.org[0], p.stride[0], p.org[1], p.org[2], p.stride[1], W, H, dh, hw, hh,
                                                                         beta_offset_div2, tc_offset_div2, c->cfg.bit_depth);
    TVC_LAUNCH_CHECK(c);
  }
  return TVC_OK;
}
