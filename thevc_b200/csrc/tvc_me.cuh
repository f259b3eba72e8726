// tvc_me.cuh -- device helpers shared by the motion-estimation kernels (tvc_me.cu, tvc_me_group.cu): TMA / mbarrier wrappers,
// the SIMD absolute-difference step, the MV rate of TComRdCost::getCost and the candidate geometry of the TZ diamond.
#pragma once
#include "tvc_internal.cuh"

namespace tvc {
struct MeMaps {
  CUtensorMap cur;
  CUtensorMap ref[8];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar)
{
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ uint32_t vsad4_acc(uint32_t a, uint32_t b, uint32_t c)
{
  uint32_t d;
  asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}


__device__ __forceinline__ uint32_t mv_comp_bits(int v)
{
  // xGetComponentBits (TComRdCost.cpp:270-284): 2*floor(log2(t)) + 1 with t = v<=0 ? -2v+1 : 2v
  uint32_t t = (v <= 0) ? (uint32_t)((-v << 1) + 1) : (uint32_t)(v << 1);
  return 2u * (31u - (uint32_t)__clz(t)) + 1u;
}
__device__ __forceinline__ uint32_t mv_cost(uint32_t lc, int x, int y, int scale, int px, int py)
{
  uint32_t bits = mv_comp_bits((x << scale) - px) + mv_comp_bits((y << scale) - py);
  return (lc * bits) >> 16;
}

constexpr uint32_t kNoCost = 0xFFFFFFFFu;

__device__ __forceinline__ int round_size(int d) { return d == 1 ? 4 : (d <= 8 ? 8 : 16); }

// i-th candidate (visiting order) of xTZ8PointDiamondSearch (TEncSearch.cpp:535-707) around (sx,sy) at
// distance d, with the reference's own border tests: an axis point (top / left / right / bottom at
// distance d) is tested against the one edge it can cross, an off-axis point is taken when the four
// axis points are inside ("check border") or else when it passes the vertical and the horizontal edge
// on its own side.  The centre CAN be outside the window (the zero vector is probed unconditionally,
// :4336-4339), so these are not the same as "point in window" and are restated as written.
// Branch-free: every lane of a batch holds a different i.
//   d == 1 : 4 points  top(2) left(4) right(5) bottom(7)
//   d <= 8 : 8 points  top(2) TL(1) TR(3) left(4) right(5) BL(6) BR(8) bottom(7); diagonals at d/2, tagged d/2
//   d  > 8 : 16 points top left right bottom, then k = 1..3: (xl,yt) (xr,yt) (xl,yb) (xr,yb), all tagged 0 / d
template <class WinT>      // any type with the window members lx, ty, rx, by
__device__ __forceinline__ bool diamond_cand(const WinT& s, int sx, int sy, int d, int i, int& x, int& y, int& pt,
                                             uint32_t& dist)
{
  int ux, uy, unit, ptn;          // offset = (ux, uy) * unit
  bool axis;
  if (d == 1) {
    // i: 0 top, 1 left, 2 right, 3 bottom
    ux = (i == 1) ? -1 : (i == 2 ? 1 : 0);
    uy = (i == 0) ? -1 : (i == 3 ? 1 : 0);
    unit = 1; axis = true;
    ptn = (0x7542 >> (4 * i)) & 15;
    dist = 1;
  } else if (d <= 8) {
    // nibble tables indexed by i (LSB first): ux+2, uy+2 in half-distance units, point number
    ux = (int)((0x23140312u >> (4 * i)) & 15) - 2;     // 0,-1,+1,-2,+2,-1,+1,0
    uy = (int)((0x43322110u >> (4 * i)) & 15) - 2;     // -2,-1,-1,0,0,+1,+1,+2
    ptn = (int)((0x78654312u >> (4 * i)) & 15);        // 2,1,3,4,5,6,8,7
    unit = d >> 1;
    axis = (ux == 0) || (uy == 0);
    dist = axis ? (uint32_t)d : (uint32_t)(d >> 1);
  } else {
    unit = d >> 2;
    if (i < 4) {
      ux = (i == 1) ? -4 : (i == 2 ? 4 : 0);
      uy = (i == 0) ? -4 : (i == 3 ? 4 : 0);
      axis = true;
    } else {
      const int k = ((i - 4) >> 2) + 1, j = (i - 4) & 3;
      ux = (j & 1) ? k : -k;
      uy = (j & 2) ? 4 - k : k - 4;
      axis = false;
    }
    ptn = 0; dist = (uint32_t)d;
  }
  x = sx + ux * unit; y = sy + uy * unit; pt = ptn;
  const bool yc = uy < 0 ? (y >= s.ty) : (y <= s.by);
  const bool xc = ux < 0 ? (x >= s.lx) : (x <= s.rx);
  const bool inside = (sy - d) >= s.ty && (sy + d) <= s.by && (sx - d) >= s.lx && (sx + d) <= s.rx;
  return axis ? (ux == 0 ? yc : xc) : (inside || (yc && xc));
}

// candidate c (visiting order) of a sweep that starts at distance 1 -> (round distance d, index i inside the round): the rounds hold
// 4, 8, 8, 8, 16, 16, 16 candidates (d = 1 .. 64), i.e. start at 0, 4, 12, 20, 28, 44, 60.  Closed form instead of walking the
// rounds: the walk was 20 % of k_me_search's instructions in the ncu source view.
__device__ __forceinline__ bool sweep_slot(int c, int dmax, int& d, int& i)
{
  if (c < 4) { d = 1; i = c; }
  else if (c < 28) { d = 2 << ((c - 4) >> 3); i = (c - 4) & 7; }
  else if (c < 76) { d = 16 << ((c - 28) >> 4); i = (c - 28) & 15; }
  else { d = 128; i = 0; return false; }
  return d <= dmax;
}


}  // namespace tvc
