// tvc_me.cu -- motion estimation on the device.
//
//  (1) k_me_sad_tables : per-frame pre-pass.  One CTA per (CTU, reference): the 192x192 u8 search
//      window and the 64x64 u8 current CTU are staged into shared memory with TMA
//      (cp.async.bulk.tensor.2d + mbarrier), then every warp walks tasks of 8 dx (equal dx mod 16,
//      so the byte alignment is warp-uniform) x 4 column quarters x 8 dy, sliding an 8-row register
//      window over the reference and accumulating VABSDIFF4.U8.ACC into even-row / odd-row sums of
//      every 4x4 block (TComRdCost.cpp:518-989 with the FEN row sub-sampling of
//      TEncSearch.cpp:324-330).  Output: uint16 tables, layout in tvc_internal.cuh.
//  (2) k_me_search     : xPatternSearch / xTZSearch (TEncSearch.cpp:4227-4474) per PU job, one warp
//      per job and one lane per candidate of a diamond round / raster batch, SADs from the tables
//      (or from the pictures), MV rate added per candidate (TComRdCost.h:196-213); the reference's
//      sequential strict-'<' update is replayed as an ordered arg-min per round.
//  (3) k_me_frac       : xPatternSearchFracDIF (TEncSearch.cpp:4476-4514): 8-tap half/quarter
//      interpolation around the integer MV and 9+9 Hadamard SATD evaluations, one CTA per job,
//      14-bit intermediates in shared memory, Hadamard butterflies in registers + warp shuffles.
#include "tvc_internal.cuh"
#include "tvc_me.cuh"
#include "tvc_interp.cuh"
#include <stdlib.h>

namespace tvc {

// ================================================================================ (1) SAD tables
// TMA needs a 16-byte aligned start coordinate: the window is loaded from its start rounded down to
// 16, 16 bytes wider, and every task adds the remainder e = start & 15 to its byte offset.
constexpr int kWinW = 208, kWinH = 192;
constexpr int kSmemWin = kWinW * kWinH;          // 36864
constexpr int kSmemCur = 64 * 64;                // 4096
constexpr int kSmemTables = kSmemWin + kSmemCur + 64;

// one warp task.  Lane parameters: X = byte offset of the lane's 16-byte cur quarter inside a window
// row for its dx (X = dx+64 + 16q, X & 15 == align16 for all lanes), dyb = first dy (0-based, i.e.
// dy+64) of its DYB consecutive dy values.  WO = (align16 >> 2) is the word offset inside the
// 16-byte chunk, SH = (align16 & 3) * 8 the funnel shift; both warp-uniform.
template <int DYB, int WO>
__device__ __forceinline__ void sad_task(const uint8_t* __restrict__ win, const uint8_t* __restrict__ cur, int X, int SH,
                                         int dyb, int q, bool active, uint16_t* __restrict__ tbl, int dxc)
{
  const uint8_t* wbase = win + (X & ~15);
  const uint8_t* cbase = cur + 16 * q;
  uint32_t R[DYB][4];      // aligned reference words of DYB consecutive window rows (slot = row & (DYB-1))
  auto load_row = [&](int row, uint32_t* dst) {
    const uint4* p = reinterpret_cast<const uint4*>(wbase + row * kWinW);
    uint4 lo = p[0], hi = p[1];
    uint32_t W[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int m = 0; m < 4; m++) dst[m] = __funnelshift_r(W[WO + m], W[WO + m + 1], SH);
  };
#pragma unroll
  for (int d = 0; d < DYB - 1; d++) load_row(dyb + d, R[d]);

  uint32_t E[DYB][4], O[DYB][4];
  for (int y0 = 0; y0 < 64; y0 += (DYB >= 4 ? DYB : 4)) {
#pragma unroll
    for (int yy = 0; yy < (DYB >= 4 ? DYB : 4); yy++) {
      const int y = y0 + yy;
      load_row(dyb + y + DYB - 1, R[(yy + DYB - 1) % DYB]);
      const uint4 c4 = *reinterpret_cast<const uint4*>(cbase + y * 64);
      const uint32_t C[4] = {c4.x, c4.y, c4.z, c4.w};
      const int r = yy & 3;
#pragma unroll
      for (int d = 0; d < DYB; d++) {
        const uint32_t* rr = R[(yy + d) % DYB];
#pragma unroll
        for (int m = 0; m < 4; m++) {
          if (r == 0) E[d][m] = vsad4_acc(C[m], rr[m], 0u);
          else if (r == 1) O[d][m] = vsad4_acc(C[m], rr[m], 0u);
          else if (r == 2) E[d][m] = vsad4_acc(C[m], rr[m], E[d][m]);
          else O[d][m] = vsad4_acc(C[m], rr[m], O[d][m]);
        }
      }
      if (r == 3 && active) {
        const int by = y >> 2;
#pragma unroll
        for (int d = 0; d < DYB; d++) {
          uint4 o;
          o.x = E[d][0] | (E[d][1] << 16);
          o.y = E[d][2] | (E[d][3] << 16);
          o.z = O[d][0] | (O[d][1] << 16);
          o.w = O[d][2] | (O[d][3] << 16);
          size_t e = ((size_t)((dyb + d) * kMeC + dxc) * kMeCandGranules + by * 4 + q) * kMeGranule;
          *reinterpret_cast<uint4*>(tbl + e) = o;
        }
      }
    }
  }
}

// Row-split form of the same task: lane = (i2: 2 dx, yb: 4 block rows, q: 4 quarters).  At step t the four yb lanes of a dx hold
// block rows 4t .. 4t+3 of the same candidates, so one warp store covers 256 contiguous bytes per candidate (two candidates per
// instruction) instead of eight 64-byte pieces: measured with tvc_ubench, 64-byte pieces reach 5.5 TB/s, 256-byte runs 6.8 TB/s.
// The price is the sliding window: every step reloads its DYB+3 window rows (26 instead of 12 shared-memory loads per 4 rows).
template <int DYB, int WO>
__device__ __forceinline__ void sad_task_rows(const uint8_t* __restrict__ win, const uint8_t* __restrict__ cur, int X, int SH,
                                              int dyb, int q, int yb, uint16_t* __restrict__ tbl, int dxc)
{
  const uint8_t* wbase = win + (X & ~15);
  const uint8_t* cbase = cur + 16 * q;
  for (int t = 0; t < 4; t++) {
    const int by = 4 * t + yb, y0 = 4 * by;
    uint32_t C[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const uint4 c4 = *reinterpret_cast<const uint4*>(cbase + (y0 + r) * 64);
      C[r][0] = c4.x; C[r][1] = c4.y; C[r][2] = c4.z; C[r][3] = c4.w;
    }
    uint32_t E[DYB][4], O[DYB][4];
#pragma unroll
    for (int j = 0; j < DYB + 3; j++) {
      // window row dyb + y0 + j meets current row r at dy offset d = j - r
      const uint4* p = reinterpret_cast<const uint4*>(wbase + (dyb + y0 + j) * kWinW);
      const uint4 lo = p[0], hi = p[1];
      const uint32_t W[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
      uint32_t R[4];
#pragma unroll
      for (int m = 0; m < 4; m++) R[m] = __funnelshift_r(W[WO + m], W[WO + m + 1], SH);
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const int d = j - r;
        if (d >= 0 && d < DYB) {
#pragma unroll
          for (int m = 0; m < 4; m++) {
            if (r == 0) E[d][m] = vsad4_acc(C[0][m], R[m], 0u);
            else if (r == 1) O[d][m] = vsad4_acc(C[1][m], R[m], 0u);
            else if (r == 2) E[d][m] = vsad4_acc(C[2][m], R[m], E[d][m]);
            else O[d][m] = vsad4_acc(C[3][m], R[m], O[d][m]);
          }
        }
      }
    }
#pragma unroll
    for (int d = 0; d < DYB; d++) {
      uint4 o;
      o.x = E[d][0] | (E[d][1] << 16);
      o.y = E[d][2] | (E[d][3] << 16);
      o.z = O[d][0] | (O[d][1] << 16);
      o.w = O[d][2] | (O[d][3] << 16);
      const size_t e = ((size_t)((dyb + d) * kMeC + dxc) * kMeCandGranules + by * 4 + q) * kMeGranule;
      *reinterpret_cast<uint4*>(tbl + e) = o;
    }
  }
}

template <int DYB>
__device__ __forceinline__ void sad_task_rows_wo(int wo, const uint8_t* win, const uint8_t* cur, int X, int SH, int dyb, int q, int yb,
                                                 uint16_t* tbl, int dxc)
{
  switch (wo) {
    case 0: sad_task_rows<DYB, 0>(win, cur, X, SH, dyb, q, yb, tbl, dxc); break;
    case 1: sad_task_rows<DYB, 1>(win, cur, X, SH, dyb, q, yb, tbl, dxc); break;
    case 2: sad_task_rows<DYB, 2>(win, cur, X, SH, dyb, q, yb, tbl, dxc); break;
    default: sad_task_rows<DYB, 3>(win, cur, X, SH, dyb, q, yb, tbl, dxc); break;
  }
}

template <int DYB>
__device__ __forceinline__ void sad_task_wo(int wo, const uint8_t* win, const uint8_t* cur, int X, int SH, int dyb, int q,
                                            bool active, uint16_t* tbl, int dxc)
{
  switch (wo) {
    case 0: sad_task<DYB, 0>(win, cur, X, SH, dyb, q, active, tbl, dxc); break;
    case 1: sad_task<DYB, 1>(win, cur, X, SH, dyb, q, active, tbl, dxc); break;
    case 2: sad_task<DYB, 2>(win, cur, X, SH, dyb, q, active, tbl, dxc); break;
    default: sad_task<DYB, 3>(win, cur, X, SH, dyb, q, active, tbl, dxc); break;
  }
}

template <int DYB, int MINB, bool ROWS = false>   // DYB consecutive dy per lane (register window), MINB resident CTAs per SM compiled
__global__ void __launch_bounds__(256, MINB)       // for, ROWS: row-split lane mapping of the bulk tasks (256-byte store runs)
k_me_sad_tables(const __grid_constant__ MeMaps maps, int num_ctus, int ctus_x, int mx, int my,
                const tvc_me_center* __restrict__ centers, uint16_t* __restrict__ tables, int ctu0, int ref0)
{
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* win = smem;
  uint8_t* cur = smem + kSmemWin;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kSmemWin + kSmemCur);
  const int ctu = blockIdx.x + ctu0, ref = blockIdx.y + ref0;     // (ctu0, ref0): first unit of a pipelined chunk
  const int cx0 = (ctu % ctus_x) * 64, cy0 = (ctu / ctus_x) * 64;
  const tvc_me_center cen = centers[(size_t)ref * num_ctus + ctu];

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int wx = mx + cx0 + cen.cx - kMeR;
  const int e = wx & 15;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, kSmemWin + kSmemCur);
    tma_load_2d(win, &maps.ref[ref], wx - e, my + cy0 + cen.cy - kMeR, bar);
    tma_load_2d(cur, &maps.cur, mx + cx0, my + cy0, bar);
  }
  mbar_wait(bar, 0);

  uint16_t* tbl = tables + ((size_t)ref * num_ctus + ctu) * kMeCtuElems;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = lane & 3, i8 = lane >> 2;
  // tasks: [0,16*NG) (a,g) 8 dx x DYB dy ; then 16 (a) last dy row ; then NG/8 for the dx=+64 column (8 dy-groups per
  // warp) ; last the (+64,+64) corner.  NG = 128 / DYB dy-groups.
  constexpr int NG = 128 / DYB, T0 = 16 * NG, T1 = T0 + 16, T2 = T1 + NG / 8;
  if (ROWS) {
    // bulk of the table (dx, dy < +64) in the row-split mapping: task = (a: dx residue mod 16, xb: 32-wide dx band, g: dy group)
    const int i2 = lane >> 4, yb = (lane >> 2) & 3;
    for (int t = warp; t < 64 * NG; t += 8) {
      const int a = t & 15, xb = (t >> 4) & 3, g = t >> 6;
      const int u = a + 16 * (2 * xb + i2), al = (a + e) & 15;
      sad_task_rows_wo<DYB>(al >> 2, win, cur, u + 16 * q + e, (al & 3) * 8, g * DYB, q, yb, tbl, u);
    }
  }
  for (int t = warp + (ROWS ? T0 : 0); t <= T2; t += 8) {
    if (t < T0) {
      int a = t & 15, g = t >> 4;
      int u = a + 16 * i8, al = (a + e) & 15;
      sad_task_wo<DYB>(al >> 2, win, cur, u + 16 * q + e, (al & 3) * 8, g * DYB, q, true, tbl, u);
    } else if (t < T1) {
      int a = t - T0;
      int u = a + 16 * i8, al = (a + e) & 15;
      sad_task_wo<1>(al >> 2, win, cur, u + 16 * q + e, (al & 3) * 8, 128, q, true, tbl, u);
    } else if (t < T2) {
      int g = (t - T1) * 8 + i8;
      sad_task_wo<DYB>(e >> 2, win, cur, 128 + 16 * q + e, (e & 3) * 8, g * DYB, q, true, tbl, 128);
    } else {
      sad_task_wo<1>(e >> 2, win, cur, 128 + 16 * q + e, (e & 3) * 8, 128, q, i8 == 0, tbl, 128);
    }
  }
}

// ---- table read: SAD of a PU (inside one CTU) at candidate (dx,dy) relative to the table centre.
// Cooperative form: lanes split the (block-row, quarter) granules; caller reduces over the warp.
__device__ __forceinline__ uint32_t table_pu_sad_partial(const uint16_t* __restrict__ tbl, int bx0, int by0, int nbx,
                                                         int nby, bool even_only, int dy, int dx, int lane)
{
  // granules: by in [by0, by0+nby), q in [bx0/4, (bx0+nbx-1)/4]
  const int q0 = bx0 >> 2, q1 = (bx0 + nbx - 1) >> 2, nq = q1 - q0 + 1;
  const int total = nby * nq;
  uint32_t acc = 0;
  for (int i = lane; i < total; i += 32) {
    int by = by0 + i / nq, q = q0 + i % nq;
    const uint4 g = *reinterpret_cast<const uint4*>(tbl + (size_t)me_granule(dy, dx, by, q) * kMeGranule);
    uint32_t ev[4] = {g.x & 0xffffu, g.x >> 16, g.y & 0xffffu, g.y >> 16};
    uint32_t od[4] = {g.z & 0xffffu, g.z >> 16, g.w & 0xffffu, g.w >> 16};
#pragma unroll
    for (int m = 0; m < 4; m++) {
      int bx = q * 4 + m;
      if (bx >= bx0 && bx < bx0 + nbx) acc += even_only ? ev[m] : (ev[m] + od[m]);
    }
  }
  return acc;
}

__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__constant__ tvc_census_pu c_census[TVC_ME_CENSUS];     // partition census of a CTU (frame pre-pass)

// ================================================================================ (2) search
// One warp per PU job, ONE LANE PER CANDIDATE.  The reference evaluates candidates one after the
// other (xTZSearchHelp, TEncSearch.cpp:312-349: strict '<' against the running best), but inside a
// diamond round, a raster scan or the 2-point step the candidate set is fixed beforehand, so all
// costs of a batch are computed in parallel (each lane sums its candidate's table granules, many
// independent 16-byte loads in flight) and the sequential update is replayed afterwards as an
// ordered arg-min per round: the first candidate in visiting order that attains the round minimum
// wins iff it is strictly below the best so far -- exactly the state the sequential loop ends in.
#ifndef TVC_SEARCH_ROW_UNROLL
#define TVC_SEARCH_ROW_UNROLL 1
#endif
#ifndef TVC_RASTER_K
#define TVC_RASTER_K 4
#endif
constexpr int kSearchRowUnroll = TVC_SEARCH_ROW_UNROLL;

// the staged search window of a (CTU, reference) group (k_me_group)
struct GrpWin {
  const uint8_t* win;           // shared memory: 208 x 192 u8, window of +-64 around (cenx, ceny), e16 bytes of alignment slack on the left
  const uint8_t* cur;           // shared memory: the CTU, pitch kGrpCurP
  const uint8_t* plane8;        // the reference's u8 luma plane, pel (0,0)
  int pitch8, cenx, ceny, e16, x0, y0;
};

struct SearchCtx {
  // SAD sources
  const uint16_t* tbl;      // table of (ref, ctu) or nullptr
  int tcx, tcy;             // table centre
  int by0, nby, q0, nq;     // PU rows in 4x4 blocks / 16-column quarters it touches
  uint32_t m01[4], m23[4];  // per touched quarter: packed-u16 masks of the blocks inside the PU
  const int16_t* org; int so;
  const int16_t* ref; int rs;   // co-located pel of the PU in the reference plane
  // group search (k_me_group): u8 search window + CTU staged in shared memory, SADs computed per lane on demand
  const uint8_t* win8;          // nullptr: not in this mode.  Else: window byte of the PU's top-left at candidate (wcx, wcy) - (64, 64)
  const uint8_t* cur8;          // the PU's top-left in the staged CTU (pitch kGrpCurP)
  const uint8_t* gref8;         // co-located byte of the PU in the reference's u8 plane (candidates beyond the window)
  int gpitch8, wcx, wcy;
  bool grp;                     // group search; a compile-time constant wherever me_search_job is inlined (the other form's code drops out)
  int w, h, sub, bi;
  uint32_t lc; int px, py;
  int lx, ty, rx, by;
  // state (IntTZSearchStruct)
  uint32_t best_sad; int best_x, best_y; uint32_t best_dist, best_round; int point_nr;
  uint32_t n_sads;
  int lane;
};

// whole-warp SAD of one candidate straight from the pictures (candidate outside the table window,
// no tables, 10-bit): xGetSAD* with iSubShift (TComRdCost.cpp:518-989)
__device__ __noinline__ uint32_t direct_sad_warp(const int16_t* __restrict__ org, int so, const int16_t* __restrict__ c, int rs,
                                                 int w, int h, int sub, int lane)
{
  int step = 1 << sub, nrows = h >> sub, total = w * nrows;
  uint32_t acc = 0;
  for (int i = lane; i < total; i += 32) {
    int r = i / w, xx = i - r * w;
    acc += (uint32_t)abs((int)org[(r * step) * so + xx] - (int)c[(ptrdiff_t)(r * step) * rs + xx]);
  }
  return warp_sum_u32(acc);
}

constexpr int kGrpWinW = 208, kGrpWinH = 192, kGrpCurP = 80;

// SAD of one candidate by ONE lane, 8-bit samples packed four to a word: the reference row is read as aligned words and shifted
// into place (the candidate's byte alignment differs from lane to lane), VABSDIFF4 accumulates four absolute differences per
// instruction.  rows = h >> sub rows at a row step of 1 << sub (TEncSearch.cpp:324-330); xGetSAD* of TComRdCost.cpp:518-989.
__device__ __forceinline__ uint32_t sad_u8_lane(const uint8_t* __restrict__ ref, int rpitch, const uint8_t* __restrict__ cur, int w4, int rows,
                                                int sub)
{
  // measured on B200 (ms of k_me_group per 1080p picture x 4 references): this plain form 5.8; two rows per step with four
  // independent accumulators 6.8; row loop unrolled per PU width behind a non-inlined call 10.0 -- the simple loop stays
  const int a = (int)((uintptr_t)ref & 3), sh = a * 8;
  const uint32_t* rb = reinterpret_cast<const uint32_t*>(ref - a);
  const uint32_t* cb = reinterpret_cast<const uint32_t*>(cur);
  const int rstride = (rpitch << sub) >> 2, cstride = (kGrpCurP << sub) >> 2;
  uint32_t acc = 0;
  for (int r = 0; r < rows; r++) {
    uint32_t prev = rb[0];
#pragma unroll 4
    for (int j = 0; j < w4; j++) {
      const uint32_t nxt = rb[j + 1];
      acc = vsad4_acc(cb[j], __funnelshift_r(prev, nxt, sh), acc);
      prev = nxt;
    }
    rb += rstride; cb += cstride;
  }
  return acc;
}

// The same loop behind ONE call for every place of the group kernel that asks for a SAD.  k_me_group inlines the TZ control code of
// a PU once per call site (start / zero, the staged sweeps, the two-point step, raster, star refinement); with the SAD loop inlined
// into each of them the kernel was 318 KB of SASS, its eight warps sat in different copies of the same loop and ncu's top stall
// was "no instruction" (4.9 warp-cycles per issue, instruction-cache hit rate 65 %).  Operands are byte offsets into the CTA's
// dynamic shared memory so that the loads stay LDS (a pointer argument would make them generic).
#ifndef TVC_GRP_VAR
#define TVC_GRP_VAR 14           // bit 0: SAD loop behind a call, bit 1: compact CU cost routine behind a call, bit 2: the whole
                                 // candidate cost behind the call, bit 3: SAD loop per PU width.  Measured on B200, ms of k_me_group per
                                 // 1080p picture x 4 references (quarter-pel predictor guesses): 0: 3.18, 1: 3.24, 2: 2.96, 3: 3.14,
                                 // 6: 3.13, 11: 2.39, 14: 2.30 (then 1.93 with the statistics summed per warp); same results for all
#endif
// bit 3 of TVC_GRP_VAR: the SAD loop per PU width.  Most census PUs are narrow (320 of 593 belong to 8x8 CUs: one or two words
// per row), where the generic loop spends two of every seven instructions on its own control; the width is uniform over the
// warp, so one switch picks a row loop with the words unrolled, two rows per trip (census heights give even row counts), the
// current block read with the widest aligned load.
template <int W4>
__device__ __forceinline__ uint32_t sad_rows_w(const uint32_t* __restrict__ rb, const uint32_t* __restrict__ cb, int sh, int rows,
                                               int rstride, int cstride)
{
  constexpr int V = (W4 % 4 == 0) ? 4 : ((W4 % 2 == 0) ? 2 : 1);
  uint32_t acc[2] = {0, 0};
#pragma unroll 1
  for (int r = 0; r < rows; r += 2) {
#pragma unroll
    for (int t = 0; t < 2; t++) {
      const uint32_t* rr = rb + t * rstride;
      const uint32_t* cc = cb + t * cstride;
      uint32_t w[W4 + 1], c[W4];
#pragma unroll
      for (int j = 0; j <= W4; j++) w[j] = rr[j];
      if (V == 4) {
#pragma unroll
        for (int j = 0; j < W4; j += 4) {
          const uint4 q = *reinterpret_cast<const uint4*>(cc + j);
          c[j] = q.x; c[j + 1] = q.y; c[j + 2] = q.z; c[j + 3] = q.w;
        }
      } else if (V == 2) {
#pragma unroll
        for (int j = 0; j < W4; j += 2) {
          const uint2 q = *reinterpret_cast<const uint2*>(cc + j);
          c[j] = q.x; c[j + 1] = q.y;
        }
      } else {
#pragma unroll
        for (int j = 0; j < W4; j++) c[j] = cc[j];
      }
#pragma unroll
      for (int j = 0; j < W4; j++) acc[t] = vsad4_acc(c[j], __funnelshift_r(w[j], w[j + 1], sh), acc[t]);
    }
    rb += 2 * rstride; cb += 2 * cstride;
  }
  return acc[0] + acc[1];
}

__device__ __forceinline__ uint32_t sad_u8_smem_body(uint32_t ref_off, uint32_t cur_off, int w4, int rows, int sub)
{
  extern __shared__ __align__(128) uint8_t gsm[];
#if TVC_GRP_VAR & 8
  const int a = (int)(ref_off & 3u), sh = a * 8;
  const uint32_t* rb = reinterpret_cast<const uint32_t*>(gsm + (ref_off - a));
  const uint32_t* cb = reinterpret_cast<const uint32_t*>(gsm + cur_off);
  const int rstride = (kGrpWinW << sub) >> 2, cstride = (kGrpCurP << sub) >> 2;
  // vector loads of the current block need its alignment (census PUs have it: a part as wide as 16 / 8 pels starts on a multiple)
  const int need = (w4 & 3) == 0 ? 15 : ((w4 & 1) == 0 ? 7 : 3);
  const int sel = ((cur_off & need) == 0 && (rows & 1) == 0) ? w4 : 0;
  switch (sel) {
    case 1: return sad_rows_w<1>(rb, cb, sh, rows, rstride, cstride);
    case 2: return sad_rows_w<2>(rb, cb, sh, rows, rstride, cstride);
    case 3: return sad_rows_w<3>(rb, cb, sh, rows, rstride, cstride);
    case 4: return sad_rows_w<4>(rb, cb, sh, rows, rstride, cstride);
    case 6: return sad_rows_w<6>(rb, cb, sh, rows, rstride, cstride);
    case 8: return sad_rows_w<8>(rb, cb, sh, rows, rstride, cstride);
    case 12: return sad_rows_w<12>(rb, cb, sh, rows, rstride, cstride);
    case 16: return sad_rows_w<16>(rb, cb, sh, rows, rstride, cstride);
    default: break;
  }
#endif
  return sad_u8_lane(gsm + ref_off, kGrpWinW, gsm + cur_off, w4, rows, sub);
}

__device__ __noinline__ uint32_t sad_u8_smem(uint32_t ref_off, uint32_t cur_off, int w4, int rows, int sub)
{
  return sad_u8_smem_body(ref_off, cur_off, w4, rows, sub);
}
__device__ __noinline__ uint32_t sad_u8_plane(const uint8_t* __restrict__ ref, int rpitch, uint32_t cur_off, int w4, int rows, int sub)
{
  extern __shared__ __align__(128) uint8_t gsm[];
  return sad_u8_lane(ref, rpitch, gsm + cur_off, w4, rows, sub);
}

// bit 2 of TVC_GRP_VAR: the whole candidate cost (window test, address, SAD, MV rate) behind the call
__device__ __noinline__ uint32_t grp_cand_cost(uint32_t win_off, uint32_t cur_off, const uint8_t* __restrict__ gref8, int gpitch8, int dx, int dy,
                                               int x, int y, int w4_rows_sub, uint32_t lc, int px, int py)
{
  extern __shared__ __align__(128) uint8_t gsm[];
  const int w4 = w4_rows_sub & 0xff, rows = (w4_rows_sub >> 8) & 0xff, sub = w4_rows_sub >> 16;
  uint32_t sad;
  if (dx >= -kMeR && dx <= kMeR && dy >= -kMeR && dy <= kMeR)
    sad = sad_u8_smem_body(win_off + (uint32_t)((dy + kMeR) * kGrpWinW + dx + kMeR), cur_off, w4, rows, sub);
  else
    sad = sad_u8_lane(gref8 + (ptrdiff_t)y * gpitch8 + x, gpitch8, gsm + cur_off, w4, rows, sub);
  return (sad << sub) + mv_cost(lc, x, y, 2, px, py);
}

// cost (SAD + MV rate at scale 2) of this lane's K candidates; kNoCost for slots without one.  The K
// table sums run interleaved so that K * granules independent 16-byte loads are in flight per lane.
// Warp-collective: every lane must call it.
template <int K>
__device__ __forceinline__ void eval_multi(const SearchCtx& s, const bool (&valid)[K], const int (&x)[K], const int (&y)[K],
                                           uint32_t (&cost)[K])
{
  if (s.grp) {
    // group search: every lane computes the SAD of its own candidates from the staged window (or, beyond it, from the u8 plane)
    const int w4 = s.w >> 2, rows = s.h >> s.sub;
#pragma unroll
    for (int k = 0; k < K; k++) {
      cost[k] = kNoCost;
      if (valid[k]) {
        const int dx = x[k] - s.wcx, dy = y[k] - s.wcy;
#if TVC_GRP_VAR & 4
        extern __shared__ __align__(128) uint8_t gsm[];
        cost[k] = grp_cand_cost((uint32_t)(s.win8 - gsm), (uint32_t)(s.cur8 - gsm), s.gref8, s.gpitch8, dx, dy, x[k], y[k],
                                w4 | (rows << 8) | (s.sub << 16), s.lc, s.px, s.py);
#else
        uint32_t sad;
#if TVC_GRP_VAR & 1
        extern __shared__ __align__(128) uint8_t gsm[];
        if (dx >= -kMeR && dx <= kMeR && dy >= -kMeR && dy <= kMeR)
          sad = sad_u8_smem((uint32_t)(s.win8 - gsm) + (uint32_t)((dy + kMeR) * kGrpWinW + dx + kMeR), (uint32_t)(s.cur8 - gsm), w4, rows, s.sub);
        else
          sad = sad_u8_plane(s.gref8 + (ptrdiff_t)y[k] * s.gpitch8 + x[k], s.gpitch8, (uint32_t)(s.cur8 - gsm), w4, rows, s.sub);
#else
        if (dx >= -kMeR && dx <= kMeR && dy >= -kMeR && dy <= kMeR)
          sad = sad_u8_lane(s.win8 + (dy + kMeR) * kGrpWinW + dx + kMeR, kGrpWinW, s.cur8, w4, rows, s.sub);
        else
          sad = sad_u8_lane(s.gref8 + (ptrdiff_t)y[k] * s.gpitch8 + x[k], s.gpitch8, s.cur8, w4, rows, s.sub);
#endif
        cost[k] = (sad << s.sub) + mv_cost(s.lc, x[k], y[k], 2, s.px, s.py);
#endif
      }
    }
    return;
  }
  uint32_t sad[K];
  bool tab[K], direct[K];
  uint32_t base[K];             // granule index inside this (ref, CTU) table (1.06 M granules: fits 32 bits)
  bool any_direct = false;
#pragma unroll
  for (int k = 0; k < K; k++) {
    const int dx = x[k] - s.tcx, dy = y[k] - s.tcy;
    const bool cov = s.tbl && dx >= -kMeR && dx <= kMeR && dy >= -kMeR && dy <= kMeR;
    tab[k] = valid[k] && cov;
    direct[k] = valid[k] && !cov;
    any_direct |= direct[k];
    base[k] = tab[k] ? me_granule(dy, dx, s.by0, s.q0) : 0u;
    sad[k] = 0;
  }
  const uint4* __restrict__ tg = reinterpret_cast<const uint4*>(s.tbl);
  constexpr uint32_t kRow = 4u;                 // next block row, same quarter (granules)
  constexpr uint32_t kQtr = 1u;                 // next quarter, same block row
  const bool all = s.sub == 0;
#pragma unroll
  for (int qi = 0; qi < 4; qi++) {
    if (qi < s.nq) {
      const uint32_t ma = s.m01[qi], mb = s.m23[qi];
      // block rows in runs of <= 8: the packed u16 pairs of a run are summed as they are (a half stays <= 8 * 4 * 2040 = 65280)
      // and unpacked once per run instead of once per granule
      for (int r0 = 0; r0 < s.nby; r0 += 8) {
        const int r1 = min(r0 + 8, s.nby);
        uint32_t pk[K];
#pragma unroll
        for (int k = 0; k < K; k++) pk[k] = 0;
#pragma unroll kSearchRowUnroll
        for (int r = r0; r < r1; r++) {
          uint4 g[K];
#pragma unroll
          for (int k = 0; k < K; k++)
            if (tab[k]) g[k] = __ldg(tg + (base[k] + qi * kQtr + (uint32_t)r * kRow));
#pragma unroll
          for (int k = 0; k < K; k++) {
            if (tab[k]) {
              uint32_t v = (g[k].x & ma) + (g[k].y & mb);         // packed u16 pairs, each <= 2 * 2040
              if (all) v += (g[k].z & ma) + (g[k].w & mb);        // <= 4 * 2040
              pk[k] += v;
            }
          }
        }
#pragma unroll
        for (int k = 0; k < K; k++) sad[k] += (pk[k] & 0xffffu) + (pk[k] >> 16);
      }
    }
  }
  if (__any_sync(0xffffffffu, any_direct)) {
#pragma unroll
    for (int k = 0; k < K; k++) {
      unsigned m = __ballot_sync(0xffffffffu, direct[k]);
      while (m) {
        int src = __ffs(m) - 1;
        m &= m - 1;
        int cx = __shfl_sync(0xffffffffu, x[k], src), cy = __shfl_sync(0xffffffffu, y[k], src);
        uint32_t v = direct_sad_warp(s.org, s.so, s.ref + (ptrdiff_t)cy * s.rs + cx, s.rs, s.w, s.h, s.sub, s.lane);
        if (s.lane == src) sad[k] = v;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < K; k++)
    cost[k] = valid[k] ? ((sad[k] << s.sub) >> s.bi) + mv_cost(s.lc, x[k], y[k], 2, s.px, s.py) : kNoCost;
}

// Diamond rounds d = 1, 2, 4, ... <= dmax around (sx,sy) (d0 must be 1: every TZ sweep starts there).  All candidates of these rounds are known
// beforehand, so they are evaluated together: candidate c (visiting order) is slot c >> 5 of lane
// c & 31.  The reference's sequential update is then replayed as an ordered arg-min.
// first_search: the reference stops when three consecutive rounds brought no improvement
// (bFirstSearchStop, uiFirstSearchRounds = 3; TEncSearch.cpp:4346-4361): candidates of later rounds were
// evaluated speculatively and are neither counted nor used; returns true when stopped.
// First search of the group kernel, in two stages.  There a SAD is real arithmetic (not a table read), and the reference usually
// stops after three or four rounds (bFirstSearchStop): rounds d = 1 .. 8 (28 candidates, one per lane) are evaluated and replayed
// first, rounds d = 16 .. 64 (48 candidates, two per lane) only when the search is still running.  Same visiting order, same
// strict-'<' replay, same count as the one-shot form.
template <int K>
__device__ __forceinline__ bool sweep_stage(SearchCtx& s, int sx, int sy, int c0, int dfirst, int dlast, int dmax)
{
  bool valid[K];
  int x[K], y[K];
  uint32_t cost[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    const int c = c0 + s.lane + 32 * k;
    int d, i, pt;
    uint32_t dist;
    valid[k] = false; x[k] = 0; y[k] = 0;
    if (sweep_slot(c, dmax, d, i) && d >= dfirst && d <= dlast) valid[k] = diamond_cand(s, sx, sy, d, i, x[k], y[k], pt, dist);
  }
  eval_multi<K>(s, valid, x, y, cost);
  int off = c0;
  for (int d = dfirst; d <= dlast && d <= dmax; d <<= 1) {
    const int sz = round_size(d);
    uint32_t c = kNoCost;
    unsigned pos = 0xffu;
#pragma unroll
    for (int k = 0; k < K; k++) {
      const int ci = c0 + s.lane + 32 * k;
      if (ci >= off && ci < off + sz) { c = cost[k]; pos = (unsigned)(ci - off); }
    }
    s.best_round += 1;
    s.n_sads += __popc(__ballot_sync(0xffffffffu, c != kNoCost));
    const uint32_t mn = __reduce_min_sync(0xffffffffu, c);
    if (mn < s.best_sad) {
      const int i = (int)__reduce_min_sync(0xffffffffu, c == mn ? pos : 0xffu);
      int bx, by, pt;
      uint32_t dist;
      diamond_cand(s, sx, sy, d, i, bx, by, pt, dist);
      s.best_sad = mn; s.best_x = bx; s.best_y = by; s.best_dist = dist; s.point_nr = pt; s.best_round = 0;
    }
    if (s.best_round >= 3) return true;
    off += sz;
  }
  return false;
}

template <int K>
__device__ __forceinline__ bool diamond_sweep(SearchCtx& s, int sx, int sy, int d0, int dmax, bool first_search)
{
  if (first_search && s.grp) {
    if (sweep_stage<1>(s, sx, sy, 0, 1, 8, dmax)) return true;
    if (dmax < 16) return false;
    return sweep_stage<2>(s, sx, sy, 28, 16, 64, dmax);
  }
  bool valid[K];
  int x[K], y[K];
  uint32_t cost[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    const int c = s.lane + 32 * k;
    int d, i, pt;
    uint32_t dist;
    valid[k] = false; x[k] = 0; y[k] = 0;
    if (sweep_slot(c, dmax, d, i)) valid[k] = diamond_cand(s, sx, sy, d, i, x[k], y[k], pt, dist);
  }
  eval_multi<K>(s, valid, x, y, cost);
  if (!first_search) {
    // star refinement: every round of the sweep runs (no early stop), so the sequential updates collapse
    // into ONE ordered arg-min over the whole sweep: first candidate in visiting order with the minimum
    uint32_t lb = kNoCost;
    int lc = 0x7fffffff, nvalid = 0;
#pragma unroll
    for (int k = 0; k < K; k++) {
      nvalid += cost[k] != kNoCost;
      if (cost[k] < lb) { lb = cost[k]; lc = s.lane + 32 * k; }     // slots of a lane are in visiting order
    }
    s.n_sads += __reduce_add_sync(0xffffffffu, (unsigned)nvalid);
    const uint32_t mn = __reduce_min_sync(0xffffffffu, lb);
    if (mn < s.best_sad) {
      const int c = (int)__reduce_min_sync(0xffffffffu, (unsigned)(lb == mn ? lc : 0x7fffffff));
      int d, i, bx, by, pt;
      uint32_t dist;
      sweep_slot(c, dmax, d, i);
      diamond_cand(s, sx, sy, d, i, bx, by, pt, dist);
      s.best_sad = mn; s.best_x = bx; s.best_y = by; s.best_dist = dist; s.point_nr = pt; s.best_round = 0;
    }
    return false;
  }
  int off = 0;
  for (int d = d0; d <= dmax; d <<= 1) {
    const int sz = round_size(d);
    // this lane's candidate of the round, if any (a round has <= 16 candidates: at most one per lane)
    uint32_t c = kNoCost;
#pragma unroll
    for (int k = 0; k < K; k++) {
      const int ci = s.lane + 32 * k;
      if (ci >= off && ci < off + sz) c = cost[k];
    }
    s.best_round += 1;
    s.n_sads += __popc(__ballot_sync(0xffffffffu, c != kNoCost));
    const uint32_t mn = __reduce_min_sync(0xffffffffu, c);
    if (mn < s.best_sad) {
      // first candidate in visiting order that attains the minimum
      const unsigned pos = (unsigned)((s.lane - off) & 31);
      const int i = (int)__reduce_min_sync(0xffffffffu, c == mn ? pos : 0xffu);
      int bx, by, pt;
      uint32_t dist;
      diamond_cand(s, sx, sy, d, i, bx, by, pt, dist);
      s.best_sad = mn; s.best_x = bx; s.best_y = by; s.best_dist = dist; s.point_nr = pt; s.best_round = 0;
    }
    if (s.best_round >= 3) return true;
    off += sz;
  }
  return false;
}

// xTZ2PointSearch (TEncSearch.cpp:351-476): the two untested neighbours of the best point; border
// tests as in the reference (only the edges named there)
__device__ __forceinline__ void two_point(SearchCtx& s)
{
  const int bx = s.best_x, by = s.best_y;
  const bool up = (by - 1) >= s.ty, dn = (by + 1) <= s.by, lf = (bx - 1) >= s.lx, rt = (bx + 1) <= s.rx;
  int x[1] = {0}, y[1] = {0};
  bool valid[1] = {false};
  if (s.lane < 2) {
    const bool f = s.lane == 0;
    switch (s.point_nr) {
      case 1: x[0] = f ? bx - 1 : bx; y[0] = f ? by : by - 1; valid[0] = f ? lf : up; break;
      case 2: x[0] = f ? bx - 1 : bx + 1; y[0] = by - 1; valid[0] = up && (f ? lf : rt); break;
      case 3: x[0] = f ? bx : bx + 1; y[0] = f ? by - 1 : by; valid[0] = f ? up : rt; break;
      case 4: x[0] = bx - 1; y[0] = f ? by + 1 : by - 1; valid[0] = lf && (f ? dn : up); break;
      case 5: x[0] = bx + 1; y[0] = f ? by - 1 : by + 1; valid[0] = rt && (f ? up : dn); break;
      case 6: x[0] = f ? bx - 1 : bx; y[0] = f ? by : by + 1; valid[0] = f ? lf : dn; break;
      case 7: x[0] = f ? bx - 1 : bx + 1; y[0] = by + 1; valid[0] = dn && (f ? lf : rt); break;
      case 8: x[0] = f ? bx + 1 : bx; y[0] = f ? by : by + 1; valid[0] = f ? rt : dn; break;
      default: break;   // the reference asserts; unreachable (distance 1 always carries a point number)
    }
  }
  uint32_t cost[1];
  eval_multi<1>(s, valid, x, y, cost);
  s.n_sads += __popc(__ballot_sync(0xffffffffu, cost[0] != kNoCost));
  const uint32_t mn = __reduce_min_sync(0xffffffffu, cost[0]);
  if (mn < s.best_sad) {
    const int src = __ffs(__ballot_sync(0xffffffffu, cost[0] == mn)) - 1;
    s.best_sad = mn;
    s.best_x = __shfl_sync(0xffffffffu, x[0], src);
    s.best_y = __shfl_sync(0xffffffffu, y[0], src);
    s.best_dist = 2; s.point_nr = 0; s.best_round = 0;
  }
}

// raster over the window with the given step (xPatternSearch with step 1, TEncSearch.cpp:4227-4283;
// the TZ raster stage with step iRaster, :4389-4400): y outer, x inner, strict '<'.  K candidates per
// lane and iteration.
template <int K>
__device__ __forceinline__ void raster_scan(SearchCtx& s, int step, uint32_t dist_tag)
{
  const int nx = (s.rx - s.lx) / step + 1, ny = (s.by - s.ty) / step + 1, N = nx * ny;
  uint32_t lbest = kNoCost;
  int lidx = 0x7fffffff;
  for (int base = 0; base < N; base += 32 * K) {
    bool valid[K];
    int x[K], y[K];
    uint32_t cost[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
      const int i = base + 32 * k + s.lane;
      valid[k] = i < N;
      const int iy = valid[k] ? i / nx : 0, ix = valid[k] ? i - iy * nx : 0;
      x[k] = s.lx + ix * step; y[k] = s.ty + iy * step;
    }
    eval_multi<K>(s, valid, x, y, cost);
#pragma unroll
    for (int k = 0; k < K; k++)
      if (cost[k] < lbest) { lbest = cost[k]; lidx = base + 32 * k + s.lane; }     // increasing index within a lane
  }
  s.n_sads += (uint32_t)N;
  const uint32_t mn = __reduce_min_sync(0xffffffffu, lbest);
  if (mn < s.best_sad) {
    const int idx = (int)__reduce_min_sync(0xffffffffu, (unsigned)(lbest == mn ? lidx : 0x7fffffff));
    s.best_sad = mn;
    s.best_x = s.lx + (idx % nx) * step;
    s.best_y = s.ty + (idx / nx) * step;
    s.best_dist = dist_tag; s.best_round = 0; s.point_nr = 0;
  }
}

// ================================================================================ (2b) shared raster stage
// In the frame pre-pass every PU of a CTU searches the same window around the same predictor, and
// the TZ raster stage (TEncSearch.cpp:4389-4400: every 5th candidate of the whole window, taken by
// about half of the PUs and then ~80% of all their SAD evaluations) visits the SAME candidates for all
// 593 PUs.  k_me_raster therefore walks the raster once per (CTU, reference): each candidate is one
// coalesced 1 KB read; its 256 even-row and 256 odd-row block SADs are turned into two 17x17 integral
// images in shared memory, from which every PU's SAD is four look-ups; each PU thread keeps the
// running (cost, first index) minimum in visiting order with strict '<' -- the state xTZSearchHelp
// would reach.  k_me_search consumes the result when its own window equals the CTU's.
struct RasterBest { uint32_t cost; int32_t idx; };
// state of a PU's search after the start / zero tests and the first diamond sweep (IntTZSearchStruct fields);
// n_sads == 0: not served by the shared stage
struct SweepState { uint32_t best_sad; int32_t best_x, best_y; uint32_t best_dist; int32_t point_nr; uint32_t n_sads; };
constexpr int kRastChunk = 4;                   // candidates per iteration
constexpr int kRastThreads = 640;               // >= TVC_ME_CENSUS PU threads
constexpr int kSweepMax = 2 + 4 + 3 * 8 + 3 * 16;   // start, zero, rounds d = 1 .. 64

// Walks N candidates in visiting order, kRastChunk at a time: threads 0..255 fetch the chunk's 1 KB candidate blocks
// (pos(i, dx, dy) gives the table coordinates, false = skip), the block SADs become two 17x17 integral images per
// candidate, then every thread runs consume(base) (PU threads read four corners per candidate).  The next chunk's
// loads fly during the scans.
template <class PosF, class MvcF, class ConsumeF>
__device__ __forceinline__ void walk_candidates(int N, const uint4* __restrict__ tg, uint32_t (*IE)[17 * 17], uint32_t (*IA)[17 * 17],
                                                uint32_t* s_mvc, int tid, PosF pos, MvcF mvc, ConsumeF consume)
{
  const int lc = tid >> 6, lg = tid & 63;
  auto fetch = [&](int base) -> uint4 {
    const int i = base + lc;
    int dx, dy;
    if (tid >= 64 * kRastChunk || i >= N || !pos(i, dx, dy)) return make_uint4(0, 0, 0, 0);
    return __ldg(tg + me_granule(dy, dx, 0, 0) + lg);
  };
  uint4 g = fetch(0);
  for (int base = 0; base < N; base += kRastChunk) {
    // (1) scatter the granule: block row lg >> 2, blocks 4*(lg & 3) .. +3
    if (tid < 64 * kRastChunk) {
      const int by = lg >> 2, bx = (lg & 3) * 4;
      uint32_t* e = &IE[lc][(by + 1) * 17 + bx + 1];
      uint32_t* a = &IA[lc][(by + 1) * 17 + bx + 1];
      const uint32_t e0 = g.x & 0xffffu, e1 = g.x >> 16, e2 = g.y & 0xffffu, e3 = g.y >> 16;
      e[0] = e0; e[1] = e1; e[2] = e2; e[3] = e3;
      a[0] = e0 + (g.z & 0xffffu); a[1] = e1 + (g.z >> 16); a[2] = e2 + (g.w & 0xffffu); a[3] = e3 + (g.w >> 16);
    } else if (tid < 64 * kRastChunk + kRastChunk) {
      const int c = tid - 64 * kRastChunk, i = base + c;
      if (i < N) s_mvc[c] = mvc(i);
    }
    g = fetch(base + kRastChunk);
    __syncthreads();
    // (2) row prefix sums: 2 images x chunk x 16 rows
    if (tid < 2 * kRastChunk * 16) {
      const int img = tid / (kRastChunk * 16), c = (tid / 16) % kRastChunk, r = tid % 16;
      uint32_t* p = (img ? IA[c] : IE[c]) + (r + 1) * 17 + 1;
      uint32_t s = 0;
#pragma unroll
      for (int x = 0; x < 16; x++) { s += p[x]; p[x] = s; }
    }
    __syncthreads();
    // (3) column prefix sums
    if (tid < 2 * kRastChunk * 16) {
      const int img = tid / (kRastChunk * 16), c = (tid / 16) % kRastChunk, x = tid % 16;
      uint32_t* p = (img ? IA[c] : IE[c]) + 17 + x + 1;
      uint32_t s = 0;
#pragma unroll
      for (int r = 0; r < 16; r++) { s += p[r * 17]; p[r * 17] = s; }
    }
    __syncthreads();
    // (4) every PU: cost of the chunk's candidates in visiting order
    consume(base);
    __syncthreads();
  }
}

// Shared stages of the frame pre-pass, one CTA per (CTU, reference).  All 593 PUs of the CTU search around the same
// predictor, so (a) the start / zero tests and the FIRST diamond sweep (xTZSearch :4320-4361: <= 78 candidates around
// the clipped predictor) and (b) the raster stage (every 5th candidate of the window) visit the same candidates for
// every PU: each candidate is read once (1 KB, coalesced) and every PU thread replays the reference's sequential
// strict-'<' update on its own SAD.  PUs whose window / start differ from the CTU's (picture border: clipMv depends on
// the CU origin), whose zero vector beats the predictor (the sweep then centres elsewhere) or whose candidates leave
// the table window fall back to k_me_search's own evaluation.
template <bool SWEEP>            // true: stage (a) only, false: stage (b) only -- two launches keep the raster at 2 CTAs per SM
__global__ void __launch_bounds__(kRastThreads)
k_me_raster(const tvc_me_job* __restrict__ jobs, const uint16_t* __restrict__ tables, const tvc_me_center* __restrict__ centers,
            int num_ctus, int raster, int bi, RasterBest* __restrict__ out, SweepState* __restrict__ sweep_out,
            unsigned long long* __restrict__ stats, int ctu0, int ref0)
{
  __shared__ uint32_t IE[kRastChunk][17 * 17];  // integral of even-row block SADs
  __shared__ uint32_t IA[kRastChunk][17 * 17];  // integral of even+odd
  __shared__ uint32_t s_mvc[kRastChunk];
  __shared__ int16_t s_cx[kSweepMax], s_cy[kSweepMax];
  __shared__ uint8_t s_cvalid[kSweepMax], s_cpt[kSweepMax], s_cfirst[kSweepMax], s_clast[kSweepMax];
  __shared__ uint32_t s_cdist[kSweepMax], s_cmv[kSweepMax];
  __shared__ int s_sweep_ok;
  const int ctu = blockIdx.x + ctu0, ref = blockIdx.y + ref0, tid = threadIdx.x;
  const size_t rc = (size_t)ref * num_ctus + ctu;
  const tvc_me_job* jb0 = jobs + rc * TVC_ME_CENSUS;
  const tvc_me_job j0 = jb0[0];                  // the 64x64 PU: its window is the CTU's
  RasterBest* o = out + rc * TVC_ME_CENSUS;
  SweepState* so = sweep_out ? sweep_out + rc * TVC_ME_CENSUS : nullptr;
  const tvc_me_center cen = centers[rc];
  const int nx = (j0.rx - j0.lx) / raster + 1, ny = (j0.by - j0.ty) / raster + 1, N = nx * ny;
  const int x_last = j0.lx + (nx - 1) * raster, y_last = j0.ty + (ny - 1) * raster;
  const bool usable = j0.w > 0 && j0.mode == TVC_ME_TZ;
  const bool covered = usable && j0.lx - cen.cx >= -kMeR && x_last - cen.cx <= kMeR &&
                       j0.ty - cen.cy >= -kMeR && y_last - cen.cy <= kMeR;
  const uint4* __restrict__ tg = reinterpret_cast<const uint4*>(tables + rc * kMeCtuElems);
  // this thread's PU
  const bool is_pu = tid < TVC_ME_CENSUS;
  const tvc_census_pu cp = c_census[is_pu ? tid : 0];
  const tvc_me_job jb = jb0[is_pu ? tid : 0];
  const bool pu_ok = is_pu && jb.w > 0;
  const bool same = pu_ok && jb.lx == j0.lx && jb.ty == j0.ty && jb.rx == j0.rx && jb.by == j0.by && jb.startx == j0.startx &&
                    jb.starty == j0.starty && jb.lambda_cost == j0.lambda_cost && jb.search_range == j0.search_range;
  const int bx0 = cp.x >> 2, by0 = cp.y >> 2, bx1 = (cp.x + cp.w) >> 2, by1 = (cp.y + cp.h) >> 2;
  const int sub = (j0.fen && cp.h > 8) ? 1 : 0;
  const int c00 = by0 * 17 + bx0, c01 = by0 * 17 + bx1, c10 = by1 * 17 + bx0, c11 = by1 * 17 + bx1;
  if (tid < 17 * kRastChunk) {                   // zero row / column of the integrals
    const int c = tid / 17, k = tid % 17;
    IE[c][k] = 0; IA[c][k] = 0; IE[c][k * 17] = 0; IA[c][k * 17] = 0;
  }

  // ---- (a) start, zero, first diamond sweep
  int nsweep = 0;
  for (int d = 1; d <= j0.search_range; d <<= 1) nsweep += round_size(d);
  const int n1 = 2 + nsweep;
  if (tid == 0) s_sweep_ok = (SWEEP && so && usable) ? 1 : 0;
  __syncthreads();
  if (SWEEP && so && usable && tid < n1) {
    int x = 0, y = 0, pt = 0;
    uint32_t dist = 0;
    bool v = true, first = false, last = false;
    if (tid == 0) { x = j0.startx; y = j0.starty; }
    else if (tid >= 2) {
      SearchCtx sc;
      sc.lx = j0.lx; sc.ty = j0.ty; sc.rx = j0.rx; sc.by = j0.by;
      int c = tid - 2, off = 0, d = 1;
      while (c >= off + round_size(d)) { off += round_size(d); d <<= 1; }
      v = diamond_cand(sc, j0.startx, j0.starty, d, c - off, x, y, pt, dist);
      first = c == off; last = c == off + round_size(d) - 1;
    }
    s_cx[tid] = (int16_t)x; s_cy[tid] = (int16_t)y; s_cvalid[tid] = v; s_cpt[tid] = (uint8_t)pt; s_cdist[tid] = dist;
    s_cfirst[tid] = first; s_clast[tid] = last;
    s_cmv[tid] = mv_cost(j0.lambda_cost, x, y, 2, j0.predx, j0.predy);
    const int dx = x - cen.cx, dy = y - cen.cy;
    if (v && (dx < -kMeR || dx > kMeR || dy < -kMeR || dy > kMeR)) s_sweep_ok = 0;     // benign race: every writer stores 0
  }
  __syncthreads();
  if (SWEEP && s_sweep_ok) {
    uint32_t best = kNoCost, best_dist = 0, n_sads = 0;
    int best_x = 0, best_y = 0, point_nr = 0, best_round = 0;
    bool run = same;                             // this PU's sequential first search is still running
    bool fallback = false;
    walk_candidates(n1, tg, IE, IA, s_mvc, tid,
      [&](int i, int& dx, int& dy) { dx = s_cx[i] - cen.cx; dy = s_cy[i] - cen.cy; return s_cvalid[i] != 0; },
      [&](int i) { return s_cmv[i]; },
      [&](int base) {
        if (!run) return;
#pragma unroll
        for (int c = 0; c < kRastChunk; c++) {
          const int i = base + c;
          if (i >= n1) break;
          if (s_cfirst[i]) best_round += 1;      // xTZ8PointDiamondSearch entry
          if (s_cvalid[i]) {
            const uint32_t* I = sub ? IE[c] : IA[c];
            const uint32_t sad = I[c11] - I[c01] - I[c10] + I[c00];
            const uint32_t cost = ((sad << sub) >> bi) + s_mvc[c];
            n_sads++;
            if (i == 0) { best = cost; best_x = s_cx[0]; best_y = s_cy[0]; }
            else if (i == 1) {
              if (cost < best) { fallback = true; run = false; break; }   // the zero vector wins: the sweep centres on (0,0) -> own search
            } else if (cost < best) {
              best = cost; best_x = s_cx[i]; best_y = s_cy[i]; best_dist = s_cdist[i]; point_nr = s_cpt[i]; best_round = 0;
            }
          }
          if (s_clast[i] && best_round >= 3) { run = false; break; }       // bFirstSearchStop, uiFirstSearchRounds = 3
        }
      });
    if (is_pu) so[tid] = (same && !fallback) ? SweepState{best, best_x, best_y, best_dist, point_nr, n_sads} : SweepState{0, 0, 0, 0, 0, 0};
  } else if (SWEEP && is_pu && so) so[tid] = SweepState{0, 0, 0, 0, 0, 0};
  if (SWEEP) {
    if (tid == 0 && stats && s_sweep_ok) atomicAdd(&stats[2], (unsigned long long)n1);
    return;
  }

  // ---- (b) raster
  if (!covered) {                                // partial CTU / window leaves the table: every PU rasters on its own
    if (is_pu) o[tid] = RasterBest{kNoCost, -1};
    return;
  }
  uint32_t best = kNoCost;
  int best_i = -1;
  walk_candidates(N, tg, IE, IA, s_mvc, tid,
    [&](int i, int& dx, int& dy) { const int iy = i / nx, ix = i - iy * nx; dx = j0.lx + ix * raster - cen.cx; dy = j0.ty + iy * raster - cen.cy; return true; },
    [&](int i) { const int iy = i / nx, ix = i - iy * nx; return mv_cost(j0.lambda_cost, j0.lx + ix * raster, j0.ty + iy * raster, 2, j0.predx, j0.predy); },
    [&](int base) {
      if (!pu_ok) return;
#pragma unroll
      for (int c = 0; c < kRastChunk; c++) {
        if (base + c < N) {
          const uint32_t* I = sub ? IE[c] : IA[c];
          const uint32_t sad = I[c11] - I[c01] - I[c10] + I[c00];
          const uint32_t cost = ((sad << sub) >> bi) + s_mvc[c];
          if (cost < best) { best = cost; best_i = base + c; }
        }
      }
    });
  if (is_pu) o[tid] = RasterBest{pu_ok ? best : kNoCost, pu_ok ? best_i : -1};
  if (tid == 0 && stats) atomicAdd(&stats[2], (unsigned long long)N);
}

// one PU job by one warp (warp-collective: all 32 lanes call it with the same j)
__device__ __forceinline__ void me_search_job(int j, const PlaneTable& pt, int cur_slot, const tvc_me_job* __restrict__ jobs,
                                              tvc_me_result* __restrict__ out, const uint16_t* __restrict__ tables,
                                              const tvc_me_center* __restrict__ centers, int num_ctus, int ctus_x, int bi,
                                              const RasterBest* __restrict__ rast, const SweepState* __restrict__ sweep,
                                              unsigned long long* __restrict__ stats, const GrpWin* gw = nullptr,
                                              const SweepState* sw_job = nullptr, unsigned long long* grp_acc = nullptr)
{
  const tvc_me_job jb = jobs[j];
  if (jb.w <= 0) {                       // census PU outside the picture (frame pre-pass)
    if ((threadIdx.x & 31) == 0) out[j] = tvc_me_result{0, 0, 0u, 0u};
    return;
  }
  SearchCtx s;
  s.lane = threadIdx.x & 31;
  s.so = pt.stride[0]; s.rs = pt.stride[0];
  s.org = pt.org[cur_slot][0] + (ptrdiff_t)jb.y * s.so + jb.x;
  s.ref = pt.org[jb.ref_slot][0] + (ptrdiff_t)jb.y * s.rs + jb.x;
  s.w = jb.w; s.h = jb.h; s.bi = bi;
  s.sub = (jb.fen && jb.h > 8) ? 1 : 0;                      // TEncSearch.cpp:324-330 / 4245-4251
  s.lc = jb.lambda_cost; s.px = jb.predx; s.py = jb.predy;
  s.lx = jb.lx; s.ty = jb.ty; s.rx = jb.rx; s.by = jb.by;
  s.tbl = nullptr; s.tcx = 0; s.tcy = 0;
  s.win8 = nullptr; s.cur8 = nullptr; s.gref8 = nullptr; s.gpitch8 = 0; s.wcx = 0; s.wcy = 0;
  s.grp = gw != nullptr;
  if (gw) {
    const int px = jb.x - gw->x0, py = jb.y - gw->y0;        // the PU inside its CTU
    s.win8 = gw->win + py * kGrpWinW + px + gw->e16;
    s.cur8 = gw->cur + py * kGrpCurP + px;
    s.gref8 = gw->plane8 + (ptrdiff_t)jb.y * gw->pitch8 + jb.x;
    s.gpitch8 = gw->pitch8; s.wcx = gw->cenx; s.wcy = gw->ceny;
  }
  int ctu = (jb.y >> 6) * ctus_x + (jb.x >> 6);
  if (tables) {
    s.tbl = tables + ((size_t)jb.ref_index * num_ctus + ctu) * kMeCtuElems;
    tvc_me_center cen = centers[(size_t)jb.ref_index * num_ctus + ctu];
    s.tcx = cen.cx; s.tcy = cen.cy;
  }
  {
    const int bx0 = (jb.x & 63) >> 2, nbx = jb.w >> 2;
    s.by0 = (jb.y & 63) >> 2; s.nby = jb.h >> 2;
    s.q0 = bx0 >> 2; s.nq = ((bx0 + nbx - 1) >> 2) - s.q0 + 1;
#pragma unroll
    for (int qi = 0; qi < 4; qi++) {
      uint32_t a = 0, b = 0;
      const int b0 = (s.q0 + qi) * 4;
      if (b0 + 0 >= bx0 && b0 + 0 < bx0 + nbx) a |= 0x0000ffffu;
      if (b0 + 1 >= bx0 && b0 + 1 < bx0 + nbx) a |= 0xffff0000u;
      if (b0 + 2 >= bx0 && b0 + 2 < bx0 + nbx) b |= 0x0000ffffu;
      if (b0 + 3 >= bx0 && b0 + 3 < bx0 + nbx) b |= 0xffff0000u;
      s.m01[qi] = a; s.m23[qi] = b;
    }
  }
  s.best_sad = kNoCost; s.best_x = 0; s.best_y = 0; s.best_dist = 0; s.best_round = 0; s.point_nr = 0;
  s.n_sads = 0;
  uint32_t served = 0;          // candidates of this job that the shared raster stage evaluated

  if (jb.mode == TVC_ME_FULL) {
    raster_scan<TVC_RASTER_K>(s, 1, 0);
  } else {
    // xTZSearch with TZ_SEARCH_CONFIGURATION (TEncSearch.cpp:293-309, 4302-4474)
    const int raster = 5, srange = jb.search_range;
    SweepState sw = {0, 0, 0, 0, 0, 0};
    if (sw_job) sw = *sw_job;
    else if (sweep) sw = sweep[j];
    if (sw.n_sads) {
      // frame pre-pass: start / zero tests and the first sweep were replayed by the shared stage (k_me_raster)
      s.best_sad = sw.best_sad; s.best_x = sw.best_x; s.best_y = sw.best_y; s.best_dist = sw.best_dist; s.point_nr = sw.point_nr;
      s.best_round = 0; s.n_sads = sw.n_sads; served = sw.n_sads;
    } else {
      // start point (the clipped predictor) then the zero vector (:4320, :4336-4339), sequentially
      const int x[1] = {s.lane == 0 ? jb.startx : 0}, y[1] = {s.lane == 0 ? jb.starty : 0};
      const bool valid[1] = {s.lane < 2};
      uint32_t cost[1];
      eval_multi<1>(s, valid, x, y, cost);
      const uint32_t c0 = __shfl_sync(0xffffffffu, cost[0], 0), c1 = __shfl_sync(0xffffffffu, cost[0], 1);
      s.n_sads += 2;
      s.best_sad = c0; s.best_x = jb.startx; s.best_y = jb.starty;
      if (c1 < c0) { s.best_sad = c1; s.best_x = 0; s.best_y = 0; }
      s.best_dist = 0; s.best_round = 0; s.point_nr = 0;
    }
    // a sweep holds 4 + 3*8 + 3*16 = 76 candidates for search range 64: 3 per lane (the ABI bounds the range)
    // first search :4346-4361.  All 76 candidates of the seven rounds are evaluated at once although the
    // reference usually stops after three rounds: measured on B200, issuing the rounds in stages (1..8,
    // 16..32, 64) costs more in serial latency (9.1 ms) than the speculative loads cost in bandwidth (8.2 ms).
    if (!sw.n_sads) diamond_sweep<3>(s, s.best_x, s.best_y, 1, srange, true);
    if (s.best_dist == 1) { s.best_dist = 0; two_point(s); }   // :4382-4386
    if ((int)s.best_dist > raster) {                           // :4389-4400
      s.best_dist = raster;
      // frame pre-pass: the raster of this PU was walked by k_me_raster if its window is the CTU's
      bool shared = false;
      if (rast) {
        const RasterBest rb = rast[j];
        const tvc_me_job& j0 = jobs[j - (j % TVC_ME_CENSUS)];
        shared = rb.idx >= 0 && j0.lx == jb.lx && j0.ty == jb.ty && j0.rx == jb.rx && j0.by == jb.by;
        if (shared) {
          const int nx = (s.rx - s.lx) / raster + 1, ny = (s.by - s.ty) / raster + 1;
          s.n_sads += (uint32_t)(nx * ny);
          served += (uint32_t)(nx * ny);
          if (rb.cost < s.best_sad) {
            s.best_sad = rb.cost;
            s.best_x = s.lx + (rb.idx % nx) * raster;
            s.best_y = s.ty + (rb.idx / nx) * raster;
            s.best_dist = raster; s.best_round = 0; s.point_nr = 0;
          }
        }
      }
      if (!shared) raster_scan<TVC_RASTER_K>(s, raster, raster);
    }
    while (s.best_dist > 0) {                                  // star refinement :4435-4468
      const int sx = s.best_x, sy = s.best_y;
      s.best_dist = 0; s.point_nr = 0;
      diamond_sweep<3>(s, sx, sy, 1, srange, false);
      if (s.best_dist == 1) {
        s.best_dist = 0;
        if (s.point_nr != 0) two_point(s);
      }
    }
  }
  if (s.lane == 0) {
    tvc_me_result r;
    r.mvx = s.best_x; r.mvy = s.best_y;
    r.sad = s.best_sad - mv_cost(s.lc, s.best_x, s.best_y, 2, s.px, s.py);
    r.n_sads = s.n_sads;
    out[j] = r;
    if (grp_acc) {
      // group kernel: the warp sums its PUs and adds once at the end (two atomics per PU on the same two words, 2.4 M per 1080p
      // picture, were 34 % of the kernel's stall samples: every warp's next PU waited for the L2 to take the previous one's pair)
      grp_acc[0] += (unsigned long long)s.n_sads;
      grp_acc[1] += (unsigned long long)s.n_sads * (unsigned)(s.w * (s.h >> s.sub));
    } else if (stats && gw) {
      atomicAdd(&stats[0], (unsigned long long)s.n_sads);
      atomicAdd(&stats[1], (unsigned long long)s.n_sads * (unsigned)(s.w * (s.h >> s.sub)));
    } else if (stats) {
      atomicAdd(&stats[0], (unsigned long long)(s.n_sads - served) * (unsigned)(s.nby * s.nq));
      atomicAdd(&stats[1], (unsigned long long)served);
    }
  }
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB)
k_me_search(PlaneTable pt, int cur_slot, int n, const tvc_me_job* __restrict__ jobs, tvc_me_result* __restrict__ out,
            const uint16_t* __restrict__ tables, const tvc_me_center* __restrict__ centers, int num_ctus, int ctus_x,
            int bi, const RasterBest* __restrict__ rast, const SweepState* __restrict__ sweep, unsigned long long* __restrict__ stats)
{
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= n) return;
  me_search_job(j, pt, cur_slot, jobs, out, tables, centers, num_ctus, ctus_x, bi, rast, sweep, stats);
}

// the jobs named by a device-resident list (the PUs the group kernel of tvc_me_group.cu handed back: candidates beyond its
// bitmap, round cap): a fixed grid walks the list, SADs straight from the pictures
__global__ void __launch_bounds__(128, 4)
k_me_search_list(PlaneTable pt, int cur_slot, const int* __restrict__ list, const int* __restrict__ count, const tvc_me_job* __restrict__ jobs,
                 tvc_me_result* __restrict__ out, int ctus_x, int bi)
{
  const int n = *count;
  for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += gridDim.x * (blockDim.x >> 5))
    me_search_job(list[i], pt, cur_slot, jobs, out, nullptr, nullptr, 0, ctus_x, bi, nullptr, nullptr, nullptr);
}

// ================================================================================ (2c) group search
// Integer search of a whole (CTU, reference) census group by ONE CTA without SAD tables.  Round 1 wrote the full +-64 tables of
// every (CTU, reference) to HBM (17 MB each, 34.8 GB per 1080p picture with four references) although only 18 % of the bytes
// were ever read; here the 208 x 192 u8 search window and the 64 x 64 u8 CTU are staged into shared memory once (two TMA box
// copies) and every SAD a search asks for is computed on the spot: warp per PU (the 593 census PUs are taken from a shared
// counter, largest first), lane per candidate, VABSDIFF4 over words of the staged window.  The TZ control flow is the per-PU
// kernel's (diamond_sweep / two_point / raster_scan with the ordered arg-min replay), so the results -- MV, ruiSAD, number of
// SADs -- are those of xTZSearch; candidates beyond the staged window (the zero vector of a far predictor, PUs whose clipMv
// differs from the CTU's at the picture border) read the reference's u8 plane in global memory through the same routine.
// ---- first search at CU level.  The PUs of a CU (13, or 5 at 8x8) share CU origin, hence clipMv, search window, start point and
// every candidate of the first search (start, zero vector, the diamond rounds around the winner), and each of them is a union of
// the CU's 4x4 grid of sub-blocks (2x2 at 8x8): 2NxN = sub-block rows {0,1} / {2,3}, 2NxnU = row 0 / rows 1-3, nLx2N = column 0 /
// columns 1-3, ...  One warp per CU, lane per candidate: the lane walks the CU ONCE, accumulating even-row and odd-row sums per
// sub-block (the FEN row sub-sampling applies per PU: h > 8), and derives the cost of all 13 PUs from row / column sums of that
// grid -- one pass instead of thirteen (the PUs cover the CU seven times over).  Then lane p replays the reference's sequential
// first search for PU p on the stored costs (strict '<' in visiting order, stop after three rounds without improvement) and leaves
// a SweepState from which the per-PU code continues (xTZ2PointSearch, raster, star refinement): the hand-off the round-1 shared
// stage used.  CUs that stick out of the picture, or whose candidates leave the staged window, are left to the per-PU code.
constexpr int kCuTasks = 1 + 4 + 16 + 64;
struct CuTask { int base, np, S, px, py; };
__device__ __forceinline__ CuTask cu_task(int t)
{
  CuTask c;
  if (t == 0) { c.S = 64; c.base = 0; c.np = 13; c.px = 0; c.py = 0; }
  else if (t < 5) { const int i = t - 1; c.S = 32; c.base = 13 + 13 * i; c.np = 13; c.px = (i & 1) * 32; c.py = (i >> 1) * 32; }
  else if (t < 21) { const int i = t - 5; c.S = 16; c.base = 65 + 13 * i; c.np = 13; c.px = (i & 3) * 16; c.py = (i >> 2) * 16; }
  else { const int i = t - 21; c.S = 8; c.base = 273 + 5 * i; c.np = 5; c.px = (i & 7) * 8; c.py = (i >> 3) * 8; }
  return c;
}

// costs of the CU's PUs at one candidate, written to cst[0 .. np): ref = window byte of the CU's top-left at the candidate
template <int S>
__device__ __forceinline__ void cu_costs(const uint8_t* __restrict__ ref, const uint8_t* __restrict__ cur, bool need_odd,
                                         const uint8_t* __restrict__ subp, uint32_t mvc, uint32_t* __restrict__ cst)
{
  constexpr int NB = S == 8 ? 2 : 4, Q = S / NB, QW = Q / 4, RW = S / 4;      // sub-blocks per side, their size in pels / words
  const int a = (int)((uintptr_t)ref & 3), sh = a * 8;
  const uint32_t* rb = reinterpret_cast<const uint32_t*>(ref - a);
  const uint32_t* cb = reinterpret_cast<const uint32_t*>(cur);
  uint32_t E[NB][NB], O[NB][NB];
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) { E[i][j] = 0; O[i][j] = 0; }
#pragma unroll
  for (int sbr = 0; sbr < NB; sbr++) {
    for (int rr = 0; rr < Q; rr += 2) {
      const uint32_t* r0 = rb + (sbr * Q + rr) * (kGrpWinW / 4);
      const uint32_t* c0 = cb + (sbr * Q + rr) * (kGrpCurP / 4);
      uint32_t prev = r0[0];
#pragma unroll
      for (int j = 0; j < RW; j++) {
        const uint32_t nxt = r0[j + 1];
        E[sbr][j / QW] = vsad4_acc(c0[j], __funnelshift_r(prev, nxt, sh), E[sbr][j / QW]);
        prev = nxt;
      }
      if (need_odd) {
        const uint32_t* r1 = r0 + kGrpWinW / 4;
        const uint32_t* c1 = c0 + kGrpCurP / 4;
        prev = r1[0];
#pragma unroll
        for (int j = 0; j < RW; j++) {
          const uint32_t nxt = r1[j + 1];
          O[sbr][j / QW] = vsad4_acc(c1[j], __funnelshift_r(prev, nxt, sh), O[sbr][j / QW]);
          prev = nxt;
        }
      }
    }
  }
  // row / column sums of the grid, per parity
  uint32_t re[NB], ce[NB], ro[NB], co[NB], te = 0, to = 0;
#pragma unroll
  for (int i = 0; i < NB; i++) { re[i] = 0; ce[i] = 0; ro[i] = 0; co[i] = 0; }
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) { re[i] += E[i][j]; ce[j] += E[i][j]; ro[i] += O[i][j]; co[j] += O[i][j]; }
#pragma unroll
  for (int i = 0; i < NB; i++) { te += re[i]; to += ro[i]; }
  auto put = [&](int p, uint32_t e, uint32_t o) {
    const uint32_t sub = subp[p];
    cst[p] = ((sub ? e : e + o) << sub) + mvc;
  };
  // census part order: 2Nx2N, 2NxN[0,1], Nx2N[0,1], 2NxnU[0,1], 2NxnD[0,1], nLx2N[0,1], nRx2N[0,1]
  put(0, te, to);
  if (NB == 2) {
    put(1, re[0], ro[0]); put(2, re[1], ro[1]); put(3, ce[0], co[0]); put(4, ce[1], co[1]);
  } else {
    put(1, re[0] + re[1], ro[0] + ro[1]); put(2, re[NB - 2] + re[NB - 1], ro[NB - 2] + ro[NB - 1]);
    put(3, ce[0] + ce[1], co[0] + co[1]); put(4, ce[NB - 2] + ce[NB - 1], co[NB - 2] + co[NB - 1]);
    put(5, re[0], ro[0]); put(6, te - re[0], to - ro[0]);
    put(7, te - re[NB - 1], to - ro[NB - 1]); put(8, re[NB - 1], ro[NB - 1]);
    put(9, ce[0], co[0]); put(10, te - ce[0], to - co[0]);
    put(11, te - ce[NB - 1], to - co[NB - 1]); put(12, ce[NB - 1], co[NB - 1]);
  }
}

// The same sums by loops over the sub-block rows and the row pairs instead of straight-line code (cu_costs<64> unrolled is 40 KB
// of SASS that a warp streams through once per candidate, pushing the other warps' loops out of the instruction cache), behind a
// call, operands as offsets into the CTA's dynamic shared memory (see sad_u8_smem).
template <int S>
__device__ __noinline__ void cu_costs_smem(uint32_t ref_off, uint32_t cur_off, bool need_odd, uint32_t subp_off, uint32_t mvc, uint32_t cst_off)
{
  extern __shared__ __align__(128) uint8_t gsm[];
  constexpr int NB = S == 8 ? 2 : 4, Q = S / NB, QW = Q / 4, RW = S / 4;
  const uint8_t* ref = gsm + ref_off;
  const int a = (int)(ref_off & 3u), sh = a * 8;                       // gsm is 128-byte aligned
  const uint32_t* rb = reinterpret_cast<const uint32_t*>(ref - a);
  const uint32_t* cb = reinterpret_cast<const uint32_t*>(gsm + cur_off);
  const uint8_t* subp = gsm + subp_off;
  uint32_t* cst = reinterpret_cast<uint32_t*>(gsm + cst_off);
  uint32_t re[NB], ce[NB], ro[NB], co[NB], te = 0, to = 0;
#pragma unroll
  for (int i = 0; i < NB; i++) { re[i] = 0; ce[i] = 0; ro[i] = 0; co[i] = 0; }
#pragma unroll 1
  for (int sbr = 0; sbr < NB; sbr++) {
    uint32_t e[NB], o[NB];
#pragma unroll
    for (int j = 0; j < NB; j++) { e[j] = 0; o[j] = 0; }
#pragma unroll 1
    for (int rr = 0; rr < Q; rr += 2) {
      const uint32_t* r0 = rb + (sbr * Q + rr) * (kGrpWinW / 4);
      const uint32_t* c0 = cb + (sbr * Q + rr) * (kGrpCurP / 4);
      uint32_t prev = r0[0];
#pragma unroll
      for (int j = 0; j < RW; j++) {
        const uint32_t nxt = r0[j + 1];
        e[j / QW] = vsad4_acc(c0[j], __funnelshift_r(prev, nxt, sh), e[j / QW]);
        prev = nxt;
      }
      if (need_odd) {
        const uint32_t* r1 = r0 + kGrpWinW / 4;
        const uint32_t* c1 = c0 + kGrpCurP / 4;
        prev = r1[0];
#pragma unroll
        for (int j = 0; j < RW; j++) {
          const uint32_t nxt = r1[j + 1];
          o[j / QW] = vsad4_acc(c1[j], __funnelshift_r(prev, nxt, sh), o[j / QW]);
          prev = nxt;
        }
      }
    }
    uint32_t rse = 0, rso = 0;
#pragma unroll
    for (int j = 0; j < NB; j++) { rse += e[j]; rso += o[j]; ce[j] += e[j]; co[j] += o[j]; }
#pragma unroll
    for (int i = 0; i < NB; i++) if (i == sbr) { re[i] = rse; ro[i] = rso; }
    te += rse; to += rso;
  }
  auto put = [&](int p, uint32_t ev, uint32_t ov) {
    const uint32_t sub = subp[p];
    cst[p] = ((sub ? ev : ev + ov) << sub) + mvc;
  };
  put(0, te, to);
  if (NB == 2) {
    put(1, re[0], ro[0]); put(2, re[1], ro[1]); put(3, ce[0], co[0]); put(4, ce[1], co[1]);
  } else {
    put(1, re[0] + re[1], ro[0] + ro[1]); put(2, re[NB - 2] + re[NB - 1], ro[NB - 2] + ro[NB - 1]);
    put(3, ce[0] + ce[1], co[0] + co[1]); put(4, ce[NB - 2] + ce[NB - 1], co[NB - 2] + co[NB - 1]);
    put(5, re[0], ro[0]); put(6, te - re[0], to - ro[0]);
    put(7, te - re[NB - 1], to - ro[NB - 1]); put(8, re[NB - 1], ro[NB - 1]);
    put(9, ce[0], co[0]); put(10, te - ce[0], to - co[0]);
    put(11, te - ce[NB - 1], to - co[NB - 1]); put(12, ce[NB - 1], co[NB - 1]);
  }
}

__device__ __forceinline__ void cu_costs_any(int S, const uint8_t* ref, const uint8_t* cur, bool need_odd, const uint8_t* subp, uint32_t mvc,
                                             uint32_t* cst)
{
#if TVC_GRP_VAR & 2
  extern __shared__ __align__(128) uint8_t gsm[];
  const uint32_t ro_ = (uint32_t)(ref - gsm), co_ = (uint32_t)(cur - gsm), so_ = (uint32_t)(subp - gsm);
  const uint32_t to_ = (uint32_t)(reinterpret_cast<const uint8_t*>(cst) - gsm);
  switch (S) {
    case 64: cu_costs_smem<64>(ro_, co_, need_odd, so_, mvc, to_); break;
    case 32: cu_costs_smem<32>(ro_, co_, need_odd, so_, mvc, to_); break;
    case 16: cu_costs_smem<16>(ro_, co_, need_odd, so_, mvc, to_); break;
    default: cu_costs_smem<8>(ro_, co_, need_odd, so_, mvc, to_); break;
  }
  return;
#endif
  switch (S) {
    case 64: cu_costs<64>(ref, cur, need_odd, subp, mvc, cst); break;
    case 32: cu_costs<32>(ref, cur, need_odd, subp, mvc, cst); break;
    case 16: cu_costs<16>(ref, cur, need_odd, subp, mvc, cst); break;
    default: cu_costs<8>(ref, cur, need_odd, subp, mvc, cst); break;
  }
}

constexpr int kCstPitch = 14;          // costs of up to 13 PUs per candidate (+ pad)
struct CuScratch { uint32_t cst[32][kCstPitch]; uint8_t subp[16]; };

// first search of one CU by one warp; writes the SweepState of its PUs (n_sads == 0: not served)
__device__ __forceinline__ void cu_first_search(const CuTask ct, const tvc_me_job* __restrict__ gjobs, const GrpWin& gw, CuScratch& C,
                                                SweepState* __restrict__ sweeps, int lane)
{
  const tvc_me_job j0 = gjobs[ct.base];                      // the 2Nx2N PU = the CU
  bool ok = j0.w == ct.S && j0.mode == TVC_ME_TZ && j0.search_range == 64;
  for (int p = 1; p < ct.np; p++) ok &= gjobs[ct.base + p].w > 0;      // every part inside the picture (they share window and start)
  if (!ok) return;
  if (lane < ct.np) C.subp[lane] = (uint8_t)((j0.fen && gjobs[ct.base + lane].h > 8) ? 1 : 0);
  __syncwarp();
  bool need_odd = false;
  for (int p = 0; p < ct.np; p++) need_odd |= C.subp[p] == 0;
  struct { int lx, ty, rx, by; } win = {j0.lx, j0.ty, j0.rx, j0.by};
  const uint8_t* wbase = gw.win + ct.py * kGrpWinW + ct.px + gw.e16;   // window byte of the CU's top-left at candidate (cenx - 64, ceny - 64)
  const uint8_t* cbase = gw.cur + ct.py * kGrpCurP + ct.px;
  auto in_window = [&](int x, int y) { const int dx = x - gw.cenx, dy = y - gw.ceny; return dx >= -kMeR && dx <= kMeR && dy >= -kMeR && dy <= kMeR; };
  auto eval = [&](int x, int y) {
    cu_costs_any(ct.S, wbase + (y - gw.ceny + kMeR) * kGrpWinW + (x - gw.cenx + kMeR), cbase, need_odd, C.subp,
                 mv_cost(j0.lambda_cost, x, y, 2, j0.predx, j0.predy), C.cst[lane]);
  };
  // ---- start point, zero vector (TEncSearch.cpp:4320, 4336-4339)
  {
    const int x = lane == 0 ? j0.startx : 0, y = lane == 0 ? j0.starty : 0;
    if (__any_sync(0xffffffffu, lane < 2 && !in_window(x, y))) return;
    if (lane < 2) eval(x, y);
    __syncwarp();
  }
  uint32_t best = kNoCost, best_dist = 0, n_sads = 2;
  int best_x = j0.startx, best_y = j0.starty, point_nr = 0, best_round = 0;
  if (lane < ct.np) {
    best = C.cst[0][lane];
    if (C.cst[1][lane] < best) { best = C.cst[1][lane]; best_x = 0; best_y = 0; }
  }
  // the sweep is centred on the winner: shared only when every PU of the CU picked the same one
  const int sx = __shfl_sync(0xffffffffu, best_x, 0), sy = __shfl_sync(0xffffffffu, best_y, 0);
  if (__any_sync(0xffffffffu, lane < ct.np && (best_x != sx || best_y != sy))) return;
  __syncwarp();
  bool running = lane < ct.np;
  // ---- rounds d = 1 .. 8 (one candidate per lane), then d = 16, 32 (32 candidates), then d = 64 (16): each stage only while a PU still searches
  for (int stage = 0; stage < 3; stage++) {
    if (!__any_sync(0xffffffffu, running)) break;
    const int c0 = stage == 0 ? 0 : (stage == 1 ? 28 : 60), nc = stage == 0 ? 28 : (stage == 1 ? 32 : 16);
    const int dfirst = stage == 0 ? 1 : (stage == 1 ? 16 : 64), dlast = stage == 0 ? 8 : (stage == 1 ? 32 : 64);
    int x = 0, y = 0, d = 1, i = 0, pt = 0;
    uint32_t dist = 0;
    bool v = false;
    if (lane < nc) { sweep_slot(c0 + lane, 64, d, i); v = diamond_cand(win, sx, sy, d, i, x, y, pt, dist); }
    if (__any_sync(0xffffffffu, v && !in_window(x, y))) return;         // a candidate beyond the staged window: the per-PU code takes the CU
    if (v) eval(x, y);
    const unsigned vmask = __ballot_sync(0xffffffffu, v);
    __syncwarp();
    if (running) {
      int off = 0;
      for (int dd = dfirst; dd <= dlast; dd <<= 1) {
        const int sz = round_size(dd);
        best_round += 1;
        int bi_ = -1;
        for (int k = 0; k < sz; k++) {
          if (!((vmask >> (off + k)) & 1u)) continue;
          n_sads++;
          const uint32_t cc = C.cst[off + k][lane];
          if (cc < best) { best = cc; bi_ = k; }
        }
        if (bi_ >= 0) {
          int ptn;
          diamond_cand(win, sx, sy, dd, bi_, best_x, best_y, ptn, best_dist);
          point_nr = ptn; best_round = 0;
        }
        if (best_round >= 3) { running = false; break; }           // bFirstSearchStop, uiFirstSearchRounds = 3
        off += sz;
      }
    }
    __syncwarp();
  }
  if (lane < ct.np) sweeps[ct.base + lane] = SweepState{best, best_x, best_y, best_dist, point_nr, n_sads};
}

struct GroupMaps {
  CUtensorMap cur;                 // u8 luma of the current picture, box 80 x 64
  CUtensorMap ref[8];              // u8 luma of each reference, box 208 x 192
  const uint8_t* ref8[8];          // the same planes, pel (0,0)
  int stride8;
};
constexpr int kGrpThreads = 256;
constexpr int kGrpSmemSweep = kGrpWinW * kGrpWinH + kGrpCurP * 64 + 64;                       // offset of the SweepState array
constexpr int kGrpSmemCu = kGrpSmemSweep + ((TVC_ME_CENSUS * (int)sizeof(SweepState) + 15) & ~15);  // offset of the per-warp CU scratch
constexpr int kGrpSmemFlags = kGrpSmemCu + (kGrpThreads / 32) * (int)sizeof(CuScratch);          // per-CU "first search done" flags
constexpr int kGrpSmem = kGrpSmemFlags + ((kCuTasks * (int)sizeof(int) + 15) & ~15);

template <int MINB>
__global__ void __launch_bounds__(kGrpThreads, MINB)
k_me_group(const __grid_constant__ GroupMaps maps, PlaneTable pt, int cur_slot, const tvc_me_job* __restrict__ jobs,
           tvc_me_result* __restrict__ out, int pic_w, int pic_h, int mx, int my, int ref_index_fixed, int ctus_x,
           unsigned long long* __restrict__ stats, int split, int cu_stage, int order, int num_refs,
           const int* __restrict__ order_list, uint32_t* __restrict__ cost_out)
{
  extern __shared__ __align__(128) uint8_t gsm[];
  const long long t_start = clock64();
  uint8_t* win = gsm;
  uint8_t* cur = gsm + kGrpWinW * kGrpWinH;
  uint64_t* bar = reinterpret_cast<uint64_t*>(gsm + kGrpWinW * kGrpWinH + kGrpCurP * 64);
  int* next = reinterpret_cast<int*>(bar + 1);            // next[0]: PU counter, next[1]: CU counter
  SweepState* sweeps = reinterpret_cast<SweepState*>(gsm + kGrpSmemSweep);
  CuScratch* cus = reinterpret_cast<CuScratch*>(gsm + kGrpSmemCu);
  volatile int* cu_done = reinterpret_cast<volatile int*>(gsm + kGrpSmemFlags);
  const int tid = threadIdx.x, lane = tid & 31;
  // split > 1: `split` CTAs share one group (each stages the window and takes every split-th PU): the single-group call of
  // tvc_me_ctu is latency-bound, one CTA would walk the 593 PUs alone
  const int part = (int)(blockIdx.x % (unsigned)split);
  // which group this CTA takes.  Groups are laid out [reference][CTU]; the far references cost most (larger motion: more rounds,
  // the raster stage), and in launch order they came last: the kernel's tail was a few SMs finishing the heaviest CTAs while the
  // rest idled (ncu: SMs active 62 % of the kernel's duration).  order 1: last group first; 2: references interleaved;
  // 3: both (a wave mixes the references, the farthest first).  order_list (picture-level calls after the first): the groups by
  // the time each took in the previous call of the same shape, longest first -- a CTU that was expensive against a reference in the
  // last picture usually is again (measured on B200, 1080p x 4 references: launch order 1.95 ms, interleaved 1.71, by last cost see DESIGN)
  int gi = (int)(blockIdx.x / (unsigned)split);
  if (order_list) gi = order_list[gi];
  else {
    const int ng = (int)(gridDim.x / (unsigned)split);
    if (order & 2) {
      const int per = ng / num_refs;                       // CTUs per reference
      if (per * num_refs == ng) gi = (gi % num_refs) * per + gi / num_refs;
    }
    if (order & 1) gi = ng - 1 - gi;
  }
  const size_t gbase = (size_t)gi * TVC_ME_CENSUS;
  const tvc_me_job j0 = jobs[gbase];                       // the 64x64 PU sits at the CTU origin; its start is the window centre
  const int x0 = j0.x, y0 = j0.y;
  const int ref = ref_index_fixed >= 0 ? ref_index_fixed : j0.ref_index;
  // window centre: the group's start point, moved so that the staged window stays inside the padded plane
  int cenx = j0.startx, ceny = j0.starty;
  {
    int lo_x = -mx - x0 + kMeR, hi_x = pic_w + mx - x0 - 64 - kMeR;
    int lo_y = -my - y0 + kMeR, hi_y = pic_h + my - y0 - 64 - kMeR;
    if (hi_x < lo_x) hi_x = lo_x;
    if (hi_y < lo_y) hi_y = lo_y;
    cenx = min(hi_x, max(lo_x, cenx)); ceny = min(hi_y, max(lo_y, ceny));
  }
  const int wx = mx + x0 + cenx - kMeR, e16 = wx & 15;     // TMA wants a 16-byte aligned start: load from wx - e16
  if (tid == 0) {
    next[0] = 0; next[1] = 0;
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar, kGrpWinW * kGrpWinH + kGrpCurP * 64);
    tma_load_2d(win, &maps.ref[ref], wx - e16, my + y0 + ceny - kMeR, bar);
    tma_load_2d(cur, &maps.cur, mx + x0, my + y0, bar);
  }
  for (int i = tid; i < TVC_ME_CENSUS; i += kGrpThreads) sweeps[i] = SweepState{0, 0, 0, 0, 0, 0};
  const bool cu_phase = split == 1 && cu_stage;
  for (int i = tid; i < kCuTasks; i += kGrpThreads) cu_done[i] = cu_phase ? 0 : 1;
  mbar_wait(bar, 0);
  GrpWin gw;
  gw.win = win; gw.cur = cur; gw.plane8 = maps.ref8[ref]; gw.pitch8 = maps.stride8;
  gw.cenx = cenx; gw.ceny = ceny; gw.e16 = e16; gw.x0 = x0; gw.y0 = y0;
  __syncthreads();
  // phase 1: the first search of every CU, one warp per CU, largest first (a group spread over several CTAs skips it: each CTA
  // would repeat the whole phase)
  if (cu_phase) {
    for (;;) {
      int t = 0;
      if (lane == 0) t = atomicAdd(next + 1, 1);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t >= kCuTasks) break;
      cu_first_search(cu_task(t), jobs + gbase, gw, cus[tid >> 5], sweeps, lane);
      __syncwarp();
      if (lane == 0) { __threadfence_block(); cu_done[t] = 1; }
    }
  }
  // phase 2: every PU from where its CU's first search left it (or from scratch), one warp per PU, largest first.  No barrier between
  // the phases: a warp that finds the CU counter exhausted goes on, and a PU whose CU is still being searched by another warp waits for
  // that CU's flag (every CU task has been taken by a running warp by then, and CU tasks never wait: no cycle)
  unsigned long long acc[2] = {0ull, 0ull};
  for (;;) {
    int k = 0;
    if (lane == 0) k = atomicAdd(next, 1);
    k = part + split * __shfl_sync(0xffffffffu, k, 0);
    if (k >= TVC_ME_CENSUS) break;
    const int cu = k < 13 ? 0 : (k < 65 ? 1 + (k - 13) / 13 : (k < 273 ? 5 + (k - 65) / 13 : 21 + (k - 273) / 5));
    while (cu_done[cu] == 0) { }
    __threadfence_block();
    me_search_job((int)(gbase + k), pt, cur_slot, jobs, out, nullptr, nullptr, 0, ctus_x, 0, nullptr, nullptr, stats, &gw, &sweeps[k], acc);
  }
  if (lane == 0 && stats && acc[0]) {
    atomicAdd(&stats[0], acc[0]);
    atomicAdd(&stats[1], acc[1]);
  }
  // the CTA's duration (its last warp's), in units of 16 clocks: the next call's launch order
  if (lane == 0 && cost_out) atomicMax(&cost_out[gi], (uint32_t)((clock64() - t_start) >> 4) + 1u);
}

// groups by descending cost: one CTA, bitonic sort of (cost, index) keys in shared memory (n <= 4096)
__global__ void __launch_bounds__(1024)
k_group_order(const uint32_t* __restrict__ cost, int n, int* __restrict__ order_list)
{
  __shared__ unsigned long long key[4096];
  for (int i = threadIdx.x; i < 4096; i += 1024)
    key[i] = i < n ? (((unsigned long long)(0xffffffffu - cost[i]) << 32) | (unsigned)i) : ~0ull;     // ascending key = descending cost, ties by index
  __syncthreads();
  for (int k = 2; k <= 4096; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < 4096; i += 1024) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = key[i], b = key[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { key[i] = b; key[l] = a; }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < n; i += 1024) order_list[i] = (int)(key[i] & 0xffffffffu);
}

__global__ void k_me_table_lookup(const uint16_t* __restrict__ tables, const tvc_me_center* __restrict__ centers,
                                  int num_ctus, int ctus_x, int ref_index, int pu_x, int pu_y, int pu_w, int pu_h, int fen,
                                  int n, const int16_t* __restrict__ cand, uint32_t* __restrict__ out, int bi)
{
  int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= n) return;
  int lane = threadIdx.x & 31;
  int ctu = (pu_y >> 6) * ctus_x + (pu_x >> 6);
  const uint16_t* tbl = tables + ((size_t)ref_index * num_ctus + ctu) * kMeCtuElems;
  tvc_me_center cen = centers[(size_t)ref_index * num_ctus + ctu];
  int dx = cand[2 * j] - cen.cx, dy = cand[2 * j + 1] - cen.cy;
  int sub = (fen && pu_h > 8) ? 1 : 0;
  uint32_t acc = 0;
  bool inside = dx >= -kMeR && dx <= kMeR && dy >= -kMeR && dy <= kMeR;
  if (inside)
    acc = table_pu_sad_partial(tbl, (pu_x & 63) >> 2, (pu_y & 63) >> 2, pu_w >> 2, pu_h >> 2, sub != 0, dy, dx, lane);
  acc = warp_sum_u32(acc);
  if (lane == 0) out[j] = inside ? ((acc << sub) >> bi) : 0xFFFFFFFFu;
}

// ================================================================================ (3) fractional
// xPatternSearchFracDIF (TEncSearch.cpp:4476-4514).  The reference builds half/quarter planes with
// filterHorLuma(isLast=false) followed by filterVerLuma(isFirst=false, isLast=true)
// (xExtDIFUpSamplingH/Q, :5982-6175) and evaluates 9 + 9 candidates (xPatternRefinement, :711-760).
// Every candidate sample therefore is: 14-bit horizontal stage at fraction fx&3, column shifted by
// (fx<0 ? -1 : 0); vertical stage at fraction fy&3, row shifted by (fy<0 ? -1 : 0); the planes the
// reference materialises are exactly these values (tests/test_gpu_parity.py::test_me_frac pins it).
//
// One job per thread group of NT threads, JPC groups per CTA.  Shared memory per job: the reference
// window R (w+8)x(h+8), the four horizontal planes H[f] over columns -1..w-1 and rows -4..h+3, the
// original block.  Candidate stage: a work unit is (Hadamard tile, horizontal fraction); TS lanes
// each own one tile column, keep the TS+8 rows of H[f] they need in registers (sliding window: one
// shared-memory load per row instead of eight) and produce the three vertical candidates of that
// horizontal fraction, Hadamard in registers (vertical) and shuffles (horizontal).
template <int MAXW, int MAXH>
struct FracSmem {
  static constexpr int RP = MAXW + 8;                  // pitch of R
  static constexpr int HC = MAXW + 1;                  // columns of a horizontal plane (picture columns -1 .. w-1)
  static constexpr int HR = MAXH + 8;                  // rows (picture rows -4 .. h+3)
  static constexpr int CP = HR + 4;                    // column pitch: planes are stored COLUMN-major so that a lane's
                                                       // vertical window is contiguous (64-bit loads, packed int16 pairs)
  int16_t R[RP * (MAXH + 8)];
  alignas(8) int16_t H[4][HC * CP];
  int16_t org[MAXW * MAXH];
  uint32_t cost[12];
  int sel[4];
};

// xPatternRefinement candidate orders (TEncSearch.cpp:47-71)
__constant__ int8_t c_refine_h[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {1, -1}, {-1, 1}, {1, 1}};
__constant__ int8_t c_refine_q[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {1, 1}};

// candidate index (in the reference's order) of offset (ox, oy) in {-1,0,1}^2 for the half / quarter pass
__device__ __forceinline__ int refine_index(bool half, int ox, int oy)
{
  // half:    (0,0) (0,-1) (0,1) (-1,0) (1,0) (-1,-1) (1,-1) (-1,1) (1,1)
  // quarter: (0,0) (0,-1) (0,1) (-1,-1) (1,-1) (-1,0) (1,0) (-1,1) (1,1)
  const int key = (oy + 1) * 3 + (ox + 1);      // 0..8: (-1,-1) (0,-1) (1,-1) (-1,0) (0,0) (1,0) (-1,1) (0,1) (1,1)
  const unsigned long long th = 0x827403615ull, tq = 0x827605413ull;   // nibble per key, LSB first
  return (int)(((half ? th : tq) >> (4 * key)) & 15);
}

// Two difference tiles through ONE Hadamard: P = dB * 65536 + dA is exact integer arithmetic under the butterflies (they are
// linear), and for 8-bit content every coefficient stays inside int16 (|.| <= 64 * 510 = 32640 even for bi-prediction targets), so the two results are
// recovered exactly from the halves.  One third of the fractional search's Hadamard work disappears (3 candidates -> 2 passes).
template <int TS>
__device__ __forceinline__ void had_cols2(const int (&dA)[TS], const int (&dB)[TS], int c, uint32_t& outA, uint32_t& outB)
{
  int d[TS];
#pragma unroll
  for (int r = 0; r < TS; r++) d[r] = dB[r] * 65536 + dA[r];
#pragma unroll
  for (int len = 1; len < TS; len <<= 1)
#pragma unroll
    for (int i = 0; i < TS; i += 2 * len)
#pragma unroll
      for (int k = i; k < i + len; k++) { int a = d[k], b = d[k + len]; d[k] = a + b; d[k + len] = a - b; }
#pragma unroll
  for (int m = 1; m < TS; m <<= 1) {
    const int sgn = (c & m) ? -1 : 1;
#pragma unroll
    for (int r = 0; r < TS; r++) d[r] = __shfl_xor_sync(0xffffffffu, d[r], m) + sgn * d[r];
  }
  uint32_t sa = 0, sb = 0;
#pragma unroll
  for (int r = 0; r < TS; r++) {
    const int lo = (int)(int16_t)(d[r] & 0xffff);
    sa = __sad(lo, 0, sa);
    sb = __sad((d[r] - lo) >> 16, 0, sb);
  }
#pragma unroll
  for (int m = 1; m < TS; m <<= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, m); sb += __shfl_xor_sync(0xffffffffu, sb, m); }
  outA = TS == 8 ? ((sa + 2) >> 2) : ((sa + 1) >> 1);
  outB = TS == 8 ? ((sb + 2) >> 2) : ((sb + 1) >> 1);
}

template <int TS>
__device__ __forceinline__ uint32_t had_cols(int (&d)[TS], int c)
{
  // lane c of a TS-lane group holds column c of a TS x TS difference tile: 2-D Hadamard, sum |.|, per-tile rounding
#pragma unroll
  for (int len = 1; len < TS; len <<= 1)
#pragma unroll
    for (int i = 0; i < TS; i += 2 * len)
#pragma unroll
      for (int k = i; k < i + len; k++) { int a = d[k], b = d[k + len]; d[k] = a + b; d[k + len] = a - b; }
#pragma unroll
  for (int m = 1; m < TS; m <<= 1) {
    // butterfly across lanes: the lane with bit m clear gets mine + other, its partner other - mine:
    // other + sgn * mine, one multiply-add per element
    const int sgn = (c & m) ? -1 : 1;
#pragma unroll
    for (int r = 0; r < TS; r++) d[r] = __shfl_xor_sync(0xffffffffu, d[r], m) + sgn * d[r];
  }
  uint32_t s = 0;
#pragma unroll
  for (int r = 0; r < TS; r++) s = __sad(d[r], 0, s);
#pragma unroll
  for (int m = 1; m < TS; m <<= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  return TS == 8 ? ((s + 2) >> 2) : ((s + 1) >> 1);      // xCalcHADs8x8 / 4x4 rounding (TComRdCost.cpp:1773,1869)
}

// one work unit: tile (tx,ty), horizontal candidate fx; produces the distortions of fy = fy0 - dq, fy0, fy0 + dq
// second-stage rounding of TComInterpolationFilter::filter (isFirst = false, isLast = true,
// TComInterpolationFilter.cpp:204-238) and of filterCopy (:124-145) with the shifts / offsets of the job's
// bit depth folded into constants: (Short)((sum + offset) >> shift), then clip to [0, maxv]
struct VRound { int shift, offset, cshift, coffset, maxv; bool pack2; };
__device__ __forceinline__ VRound make_vround(int bd)
{
  VRound v;
  const int head = kIfPrec - bd;
  v.shift = kIfFilt + head; v.offset = (1 << (v.shift - 1)) + (kIfOffs << kIfFilt);
  v.cshift = head; v.coffset = kIfOffs + (head ? (1 << (head - 1)) : 0);
  v.maxv = (1 << bd) - 1;
  v.pack2 = false;
  return v;
}
// The reference casts to Short before clipping.  Here the operands are this kernel's own first-stage
// outputs (|H| <= 12272 << (bd-8) ... bounded by the pel range), so (sum + offset) >> shift lies far inside
// the int16 range and the cast is the identity; the clip is one min-with-relu instruction.
__device__ __forceinline__ int vround_filter(int sum_plus_offset, const VRound& v)
{
  return __vimin_s32_relu(sum_plus_offset >> v.shift, v.maxv);
}
__device__ __forceinline__ int vround_copy(int smp, const VRound& v)
{
  return __vimin_s32_relu((smp + v.coffset) >> v.cshift, v.maxv);
}

// 8 luma taps of fraction g packed as signed bytes: TA = c0..c3, TB = c4..c7 (operands of dp2a.lo / dp2a.hi)
__device__ __forceinline__ void packed_taps(int g, int& ta, int& tb)
{
  // c_luma_taps is int8[4][8], 8-byte aligned: the eight taps of a fraction are two little-endian words already
  const int2 t = reinterpret_cast<const int2*>(c_luma_taps)[g];
  ta = t.x; tb = t.y;
}

template <int TS, int CP>
__device__ __forceinline__ void frac_unit(const int16_t* __restrict__ H, int plane_elems, const int16_t* __restrict__ org, int opitch,
                                          int fx, int fy0, int dq, int tx, int ty, int c, bool hadamard, const VRound& vr, uint32_t (&out)[3])
{
  const int ix = fx < 0 ? -1 : 0, f = fx & 3;
  // column tx+c+ix of the plane (column index 0 is picture column -1), rows ty-4 .. ty+TS+3 (row index 0 is picture
  // row -4): TS+8 consecutive int16 = (TS+8)/2 packed pairs; ty is a multiple of 4 and CP of 4: 8-byte aligned
  constexpr int NW = (TS + 8) / 2;
  const uint2* p = reinterpret_cast<const uint2*>(H + f * plane_elems + (tx + c + ix + 1) * CP + ty);
  int W[NW];                                   // W[j] = rows (2j, 2j+1) of the window
#pragma unroll
  for (int j = 0; j < NW / 2; j++) { const uint2 q = p[j]; W[2 * j] = (int)q.x; W[2 * j + 1] = (int)q.y; }
  int P[NW - 1];                               // P[j] = rows (2j+1, 2j+2)
#pragma unroll
  for (int j = 0; j < NW - 1; j++) P[j] = (int)__funnelshift_r((unsigned)W[j], (unsigned)W[j + 1], 16);
  int o[TS];
#pragma unroll
  for (int r = 0; r < TS; r++) o[r] = org[(ty + r) * opitch + tx + c];
  // difference column of candidate k (fy = fy0 + (k - 1) * dq)
  auto diff = [&](int k, int (&d)[TS]) {
    const int fy = fy0 + (k - 1) * dq;
    const bool up = fy < 0;                                                  // row shift -1
    const int g = fy & 3;
    // fy is the same for every thread of the CTA (it depends on the pass and on k only): uniform branches
    if (g == 0) {
      // integer row: the reference takes the filterCopy branch (TComInterpolationFilter.cpp:124-145); fy == 0 has
      // no row shift, so the sample is window row r + 4 (dp2a with taps (1,0) / (0,1) extracts a sign-extended half)
#pragma unroll
      for (int r = 0; r < TS; r++) {
        const int smp = ((r + 4) & 1) ? __dp2a_lo(W[(r + 4) >> 1], 0x0100, 0) : __dp2a_lo(W[(r + 4) >> 1], 0x0001, 0);
        d[r] = o[r] - vround_copy(smp, vr);
      }
    } else {
      int ta, tb;
      packed_taps(g, ta, tb);
      // output r filters window rows s0 .. s0+7 with s0 = r (row shift -1) or r + 1: four dp2a on packed row pairs
#pragma unroll
      for (int r = 0; r < TS; r++) {
        int sum = vr.offset;
        if (up) {
          if (r & 1) { const int j = (r - 1) >> 1; sum = __dp2a_lo(P[j], ta, sum); sum = __dp2a_hi(P[j + 1], ta, sum); sum = __dp2a_lo(P[j + 2], tb, sum); sum = __dp2a_hi(P[j + 3], tb, sum); }
          else { const int j = r >> 1; sum = __dp2a_lo(W[j], ta, sum); sum = __dp2a_hi(W[j + 1], ta, sum); sum = __dp2a_lo(W[j + 2], tb, sum); sum = __dp2a_hi(W[j + 3], tb, sum); }
        } else {
          if (r & 1) { const int j = (r + 1) >> 1; sum = __dp2a_lo(W[j], ta, sum); sum = __dp2a_hi(W[j + 1], ta, sum); sum = __dp2a_lo(W[j + 2], tb, sum); sum = __dp2a_hi(W[j + 3], tb, sum); }
          else { const int j = r >> 1; sum = __dp2a_lo(P[j], ta, sum); sum = __dp2a_hi(P[j + 1], ta, sum); sum = __dp2a_lo(P[j + 2], tb, sum); sum = __dp2a_hi(P[j + 3], tb, sum); }
        }
        d[r] = o[r] - vround_filter(sum, vr);
      }
    }
  };
  if (hadamard && vr.pack2) {
    // 8-bit content: candidates 0 and 1 share one Hadamard (had_cols2), candidate 2 takes its own
    int d0[TS], d1[TS];
    diff(0, d0);
    diff(1, d1);
    had_cols2<TS>(d0, d1, c, out[0], out[1]);
    diff(2, d0);
    out[2] = had_cols<TS>(d0, c);
    return;
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    int d[TS];
    diff(k, d);
    if (hadamard) out[k] = had_cols<TS>(d, c);
    else {
      uint32_t s = 0;
#pragma unroll
      for (int r = 0; r < TS; r++) s = __sad(d[r], 0, s);
#pragma unroll
      for (int m = 1; m < TS; m <<= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
      out[k] = s;
    }
  }
}

// one job by one thread group (NT threads: the CTA when JPC == 1, else a warp).  LISTED: called from the loop of k_me_frac_listed
// (the job is known to be unserved; a group returns on its own).
template <int MAXW, int MAXH, int NT, int JPC, bool LISTED>
__device__ __forceinline__ void frac_job_run(const PlaneTable& pt, int cur_slot, const tvc_frac_job* __restrict__ jobs,
                                             tvc_frac_result* __restrict__ out, int bd, FracSmem<MAXW, MAXH>& S, int tid, int ji, bool have,
                                             const uint8_t* __restrict__ done)
{
  using SM = FracSmem<MAXW, MAXH>;
  auto sync = [&]() { if (JPC == 1) __syncthreads(); else __syncwarp(); };
  const tvc_frac_job jb = jobs[ji];
  const bool served = !LISTED && have && done && done[ji];  // the CU-level kernel (k_me_frac_cu) already wrote this job's result
  const bool live = have && !served && jb.w > 0 && jb.w <= MAXW && jb.h <= MAXH;
  if (!LISTED && JPC > 1 && done) {                        // every job of this CTA already served: leave at once
    if (__syncthreads_and(!have || served)) return;
  }
  if (!live) {
    if (have && !served && tid == 0 && jb.w <= 0) out[ji] = tvc_frac_result{0, 0, 0, 0, 0u, 0u};   // census PU outside the picture
    if (JPC == 1 || LISTED) return;
  }
  const int w = live ? jb.w : 4, h = live ? jb.h : 4, bi = bd - 8;
  const int gstride = pt.stride[0];
  int wide = 0;
  if (live) {
    const int16_t* ref = pt.org[jb.ref_slot][0] + (ptrdiff_t)(jb.y + jb.imvy - 4) * gstride + jb.x + jb.imvx - 4;
    const int16_t* cur = pt.org[cur_slot][0] + (ptrdiff_t)jb.y * gstride + jb.x;
    // n / d for the small loop counters below as one multiply + shift: m = ceil(2^20 / d) is exact for n < 2^20 / d (n <= 5184, d <= 72);
    // the integer divisions were 10-18 % of the small size classes' instructions in the ncu source view
    const int rw = w + 8, rh = h + 8;
    const unsigned m_rw = ((1u << 20) + rw - 1) / rw, m_w = ((1u << 20) + w - 1) / w;
    for (int i = tid; i < rw * rh; i += NT) {
      int r = (int)(((unsigned)i * m_rw) >> 20), x = i - r * rw;
      S.R[r * SM::RP + x] = ref[(ptrdiff_t)r * gstride + x];
    }
    for (int i = tid; i < w * h; i += NT) {
      int r = (int)(((unsigned)i * m_w) >> 20), x = i - r * w;
      const int16_t ov = cur[(ptrdiff_t)r * gstride + x];
      S.org[r * MAXW + x] = ov;
      wide |= (ov < -255) | (ov > 510);        // beyond an 8-bit bi-prediction target (2 * org - pred): |d| could exceed 510
    }
    if (tid < 12) S.cost[tid] = 0;
    if (tid == 0) S.sel[2] = 0;
  }
  sync();
  if (live) {
    if (wide) S.sel[2] = 1;                    // every writer stores the same value; read after the next barrier
    // horizontal stage (isFirst, !isLast): H[f][r][xi], xi = 0..w <-> picture column xi-1, r = 0..h+7 <-> row r-4
    const int hw = w + 1, hh = h + 8;
    const unsigned m_hw = ((1u << 20) + hw - 1) / hw;
    for (int i = tid; i < hw * hh; i += NT) {
      int r = (int)(((unsigned)i * m_hw) >> 20), xi = i - r * hw;
      const int16_t* p = &S.R[r * SM::RP + xi];          // taps: picture columns (xi-1)-3 .. (xi-1)+4 = R columns xi .. xi+7
      int t[8];
#pragma unroll
      for (int k = 0; k < 8; k++) t[k] = p[k];
      S.H[0][xi * SM::CP + r] = if_copy(t[3], true, false, bd);
#pragma unroll
      for (int f = 1; f < 4; f++) {
        int sum = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) sum += t[k] * (int)c_luma_taps[f][k];
        S.H[f][xi * SM::CP + r] = if_round(sum, true, false, bd);
      }
    }
  }
  sync();

  VRound vr = make_vround(bd);
  // two candidates per Hadamard pass need every coefficient inside int16: |d| <= 510 (64 * 510 = 32640), i.e. 8-bit content
  // including bi-prediction targets
  vr.pack2 = bd == 8 && live && S.sel[2] == 0;
  const bool t8 = jb.hadamard && ((w & 7) == 0) && ((h & 7) == 0);
  const int TS = t8 ? 8 : 4;
  const int tiles_x = t8 ? (w >> 3) : (w >> 2), tiles = tiles_x * (t8 ? (h >> 3) : (h >> 2));
  const int groups = t8 ? NT / 8 : NT / 4, ug = t8 ? (tid >> 3) : (tid >> 2), c = t8 ? (tid & 7) : (tid & 3);
  const unsigned m_tiles = ((1u << 20) + tiles - 1) / tiles, m_tx = ((1u << 20) + tiles_x - 1) / tiles_x;
  int basex = 0, basey = 0, hx = 0, hy = 0;
  uint32_t cost_half = 0;
  for (int pass = 0; pass < 2; pass++) {
    const int dq = pass == 0 ? 2 : 1;
    if (live) {
      const int units = 3 * tiles;
      const int iters = t8 ? (units + NT / 8 - 1) / (NT / 8) : (units + NT / 4 - 1) / (NT / 4);      // rounded up: every lane of a warp executes the shuffles
      for (int it = 0; it < iters; it++) {
        const int u = it * groups + ug;
        const bool valid = u < units;
        const int uu = valid ? u : 0;
        const int fxi = (int)(((unsigned)uu * m_tiles) >> 20), t = uu - fxi * tiles;
        const int trow = (int)(((unsigned)t * m_tx) >> 20);
        const int tyy = trow * TS, txx = (t - trow * tiles_x) * TS;
        const int ox = fxi - 1;
        uint32_t v[3];
        if (t8) frac_unit<8, SM::CP>(&S.H[0][0], SM::HC * SM::CP, S.org, MAXW, basex + ox * dq, basey, dq, txx, tyy, c, true, vr, v);
        else frac_unit<4, SM::CP>(&S.H[0][0], SM::HC * SM::CP, S.org, MAXW, basex + ox * dq, basey, dq, txx, tyy, c, jb.hadamard != 0, vr, v);
        if (valid && c == 0) {
#pragma unroll
          for (int k = 0; k < 3; k++) atomicAdd(&S.cost[refine_index(pass == 0, ox, k - 1)], v[k]);
        }
      }
    }
    sync();
    if (tid < 32) {
      // xPatternRefinement (TEncSearch.cpp:730-757): dist >> bitIncrement + rate, strict '<' in the listed order =
      // the first candidate that attains the minimum.  Lane i holds candidate i.
      const int scale = pass == 0 ? 1 : 0;
      const int ax = pass == 0 ? (jb.imvx << 1) : (((jb.imvx << 1) + hx) << 1);
      const int ay = pass == 0 ? (jb.imvy << 1) : (((jb.imvy << 1) + hy) << 1);
      const int i = tid < 9 ? tid : 0;
      const int8_t* rf = pass == 0 ? c_refine_h[i] : c_refine_q[i];
      uint32_t dcost = 0xFFFFFFFFu;
      if (live && tid < 9) dcost = (S.cost[i] >> bi) + mv_cost(jb.lambda_cost, ax + rf[0], ay + rf[1], scale, jb.predx, jb.predy);
      const uint32_t best = __reduce_min_sync(0xffffffffu, dcost);
      const int best_i = __ffs(__ballot_sync(0xffffffffu, dcost == best)) - 1;
      __syncwarp();
      if (live && tid == best_i) { S.sel[0] = rf[0]; S.sel[1] = rf[1]; S.cost[9] = best; }
      if (live && tid < 9) S.cost[tid] = 0;
    }
    sync();
    if (pass == 0) {
      hx = S.sel[0]; hy = S.sel[1]; cost_half = S.cost[9];
      basex = hx * 2; basey = hy * 2;
    } else if (live && tid == 0) {
      tvc_frac_result r;
      r.halfx = hx; r.halfy = hy; r.qtrx = S.sel[0]; r.qtry = S.sel[1];
      r.cost_half = cost_half; r.cost = S.cost[9];
      out[ji] = r;
    }
    sync();
  }
}

template <int MAXW, int MAXH, int NT, int JPC>
__global__ void __launch_bounds__(NT * JPC)
k_me_frac(PlaneTable pt, int cur_slot, int n, const tvc_frac_job* __restrict__ jobs, tvc_frac_result* __restrict__ out, int bd,
          int span, int stride, int first, const uint8_t* __restrict__ done)
{
  using SM = FracSmem<MAXW, MAXH>;
  extern __shared__ __align__(16) uint8_t fsm[];
  const int grp = threadIdx.x / NT, tid = threadIdx.x % NT;
  SM& S = reinterpret_cast<SM*>(fsm)[grp];
  // job index: the frame pre-pass launches one census size class at a time (span jobs out of every stride)
  const int b = blockIdx.x * JPC + grp;
  const bool have = b < n;
  const int ji = have ? (b / span) * stride + first + (b % span) : 0;
  frac_job_run<MAXW, MAXH, NT, JPC, false>(pt, cur_slot, jobs, out, bd, S, tid, ji, have, done);
}

// The census jobs the CU-level kernels left unserved (`done` == 0: CUs whose PUs found different integer vectors, CUs at the picture
// border; 9 % of the jobs of a 1080p picture), as four lists by census size class.  With one CTA per job (or four) over ALL jobs the
// per-PU kernels launched 402 000 CTAs per picture of which nine in ten read a flag and left: 0.95 ms at the CTA launch rate.
__global__ void __launch_bounds__(256)
k_frac_rest_lists(const uint8_t* __restrict__ done, int n, int groups, int* __restrict__ lists, int* __restrict__ counts)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
  const bool rest = i < n && done[i] == 0;
  const int k = i < n ? i % TVC_ME_CENSUS : 0;
  const int cls = k < 13 ? 0 : (k < 65 ? 1 : (k < 273 ? 2 : 3));
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const unsigned m = __ballot_sync(0xffffffffu, rest && cls == c);
    if (!m) continue;
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(&counts[c], __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    const int first = c == 0 ? 0 : (c == 1 ? 13 : (c == 2 ? 65 : 273));
    if (rest && cls == c) lists[(size_t)groups * first + base + __popc(m & ((1u << lane) - 1))] = i;
  }
}

// the listed jobs of one size class: a grid that fills the machine once, thread group g takes list entries g, g + G, g + 2G, ...
template <int MAXW, int MAXH, int NT, int JPC>
__global__ void __launch_bounds__(NT * JPC)
k_me_frac_listed(PlaneTable pt, int cur_slot, const int* __restrict__ list, const int* __restrict__ count, const tvc_frac_job* __restrict__ jobs,
                 tvc_frac_result* __restrict__ out, int bd)
{
  using SM = FracSmem<MAXW, MAXH>;
  extern __shared__ __align__(16) uint8_t fsm[];
  const int grp = threadIdx.x / NT, tid = threadIdx.x % NT;
  SM& S = reinterpret_cast<SM*>(fsm)[grp];
  const int n = *count;
  for (int i = blockIdx.x * JPC + grp; i < n; i += gridDim.x * JPC) {
    frac_job_run<MAXW, MAXH, NT, JPC, true>(pt, cur_slot, jobs, out, bd, S, tid, list[i], true, nullptr);
    if (JPC == 1) __syncthreads(); else __syncwarp();
  }
}

// ================================================================================ (3b) fractional search, CU level
// In the census every CU carries 13 PUs (5 at 8x8) that tile the same block seven (three) times over.  When the integer search
// gave all of them the SAME vector -- the common case: coherent motion inside a CU -- they share the reference window, the
// horizontal planes and every Hadamard tile: a tile's SATD at a fractional candidate is one number whichever PU sums it.  One CTA
// (thread group) per CU then runs xPatternSearchFracDIF once over the CU: pass 1 evaluates the nine half-sample candidates tile by
// tile and adds each tile into every PU that contains it (8x8 tiles for PUs whose sides are multiples of 8, 4x4 tiles for the
// others, as xGetHADs tiles them, TComRdCost.cpp:2186-2287), each PU picks its own best half offset (xPatternRefinement with its
// own sums); pass 2 does the same for the nine quarter-sample candidates around every DISTINCT half offset the PUs chose, tiles
// restricted to the PUs that chose it.  CUs whose PUs disagree (or that stick out of the picture) are left to the per-PU kernel
// (`done` stays 0).  Results are identical: the sums are the same integers added in another order.
constexpr int kCuMaxPu = 33;        // a 16x16 CU with its four 8x8 children: 13 + 4 x 5 PUs

template <int CUW>
struct FracCuSmem {
  FracSmem<CUW, CUW> f;
  uint32_t cost[kCuMaxPu][12];
  int8_t px[kCuMaxPu], py[kCuMaxPu], pw[kCuMaxPu], ph[kCuMaxPu];      // PU rectangles inside the CU
  int8_t hx[kCuMaxPu], hy[kCuMaxPu], t8[kCuMaxPu];
  uint32_t cost_half[kCuMaxPu];
  int jidx[kCuMaxPu];             // census job of each PU
  unsigned long long mask8[(CUW / 8) * (CUW / 8)], mask4[(CUW / 4) * (CUW / 4)];      // per Hadamard tile: the PUs that contain it
  int agree, np;
};

template <int CUW, int NT, int JPC, bool CHILD>     // CHILD (16x16 only): the CU's four 8x8 children join when they share the vector too
__global__ void __launch_bounds__(NT * JPC)
k_me_frac_cu(PlaneTable pt, int cur_slot, int ncu, const tvc_frac_job* __restrict__ jobs, tvc_frac_result* __restrict__ out,
             uint8_t* __restrict__ done, int bd, int cu_per_group, int np, int first, unsigned long long* __restrict__ stats)
{
  using SM = FracCuSmem<CUW>;
  using FS = FracSmem<CUW, CUW>;
  extern __shared__ __align__(16) uint8_t fsm[];
  const int grp = threadIdx.x / NT, tid = threadIdx.x % NT;
  SM& S = reinterpret_cast<SM*>(fsm)[grp];
  const int b = blockIdx.x * JPC + grp;
  const bool have = b < ncu;
  const int base = have ? (b / cu_per_group) * TVC_ME_CENSUS + first + (b % cu_per_group) * np : 0;
  auto sync = [&]() { if (JPC == 1) __syncthreads(); else __syncwarp(); };
  const tvc_frac_job j0 = jobs[base];                    // the 2Nx2N PU = the CU
  if (tid == 0) { S.agree = have ? 1 : 0; S.np = CHILD ? np + 20 : np; }
  sync();
  // census job of PU k: the CU's own parts, then (CHILD) the five parts of each 8x8 child (depth 3 of the census: 8 x 8 CUs in
  // raster order from index 273, five PUs each; the 16x16 CU `cu` of the 4 x 4 grid covers children (2 cx + i, 2 cy + j))
  const int npx = CHILD ? np + 20 : np;
  if (have && !(done[base])) {
    for (int k = tid; k < npx; k += NT) {
      int ji = base + k;
      if (CHILD && k >= np) {
        const int cu = b % cu_per_group, cx = cu & 3, cy = cu >> 2, ch = (k - np) / 5, part = (k - np) % 5;
        ji = (b / cu_per_group) * TVC_ME_CENSUS + 273 + (((2 * cy + (ch >> 1)) * 8) + 2 * cx + (ch & 1)) * 5 + part;
      }
      const tvc_frac_job jk = jobs[ji];
      const bool same = jk.w > 0 && j0.w == CUW && jk.imvx == j0.imvx && jk.imvy == j0.imvy;
      if (!same) { if (k < np) S.agree = 0; else S.np = np; }      // a child that differs: the CU goes alone (benign races: same value from every writer)
      S.jidx[k] = ji;
      S.px[k] = (int8_t)(jk.x - j0.x); S.py[k] = (int8_t)(jk.y - j0.y); S.pw[k] = (int8_t)jk.w; S.ph[k] = (int8_t)jk.h;
      S.t8[k] = (jk.hadamard && !(jk.w & 7) && !(jk.h & 7)) ? 1 : 0;
      S.hx[k] = 0; S.hy[k] = 0;
#pragma unroll
      for (int i = 0; i < 12; i++) S.cost[k][i] = 0;
    }
  } else if (tid == 0) S.agree = 0;              // already served (an 8x8 CU taken by its parent's launch)
  sync();
  const bool live = have && S.agree != 0;
  if (!live && JPC == 1) return;
  np = live ? S.np : 0;                          // PUs served here: the CU's own, plus its children when they all agree
  for (int t = tid; t < (CUW / 8) * (CUW / 8) + (CUW / 4) * (CUW / 4); t += NT) {
    const bool is8 = t < (CUW / 8) * (CUW / 8);
    const int tt = is8 ? t : t - (CUW / 8) * (CUW / 8), n = is8 ? CUW / 8 : CUW / 4, ts = is8 ? 8 : 4;
    const int tx = (tt % n) * ts, ty = (tt / n) * ts;
    unsigned long long m = 0;
    for (int k = 0; k < np; k++)
      if ((S.t8[k] != 0) == is8 && tx >= S.px[k] && tx < S.px[k] + S.pw[k] && ty >= S.py[k] && ty < S.py[k] + S.ph[k]) m |= 1ull << k;
    if (is8) S.mask8[tt] = m; else S.mask4[tt] = m;
  }
  const int w = CUW, h = CUW, bi = bd - 8;
  const int gstride = pt.stride[0];
  int wide = 0;
  if (tid == 0) S.f.sel[2] = 0;
  if (live) {
    const int16_t* ref = pt.org[j0.ref_slot][0] + (ptrdiff_t)(j0.y + j0.imvy - 4) * gstride + j0.x + j0.imvx - 4;
    const int16_t* cur = pt.org[cur_slot][0] + (ptrdiff_t)j0.y * gstride + j0.x;
    constexpr int rw = CUW + 8;
    for (int i = tid; i < rw * rw; i += NT) {
      const int r = i / rw, x = i - r * rw;
      S.f.R[r * FS::RP + x] = ref[(ptrdiff_t)r * gstride + x];
    }
    for (int i = tid; i < w * h; i += NT) {
      const int r = i / CUW, x = i - r * CUW;
      const int16_t ov = cur[(ptrdiff_t)r * gstride + x];
      S.f.org[r * CUW + x] = ov;
      wide |= (ov < -255) | (ov > 510);
    }
  }
  sync();
  if (live) {
    if (wide) S.f.sel[2] = 1;
    constexpr int hw = CUW + 1, hh = CUW + 8;
    for (int i = tid; i < hw * hh; i += NT) {
      const int r = i / hw, xi = i - r * hw;
      const int16_t* p = &S.f.R[r * FS::RP + xi];
      int t[8];
#pragma unroll
      for (int k = 0; k < 8; k++) t[k] = p[k];
      S.f.H[0][xi * FS::CP + r] = if_copy(t[3], true, false, bd);
#pragma unroll
      for (int f = 1; f < 4; f++) {
        int sum = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) sum += t[k] * (int)c_luma_taps[f][k];
        S.f.H[f][xi * FS::CP + r] = if_round(sum, true, false, bd);
      }
    }
  }
  sync();
  VRound vr = make_vround(bd);
  vr.pack2 = bd == 8 && live && S.f.sel[2] == 0;
  const bool hadamard = j0.hadamard != 0;
  bool any8 = false, any4 = false;
  if (live)
    for (int k = 0; k < np; k++) { any8 |= S.t8[k] != 0; any4 |= S.t8[k] == 0; }

  for (int pass = 0; pass < 2; pass++) {
    const int dq = pass == 0 ? 2 : 1;
    if (live) {
      // pass 1: one base (the integer vector); pass 2: every distinct half offset the PUs picked, in xPatternRefinement's order
      for (int hb = 0; hb < (pass == 0 ? 1 : 9); hb++) {
        const int bhx = pass == 0 ? 0 : c_refine_h[hb][0], bhy = pass == 0 ? 0 : c_refine_h[hb][1];
        unsigned long long m8 = 0, m4 = 0;            // PUs that take part in this base, by tile type
        for (int k = 0; k < np; k++)
          if (pass == 0 || (S.hx[k] == bhx && S.hy[k] == bhy)) { if (S.t8[k]) m8 |= 1ull << k; else m4 |= 1ull << k; }
        if (!(m8 | m4)) continue;
        const int basex = bhx * 2, basey = bhy * 2;
        for (int tt = 0; tt < 2; tt++) {
          const unsigned long long mem = tt == 0 ? m8 : m4;
          if (!mem) continue;
          const int TS = tt == 0 ? 8 : 4;
          const int tiles_x = CUW / TS, tiles = tiles_x * tiles_x, units = 3 * tiles;
          const int groups = NT / TS, ug = tid / TS, c = tid % TS;
          const int iters = (units + groups - 1) / groups;
          for (int it = 0; it < iters; it++) {
            const int u = it * groups + ug;
            const bool valid = u < units;
            const int uu = valid ? u : 0;
            const int fxi = uu / tiles, t = uu - fxi * tiles;
            const int trow = t / tiles_x, tyy = trow * TS, txx = (t - trow * tiles_x) * TS;
            // PUs of this base and tile type that contain the tile
            const unsigned long long hit = mem & (tt == 0 ? S.mask8[t] : S.mask4[t]);
            const int ox = fxi - 1;
            uint32_t v[3] = {0, 0, 0};
            // frac_unit shuffles with the full-warp mask: the whole warp takes it or skips it (pass 2: tiles outside every PU of this base)
            if (__any_sync(0xffffffffu, hit != 0)) {
              if (tt == 0) frac_unit<8, FS::CP>(&S.f.H[0][0], FS::HC * FS::CP, S.f.org, CUW, basex + ox * dq, basey, dq, txx, tyy, c, true, vr, v);
              else frac_unit<4, FS::CP>(&S.f.H[0][0], FS::HC * FS::CP, S.f.org, CUW, basex + ox * dq, basey, dq, txx, tyy, c, hadamard, vr, v);
            }
            if (valid && c == 0) {
              for (unsigned long long hm = hit; hm; hm &= hm - 1) {
                const int k = __ffsll((long long)hm) - 1;
#pragma unroll
                for (int q = 0; q < 3; q++) atomicAdd(&S.cost[k][refine_index(pass == 0, ox, q - 1)], v[q]);
              }
            }
          }
        }
      }
    }
    sync();
    // xPatternRefinement (TEncSearch.cpp:730-757), one lane per PU: dist >> bitIncrement + rate, strict '<' in the listed order
    if (tid < 32 && live) {
      for (int k = tid; k < np; k += 32) {
        const int scale = pass == 0 ? 1 : 0;
        const int phx = S.hx[k], phy = S.hy[k];
        const int ax = pass == 0 ? (j0.imvx << 1) : (((j0.imvx << 1) + phx) << 1);
        const int ay = pass == 0 ? (j0.imvy << 1) : (((j0.imvy << 1) + phy) << 1);
        uint32_t best = 0xFFFFFFFFu;
        int best_i = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) {
          const int8_t* rf = pass == 0 ? c_refine_h[i] : c_refine_q[i];
          const uint32_t dcost = (S.cost[k][i] >> bi) + mv_cost(j0.lambda_cost, ax + rf[0], ay + rf[1], scale, j0.predx, j0.predy);
          if (dcost < best) { best = dcost; best_i = i; }
          S.cost[k][i] = 0;
        }
        const int8_t* rf = pass == 0 ? c_refine_h[best_i] : c_refine_q[best_i];
        if (pass == 0) { S.hx[k] = rf[0]; S.hy[k] = rf[1]; S.cost_half[k] = best; }
        else {
          tvc_frac_result r;
          r.halfx = phx; r.halfy = phy; r.qtrx = rf[0]; r.qtry = rf[1];
          r.cost_half = S.cost_half[k]; r.cost = best;
          out[S.jidx[k]] = r;
          done[S.jidx[k]] = 1;
        }
      }
      if (pass == 1 && tid == 0 && stats) atomicAdd(&stats[2], (unsigned long long)np);
    }
    sync();
  }
}

// ================================================================================ (4) frame pre-pass
static void build_census(tvc_census_pu* out)
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
          out[n].x = (int16_t)(cx + parts[k][0]); out[n].y = (int16_t)(cy + parts[k][1]);
          out[n].w = (int16_t)parts[k][2]; out[n].h = (int16_t)parts[k][3];
          out[n].cu_x = (int16_t)cx; out[n].cu_y = (int16_t)cy;
          n++;
        }
      }
  }
}

// TComDataCU::clipMv (TComDataCU.cpp:3505-3517), quarter pels
__device__ __forceinline__ void clip_mv(int pic_w, int pic_h, int cu_x, int cu_y, int& x, int& y)
{
  int hmax = (pic_w + 8 - cu_x - 1) * 4, hmin = (-64 - 8 - cu_x + 1) * 4;
  int vmax = (pic_h + 8 - cu_y - 1) * 4, vmin = (-64 - 8 - cu_y + 1) * 4;
  x = min(hmax, max(hmin, x));
  y = min(vmax, max(vmin, y));
}

// census PU k of CTU `ctu` against one reference with predictor p (quarter pels): the job xMotionEstimation would hand to
// xTZSearch (window by xSetSearchRange, start = clipped predictor >> 2)
__device__ __forceinline__ tvc_me_job census_job(int pic_w, int pic_h, int ctus_x, int ctu, int k, int ref_index, int ref_slot,
                                                 tvc_me_center p, const tvc_me_frame_cfg& cfg)
{
  const tvc_census_pu cp = c_census[k];
  int x0 = (ctu % ctus_x) * 64, y0 = (ctu / ctus_x) * 64;
  tvc_me_job j;
  j.ref_index = ref_index; j.ref_slot = ref_slot;
  j.x = x0 + cp.x; j.y = y0 + cp.y; j.w = cp.w; j.h = cp.h;
  if (j.x + j.w > pic_w || j.y + j.h > pic_h) j.w = 0;
  j.mode = TVC_ME_TZ; j.fen = cfg.fen; j.search_range = cfg.search_range;
  const int cu_x = x0 + cp.cu_x, cu_y = y0 + cp.cu_y;
  // xSetSearchRange (TEncSearch.cpp:4209-4225)
  int px = p.cx, py = p.cy;
  clip_mv(pic_w, pic_h, cu_x, cu_y, px, py);
  int lx = px - (cfg.search_range << 2), ty = py - (cfg.search_range << 2);
  int rx = px + (cfg.search_range << 2), by = py + (cfg.search_range << 2);
  clip_mv(pic_w, pic_h, cu_x, cu_y, lx, ty);
  clip_mv(pic_w, pic_h, cu_x, cu_y, rx, by);
  j.lx = lx >> 2; j.ty = ty >> 2; j.rx = rx >> 2; j.by = by >> 2;
  j.predx = p.cx; j.predy = p.cy;
  j.startx = px >> 2; j.starty = py >> 2;          // xTZSearch :4311-4312
  j.lambda_cost = cfg.lambda_cost;
  return j;
}

__global__ void k_me_frame_jobs(int pic_w, int pic_h, int num_ctus, int ctus_x, int num_refs, const int* __restrict__ ref_slots,
                                const tvc_me_center* __restrict__ pred, tvc_me_frame_cfg cfg, tvc_me_job* __restrict__ jobs)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int total = num_refs * num_ctus * TVC_ME_CENSUS;
  if (i >= total) return;
  int k = i % TVC_ME_CENSUS, rc = i / TVC_ME_CENSUS, ctu = rc % num_ctus, ref = rc / num_ctus;
  jobs[i] = census_job(pic_w, pic_h, ctus_x, ctu, k, ref, ref_slots[ref], pred[rc], cfg);
}

// one (CTU, reference) group with an explicit predictor (tvc_me_ctu)
__global__ void k_me_ctu_jobs(int pic_w, int pic_h, int ctus_x, int ctu, int ref_index, int ref_slot, tvc_me_center pred,
                              tvc_me_frame_cfg cfg, tvc_me_job* __restrict__ jobs)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= TVC_ME_CENSUS) return;
  jobs[k] = census_job(pic_w, pic_h, ctus_x, ctu, k, ref_index, ref_slot, pred, cfg);
}

__global__ void k_me_frame_frac_jobs(int n, const tvc_me_job* __restrict__ jobs, const tvc_me_result* __restrict__ res,
                                     int hadamard, tvc_frac_job* __restrict__ fj)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const tvc_me_job j = jobs[i];
  const tvc_me_result r = res[i];
  tvc_frac_job f;
  f.ref_slot = j.ref_slot; f.x = j.x; f.y = j.y; f.w = j.w; f.h = j.h;
  f.imvx = r.mvx; f.imvy = r.mvy; f.predx = j.predx; f.predy = j.predy;
  f.lambda_cost = j.lambda_cost; f.hadamard = hadamard;
  fj[i] = f;
}

// integer + fractional result of a job in 16 bytes (tvc_me_packed): what TEncSearch::xMotionEstimation takes from the two stages
__global__ void k_me_pack(int n, const tvc_me_result* __restrict__ ri, const tvc_frac_result* __restrict__ rf, tvc_me_packed* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const tvc_me_result a = ri[i];
  const tvc_frac_result b = rf[i];
  tvc_me_packed p;
  p.mvx = (int16_t)a.mvx; p.mvy = (int16_t)a.mvy;
  p.halfx = (int8_t)b.halfx; p.halfy = (int8_t)b.halfy; p.qtrx = (int8_t)b.qtrx; p.qtry = (int8_t)b.qtry;
  p.sad = a.n_sads ? a.sad : 0xFFFFFFFFu;
  p.cost = a.n_sads ? b.cost : 0xFFFFFFFFu;
  out[i] = p;
}

// ================================================================================ micro-benchmarks
__global__ void k_ub_vabsdiff4(uint32_t* out, int iters, uint32_t a0, uint32_t b0)
{
  uint32_t a = a0 + threadIdx.x, b = b0 * threadIdx.x;
  uint32_t acc[8] = {0, 1, 2, 3, 4, 5, 6, 7};
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = vsad4_acc(a, b ^ acc[(k + 1) & 7], acc[k]);
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ub_iadd3(uint32_t* out, int iters, uint32_t a0, uint32_t b0)
{
  uint32_t a = a0 + threadIdx.x, b = b0 * threadIdx.x;
  uint32_t acc[8] = {0, 1, 2, 3, 4, 5, 6, 7};
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) asm volatile("add.u32 %0, %0, %1;" : "+r"(acc[k]) : "r"(a));
#pragma unroll
    for (int k = 0; k < 8; k++) asm volatile("add.u32 %0, %0, %1;" : "+r"(acc[k]) : "r"(b));
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ub_imad(uint32_t* out, int iters, uint32_t a0, uint32_t b0)
{
  uint32_t a = a0 + threadIdx.x, b = b0 * threadIdx.x + 3;
  uint32_t acc[8] = {0, 1, 2, 3, 4, 5, 6, 7};
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(acc[k]) : "r"(a), "r"(b));
#pragma unroll
    for (int k = 0; k < 8; k++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(acc[k]) : "r"(b), "r"(a));
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ub_dp2a(uint32_t* out, int iters, uint32_t a0, uint32_t b0)
{
  int a = (int)(a0 + threadIdx.x), b = (int)(b0 * threadIdx.x + 3);
  int acc[8] = {0, 1, 2, 3, 4, 5, 6, 7};
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = __dp2a_lo(a, b, acc[k]);
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = __dp2a_hi(b, a, acc[k]);
  }
  int s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)s;
}
__global__ void k_ub_hbm_write(uint4* __restrict__ dst, size_t n16, uint32_t seed)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint4 v = make_uint4(seed, seed + 1, seed + 2, seed + 3);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = v;
}
// the SAD-table store pattern: a warp instruction writes 8 segments of 64 bytes that lie 16 KB apart (8 dx values x 4 quarters),
// a lane's next store is the adjacent 64 bytes (next block row); seg_bytes = 64 reproduces it, 512 is one contiguous run per warp
__global__ void k_ub_hbm_write_pat(uint4* __restrict__ dst, size_t n16, uint32_t seed, int seg16)
{
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const uint4 v = make_uint4(seed, seed + 1, seed + 2, seed + 3);
  // the buffer is cut into 1 KB records (64 x 16 B); a warp owns 32 / seg16 records at a time, spaced 16 records apart
  const int nseg = 32 / seg16, sl = lane / seg16, q = lane % seg16;
  const size_t recs = n16 / 64;
  for (size_t r0 = warp * 16 * nseg; r0 + 16 * nseg <= recs; r0 += nwarps * 16 * nseg)
    for (int a = 0; a < 16; a++)                       // 16 interleaved record groups, like the 16 dx residues of the table tasks
      for (int part = 0; part < 64 / seg16; part++) dst[(r0 + a + 16 * (size_t)sl) * 64 + part * seg16 + q] = v;
}
__global__ void k_ub_lds128(uint32_t* out, int iters)
{
  __shared__ uint4 buf[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_uint4(i, i + 1, i + 2, i + 3);
  __syncthreads();
  uint4 acc = make_uint4(0, 0, 0, 0);
  int idx = threadIdx.x;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      uint4 v = buf[(idx + k * 32) & 1023];
      acc.x += v.x; acc.y ^= v.y; acc.z += v.z; acc.w ^= v.w;
    }
    idx = (idx + acc.x) & 1023;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <int MAXW, int MAXH, int NT, int JPC>
static int launch_frac_class(tvc_ctx* c, int cur_slot, int n, const tvc_frac_job* jobs_dev, tvc_frac_result* out_dev, int span,
                             int stride, int first, const uint8_t* done = nullptr)
{
  if (n <= 0) return TVC_OK;
  constexpr size_t smem = sizeof(FracSmem<MAXW, MAXH>) * JPC;
  static bool attr_set = false;
  if (!attr_set) {
    TVC_CUDA(c, cudaFuncSetAttribute(k_me_frac<MAXW, MAXH, NT, JPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  k_me_frac<MAXW, MAXH, NT, JPC><<<(n + JPC - 1) / JPC, NT * JPC, smem, c->stream>>>(c->planes, cur_slot, n, jobs_dev, out_dev,
                                                                                      c->cfg.bit_depth, span, stride, first, done);
  TVC_LAUNCH_CHECK(c);
  return TVC_OK;
}

template <int MAXW, int MAXH, int NT, int JPC>
static int launch_frac_listed(tvc_ctx* c, int cur_slot, const int* list, const int* count, const tvc_frac_job* jobs_dev, tvc_frac_result* out_dev)
{
  constexpr size_t smem = sizeof(FracSmem<MAXW, MAXH>) * JPC;
  static int grid = 0;
  if (!grid) {
    TVC_CUDA(c, cudaFuncSetAttribute(k_me_frac_listed<MAXW, MAXH, NT, JPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    TVC_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_me_frac_listed<MAXW, MAXH, NT, JPC>, NT * JPC, smem));
    grid = kNumSM * (occ > 0 ? occ : 1);
  }
  k_me_frac_listed<MAXW, MAXH, NT, JPC><<<grid, NT * JPC, smem, c->stream>>>(c->planes, cur_slot, list, count, jobs_dev, out_dev, c->cfg.bit_depth);
  TVC_LAUNCH_CHECK(c);
  return TVC_OK;
}

// census == true: jobs are laid out [ref*ctu][593] in census order; one launch per CU depth so that the
// thread count and shared memory of a CTA fit the PU sizes of that depth
template <int CUW, int NT, int JPC, bool CHILD = false>
static int launch_frac_cu(tvc_ctx* c, int cur_slot, int groups, const tvc_frac_job* jobs_dev, tvc_frac_result* out_dev, uint8_t* done,
                          int cu_per_group, int np, int first)
{
  constexpr size_t smem = sizeof(FracCuSmem<CUW>) * JPC;
  static bool attr_set = false;
  if (!attr_set) {
    TVC_CUDA(c, cudaFuncSetAttribute(k_me_frac_cu<CUW, NT, JPC, CHILD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int ncu = groups * cu_per_group;
  k_me_frac_cu<CUW, NT, JPC, CHILD><<<(ncu + JPC - 1) / JPC, NT * JPC, smem, c->stream>>>(c->planes, cur_slot, ncu, jobs_dev, out_dev, done,
                                                                                          c->cfg.bit_depth, cu_per_group, np, first, me_fused_enabled(c) ? c->fr_stats : nullptr);
  TVC_LAUNCH_CHECK(c);
  return TVC_OK;
}

// census == true: jobs are laid out [ref*ctu][593] in census order.  First the CU-level kernels (one launch per CU depth) serve every CU
// whose PUs share their integer vector; then one per-PU launch per depth -- thread count and shared memory of a CTA fit the PU sizes of
// that depth -- serves the rest (TVC_FRAC_CU=0: everything per PU)
static int launch_frac(tvc_ctx* c, int cur_slot, int n, const tvc_frac_job* jobs_dev, tvc_frac_result* out_dev, bool census,
                       uint8_t* done_buf = nullptr)
{
  ProfScope ps(c, TVC_PH_ME_FRAC);
  if (!census) return launch_frac_class<64, 64, 256, 1>(c, cur_slot, n, jobs_dev, out_dev, n, 0, 0);
  const int groups = n / TVC_ME_CENSUS;
  int r;
  static int use_cu = -1;
  if (use_cu < 0) { const char* e = getenv("TVC_FRAC_CU"); use_cu = e ? atoi(e) : 1; }
  uint8_t* done = nullptr;
  if (use_cu && done_buf) {
    done = done_buf;                                     // a caller with its own flags (calls in flight on a side stream)
    TVC_CUDA(c, cudaMemsetAsync(done, 0, (size_t)n, c->stream));
  } else if (use_cu) {
    if ((size_t)n > c->frac_done_cap) {
      if (c->frac_done) cudaFree(c->frac_done);
      c->frac_done = nullptr; c->frac_done_cap = 0;
      TVC_CUDA(c, cudaMalloc(&c->frac_done, (size_t)n));
      c->frac_done_cap = (size_t)n;
    }
    done = (uint8_t*)c->frac_done;
    TVC_CUDA(c, cudaMemsetAsync(done, 0, (size_t)n, c->stream));
  }
  if (use_cu) {
    if ((r = launch_frac_cu<64, 256, 1>(c, cur_slot, groups, jobs_dev, out_dev, done, 1, 13, 0))) return r;
    if ((r = launch_frac_cu<32, 128, 1>(c, cur_slot, groups, jobs_dev, out_dev, done, 4, 13, 13))) return r;
    if ((r = launch_frac_cu<16, 32, 4, true>(c, cur_slot, groups, jobs_dev, out_dev, done, 16, 13, 65))) return r;      // + its 8x8 children
    if ((r = launch_frac_cu<8, 32, 4>(c, cur_slot, groups, jobs_dev, out_dev, done, 64, 5, 273))) return r;
  }
  static int use_lists = -1;
  if (use_lists < 0) { const char* e = getenv("TVC_FRAC_LISTS"); use_lists = e ? atoi(e) : 1; }
  if (done && use_lists && groups >= kNumSM) {
    // picture-level call: the per-PU kernels walk lists of the unserved jobs (a single group keeps the direct launch: few CTAs)
    if ((size_t)n + 4 > c->frac_list_cap) {
      if (c->frac_list) cudaFree(c->frac_list);
      c->frac_list = nullptr; c->frac_list_cap = 0;
      TVC_CUDA(c, cudaMalloc(&c->frac_list, ((size_t)n + 4) * sizeof(int)));
      c->frac_list_cap = (size_t)n + 4;
    }
    int* counts = (int*)c->frac_list;                    // four counters, then the lists (class c at groups * first(c))
    int* lists = counts + 4;
    TVC_CUDA(c, cudaMemsetAsync(counts, 0, 4 * sizeof(int), c->stream));
    k_frac_rest_lists<<<(n + 255) / 256, 256, 0, c->stream>>>(done, n, groups, lists, counts);
    TVC_LAUNCH_CHECK(c);
    if ((r = launch_frac_listed<64, 64, 256, 1>(c, cur_slot, lists, counts + 0, jobs_dev, out_dev))) return r;
    if ((r = launch_frac_listed<32, 32, 128, 1>(c, cur_slot, lists + (size_t)groups * 13, counts + 1, jobs_dev, out_dev))) return r;
    if ((r = launch_frac_listed<16, 16, 32, 4>(c, cur_slot, lists + (size_t)groups * 65, counts + 2, jobs_dev, out_dev))) return r;
    return launch_frac_listed<8, 8, 32, 4>(c, cur_slot, lists + (size_t)groups * 273, counts + 3, jobs_dev, out_dev);
  }
  if ((r = launch_frac_class<64, 64, 256, 1>(c, cur_slot, groups * 13, jobs_dev, out_dev, 13, TVC_ME_CENSUS, 0, done))) return r;
  if ((r = launch_frac_class<32, 32, 128, 1>(c, cur_slot, groups * 52, jobs_dev, out_dev, 52, TVC_ME_CENSUS, 13, done))) return r;
  if ((r = launch_frac_class<16, 16, 32, 4>(c, cur_slot, groups * 208, jobs_dev, out_dev, 208, TVC_ME_CENSUS, 65, done))) return r;
  return launch_frac_class<8, 8, 32, 4>(c, cur_slot, groups * 320, jobs_dev, out_dev, 320, TVC_ME_CENSUS, 273, done);
}

}  // namespace tvc

using namespace tvc;

// group kernel over `ngroups` census groups (jobs laid out [group][593]).  ref_index_fixed >= 0: every group searches
// maps.ref[ref_index_fixed] (tvc_me_ctu), else the group's own job.ref_index.
int tvc_launch_me_group(tvc_ctx* c, int cur_slot, int ngroups, const tvc_me_job* jobs_dev, tvc_me_result* out_dev, int num_refs,
                        const int* ref_slots, int ref_index_fixed, unsigned long long* stats)
{
  if (!c->pics[cur_slot].has_tmap || c->cfg.bit_depth != 8) return set_err(c, TVC_ERR_STATE, "group search: needs the 8-bit u8 planes and tensor maps");
  static int cu_stage = 1;
  static int minb = -1;          // tuning knob: resident CTAs per SM the kernel is compiled for (2: 128 registers, no spill; 3: 80)
  if (minb < 0) {
    const char* e = getenv("TVC_GROUP_MINB");
    minb = e ? atoi(e) : 2;
    const char* e2 = getenv("TVC_GROUP_CU");       // 0: every PU searches on its own from the start (no CU-level first search)
    cu_stage = e2 ? atoi(e2) : 1;
    TVC_CUDA(c, cudaFuncSetAttribute(k_me_group<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGrpSmem));
    TVC_CUDA(c, cudaFuncSetAttribute(k_me_group<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGrpSmem));
  }
  GroupMaps maps;
  memset(&maps, 0, sizeof(maps));
  const Pic& p = c->pics[cur_slot];
  maps.cur = p.tmap_cur80;
  for (int r = 0; r < num_refs && r < 8; r++) {
    const Pic& rp = c->pics[ref_slots[r]];
    if (!rp.has_tmap) return set_err(c, TVC_ERR_ARG, "group search: reference slot without tensor map");
    maps.ref[r] = rp.tmap_ref;
    maps.ref8[r] = rp.org8;
  }
  maps.stride8 = p.stride8;
  ProfScope ps(c, TVC_PH_ME_SEARCH);
  // few groups (tvc_me_ctu: one): spread each over several CTAs so that the call fills the machine
  const int split = ngroups >= 2 * kNumSM ? 1 : (ngroups >= kNumSM / 4 ? 4 : 24);
  static int order = -1;           // 0 launch order, 1 reversed, 2 references interleaved, 3 both, 4 (default): by the previous call's cost
  if (order < 0) { const char* e = getenv("TVC_GROUP_ORDER"); order = e ? atoi(e) : 4; }
  const int nr = num_refs > 0 ? num_refs : 1;
  // picture-level calls on the context's own stream keep a cost per group and launch the next call of the same shape longest first
  const int* order_list = nullptr;
  uint32_t* cost_out = nullptr;
  if (order == 4 && split == 1 && ngroups <= 4096 && ref_index_fixed < 0) {
    if (!c->grp_cost) {
      TVC_CUDA(c, cudaMalloc(&c->grp_cost, 4096 * sizeof(uint32_t)));
      TVC_CUDA(c, cudaMalloc(&c->grp_order, 4096 * sizeof(int)));
      c->grp_order_n = 0;
    }
    if (c->grp_order_n == ngroups) order_list = (const int*)c->grp_order;
    cost_out = (uint32_t*)c->grp_cost;
    TVC_CUDA(c, cudaMemsetAsync(cost_out, 0, (size_t)ngroups * sizeof(uint32_t), c->stream));
  }
  const int ord = order == 4 ? 2 : order;
  if (minb >= 3)
    k_me_group<3><<<ngroups * split, kGrpThreads, kGrpSmem, c->stream>>>(maps, c->planes, cur_slot, jobs_dev, out_dev, c->cfg.width, c->cfg.height,
                                                                         p.mx[0], p.my[0], ref_index_fixed, c->num_ctus_x, stats, split, cu_stage, ord, nr,
                                                                         order_list, cost_out);
  else
    k_me_group<2><<<ngroups * split, kGrpThreads, kGrpSmem, c->stream>>>(maps, c->planes, cur_slot, jobs_dev, out_dev, c->cfg.width, c->cfg.height,
                                                                         p.mx[0], p.my[0], ref_index_fixed, c->num_ctus_x, stats, split, cu_stage, ord, nr,
                                                                         order_list, cost_out);
  if (cost_out) {
    k_group_order<<<1, 1024, 0, c->stream>>>(cost_out, ngroups, (int*)c->grp_order);
    c->grp_order_n = ngroups;
  }
  TVC_LAUNCH_CHECK(c);
  return TVC_OK;
}

int tvc_launch_me_search_list(tvc_ctx* c, int cur_slot, const int* list_dev, const int* count_dev, const tvc_me_job* jobs_dev,
                              tvc_me_result* out_dev)
{
  k_me_search_list<<<kNumSM * 4, 128, 0, c->stream>>>(c->planes, cur_slot, list_dev, count_dev, jobs_dev, out_dev, c->num_ctus_x, c->bi);
  TVC_LAUNCH_CHECK(c);
  return TVC_OK;
}

namespace tvc {
bool me_fused_enabled(const tvc_ctx* c)
{
  int on = c->me_fused;
  if (on < 0) {
    static int env = -1;
    if (env < 0) { const char* e = getenv("TVC_ME_FUSED"); env = e ? atoi(e) : 1; }
    on = env;
  }
  return on != 0 && c->cfg.bit_depth == 8 && !c->pics.empty() && c->pics[0].has_tmap;
}
}  // namespace tvc

extern "C" {

size_t tvc_me_table_bytes(tvc_ctx* c, int num_refs)
{
  if (!c || num_refs <= 0) return 0;
  return (size_t)num_refs * c->num_ctus_x * c->num_ctus_y * kMeCtuElems * sizeof(uint16_t);
}

int tvc_me_set_fused(tvc_ctx* c, int on)
{
  if (!c) return TVC_ERR_ARG;
  c->me_fused = on < 0 ? -1 : (on ? 1 : 0);
  return TVC_OK;
}

int tvc_me_uses_tables(tvc_ctx* c) { return (c && !me_fused_enabled(c) && c->cfg.bit_depth == 8) ? 1 : 0; }

int tvc_me_reserve(tvc_ctx* c, int num_refs)
{
  if (!c || num_refs <= 0 || num_refs > 8) return set_err(c, TVC_ERR_ARG, "tvc_me_reserve: 1..8 references");
  const size_t need = tvc_me_table_bytes(c, num_refs);
  if (need <= c->me_table_bytes) return TVC_OK;
  if (c->me_tables) { cudaStreamSynchronize(c->stream); cudaFree(c->me_tables); }
  c->me_tables = nullptr; c->me_table_bytes = 0; c->me_num_refs = 0;
  if (cudaMalloc(&c->me_tables, need) != cudaSuccess) { cudaGetLastError(); return set_err(c, TVC_ERR_NOMEM, "tvc_me_reserve: cannot allocate %zu bytes of SAD tables", need); }
  c->me_table_bytes = need;
  return TVC_OK;
}

// side streams + events of the pipelined frame pre-pass (created on first use)
static int ensure_pipe(tvc_ctx* c)
{
  if (c->pipe[0]) return TVC_OK;
  int least = 0, greatest = 0;
  TVC_CUDA(c, cudaDeviceGetStreamPriorityRange(&least, &greatest));
  const int mid = greatest + 1 <= least ? greatest + 1 : greatest;
  TVC_CUDA(c, cudaStreamCreateWithPriority(&c->pipe[0], cudaStreamNonBlocking, mid));
  TVC_CUDA(c, cudaStreamCreateWithPriority(&c->pipe[1], cudaStreamNonBlocking, greatest));
  c->pipe_ev.resize(1 + 2 * kMaxPipeChunks * 8);
  for (auto& e : c->pipe_ev) TVC_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return TVC_OK;
}

// allocation, argument checks and the upload of the clamped table centres: everything of the pre-pass but the kernel
static int prepass_prepare(tvc_ctx* c, int cur_slot, int num_refs, const int* ref_slots, const tvc_me_center* centers)
{
  if (!c || !valid_slot(c, cur_slot) || num_refs <= 0 || num_refs > 8 || !ref_slots)
    return set_err(c, TVC_ERR_ARG, "tvc_me_prepass: bad argument (1..8 references)");
  if (c->cfg.bit_depth != 8) return set_err(c, TVC_ERR_ARG, "tvc_me_prepass: the SAD-table path is the 8-bit u8 SIMD path");
  for (int r = 0; r < num_refs; r++)
    if (!valid_slot(c, ref_slots[r]) || !c->pics[ref_slots[r]].has_tmap) return set_err(c, TVC_ERR_ARG, "tvc_me_prepass: bad reference slot or no tensor map");
  if (!c->pics[cur_slot].has_tmap) return set_err(c, TVC_ERR_STATE, "tvc_me_prepass: tensor maps unavailable (cuTensorMapEncodeTiled)");
  const int nctu = c->num_ctus_x * c->num_ctus_y;
  size_t need = tvc_me_table_bytes(c, num_refs);
  if (need > c->me_table_bytes) {
    if (c->me_tables) cudaFree(c->me_tables);
    c->me_tables = nullptr; c->me_table_bytes = 0;
    if (cudaMalloc(&c->me_tables, need) != cudaSuccess) { cudaGetLastError(); return set_err(c, TVC_ERR_NOMEM, "tvc_me_prepass: cannot allocate %zu bytes of SAD tables", need); }
    c->me_table_bytes = need;
  }
  if (!c->me_centers) TVC_CUDA(c, cudaMalloc(&c->me_centers, sizeof(tvc_me_center) * 8 * nctu));
  // clamp the centres so that the 192x192 window stays inside the padded plane
  const Pic& p = c->pics[cur_slot];
  int r;
  if ((r = stage_acquire(c, c->me_stage, c->me_ev, sizeof(tvc_me_center) * (size_t)num_refs * nctu))) return r;
  tvc_me_center* hc = (tvc_me_center*)c->me_stage.host;
  for (int rf = 0; rf < num_refs; rf++)
    for (int k = 0; k < nctu; k++) {
      int x0 = (k % c->num_ctus_x) * 64, y0 = (k / c->num_ctus_x) * 64;
      tvc_me_center ce = centers ? centers[(size_t)rf * nctu + k] : tvc_me_center{0, 0};
      int lo_x = -p.mx[0] - x0 + kMeR, hi_x = p.w[0] + p.mx[0] - x0 - 64 - kMeR;
      int lo_y = -p.my[0] - y0 + kMeR, hi_y = p.h[0] + p.my[0] - y0 - 64 - kMeR;
      if (hi_x < lo_x) hi_x = lo_x;
      if (hi_y < lo_y) hi_y = lo_y;
      ce.cx = ce.cx < lo_x ? lo_x : (ce.cx > hi_x ? hi_x : ce.cx);
      ce.cy = ce.cy < lo_y ? lo_y : (ce.cy > hi_y ? hi_y : ce.cy);
      hc[(size_t)rf * nctu + k] = ce;
    }
  TVC_CUDA(c, cudaMemcpyAsync(c->me_centers, hc, sizeof(tvc_me_center) * (size_t)num_refs * nctu, cudaMemcpyHostToDevice, c->stream));
  TVC_CUDA(c, cudaEventRecord(c->me_ev, c->stream));
  c->me_num_refs = num_refs;
  c->me_cur_slot = cur_slot;
  for (int rf = 0; rf < num_refs; rf++) c->me_ref_slots[rf] = ref_slots[rf];
  return TVC_OK;
}

// SAD tables of CTUs [ctu0, ctu0 + nct) x references [ref0, ref0 + nrf) of the prepared pre-pass, on c->stream
static int launch_tables(tvc_ctx* c, int ctu0, int nct, int ref0, int nrf)
{
  const int nctu = c->num_ctus_x * c->num_ctus_y;
  const Pic& p = c->pics[c->me_cur_slot];
  MeMaps maps;
  memset(&maps, 0, sizeof(maps));
  maps.cur = p.tmap_cur;
  for (int rf = 0; rf < c->me_num_refs; rf++) maps.ref[rf] = c->pics[c->me_ref_slots[rf]].tmap_ref;
  // tuning knob (measured on B200, ms per 1080p picture x 4 references): 81 = row-split lane mapping, 8-row window, 2 CTAs/SM:
  // 5.4 (default); 8 = column mapping (64-byte store pieces): 7.0; 4 / 43 = 4-row window at 4 / 3 CTAs/SM: 7.4 / 7.0; 41: 6.6
  static int variant = -1, pad = 0;
  if (variant < 0) {
    const char* ev = getenv("TVC_TABLE_DYB");
    variant = ev ? atoi(ev) : 81;
    // TVC_TABLE_SMEM_PAD=<KB>: unused dynamic shared memory that caps the resident table CTAs per SM (72 -> one), leaving registers
    // for the search / fractional-search CTAs of the pipelined form (TVC_ME_PIPE)
    const char* ep = getenv("TVC_TABLE_SMEM_PAD");
    pad = ep ? atoi(ep) * 1024 : 0;
    TVC_CUDA(c, cudaFuncSetAttribute(k_me_sad_tables<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTables + pad));
    TVC_CUDA(c, cudaFuncSetAttribute(k_me_sad_tables<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTables + pad));
    TVC_CUDA(c, cudaFuncSetAttribute(k_me_sad_tables<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTables + pad));
    TVC_CUDA(c, cudaFuncSetAttribute(k_me_sad_tables<8, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTables + pad));
    TVC_CUDA(c, cudaFuncSetAttribute(k_me_sad_tables<4, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTables + pad));
  }
  const size_t smem_tables = kSmemTables + pad;
  dim3 grd(nct, nrf);
  ProfScope ps(c, TVC_PH_ME_TABLES);
  if (variant == 4) k_me_sad_tables<4, 4><<<grd, 256, smem_tables, c->stream>>>(maps, nctu, c->num_ctus_x, p.mx[0], p.my[0], c->me_centers, c->me_tables, ctu0, ref0);
  else if (variant == 43) k_me_sad_tables<4, 3><<<grd, 256, smem_tables, c->stream>>>(maps, nctu, c->num_ctus_x, p.mx[0], p.my[0], c->me_centers, c->me_tables, ctu0, ref0);
  else if (variant == 81) k_me_sad_tables<8, 2, true><<<grd, 256, smem_tables, c->stream>>>(maps, nctu, c->num_ctus_x, p.mx[0], p.my[0], c->me_centers, c->me_tables, ctu0, ref0);
  else if (variant == 41) k_me_sad_tables<4, 3, true><<<grd, 256, smem_tables, c->stream>>>(maps, nctu, c->num_ctus_x, p.mx[0], p.my[0], c->me_centers, c->me_tables, ctu0, ref0);
  else k_me_sad_tables<8, 2><<<grd, 256, smem_tables, c->stream>>>(maps, nctu, c->num_ctus_x, p.mx[0], p.my[0], c->me_centers, c->me_tables, ctu0, ref0);
  TVC_LAUNCH_CHECK(c);
  return TVC_OK;
}

int tvc_me_prepass(tvc_ctx* c, int cur_slot, int num_refs, const int* ref_slots, const tvc_me_center* centers)
{
  int r = prepass_prepare(c, cur_slot, num_refs, ref_slots, centers);
  if (r) return r;
  return launch_tables(c, 0, c->num_ctus_x * c->num_ctus_y, 0, num_refs);
}

int tvc_me_tables_dev(tvc_ctx* c, void** tables, tvc_me_center** centers_dev)
{
  if (!c || !c->me_tables || c->me_num_refs == 0) return set_err(c, TVC_ERR_STATE, "tvc_me_tables_dev: no pre-pass has run");
  if (tables) *tables = c->me_tables;
  if (centers_dev) *centers_dev = c->me_centers;
  return TVC_OK;
}

int tvc_me_table_lookup(tvc_ctx* c, int ref_index, int pu_x, int pu_y, int pu_w, int pu_h, int fen, int n,
                        const int16_t* cand_xy, uint32_t* out)
{
  if (!c || !c->me_tables || c->me_num_refs == 0) return set_err(c, TVC_ERR_STATE, "tvc_me_table_lookup: no pre-pass has run");
  if (ref_index < 0 || ref_index >= c->me_num_refs || n < 0 || (n && (!cand_xy || !out)) || (pu_w & 3) || (pu_h & 3) ||
      pu_w <= 0 || pu_h <= 0 || pu_x < 0 || pu_y < 0 || (pu_x & 3) || (pu_y & 3) || (pu_x & 63) + pu_w > 64 || (pu_y & 63) + pu_h > 64 ||
      pu_x >= c->num_ctus_x * 64 || pu_y >= c->num_ctus_y * 64)
    return set_err(c, TVC_ERR_ARG, "tvc_me_table_lookup: bad argument");
  if (n == 0) return TVC_OK;
  int r;
  if ((r = ensure_scratch(c, c->in, (size_t)n * 4))) return r;
  if ((r = ensure_scratch(c, c->out, (size_t)n * 4))) return r;
  memcpy(c->in.host, cand_xy, (size_t)n * 4);
  TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, c->in.host, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  k_me_table_lookup<<<(n + 3) / 4, 128, 0, c->stream>>>(c->me_tables, c->me_centers, c->num_ctus_x * c->num_ctus_y, c->num_ctus_x,
                                                         ref_index, pu_x, pu_y, pu_w, pu_h, fen, n, (const int16_t*)c->in.dev,
                                                         (uint32_t*)c->out.dev, c->bi);
  TVC_LAUNCH_CHECK(c);
  TVC_CUDA(c, cudaMemcpyAsync(c->out.host, c->out.dev, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(out, c->out.host, (size_t)n * 4);
  return TVC_OK;
}

static int launch_search(tvc_ctx* c, int cur_slot, int use_tables, int n, const tvc_me_job* jobs_dev, tvc_me_result* out_dev,
                         const RasterBest* rast, const SweepState* sweep, unsigned long long* stats)
{
  if (!c || !valid_slot(c, cur_slot) || n < 0 || (n && (!jobs_dev || !out_dev))) return set_err(c, TVC_ERR_ARG, "tvc_me_search_batch_dev: bad argument");
  if (use_tables && (!c->me_tables || c->me_num_refs == 0 || c->me_cur_slot != cur_slot))
    return set_err(c, TVC_ERR_STATE, "tvc_me_search_batch: tables requested but no pre-pass for this picture");
  if (n == 0) return TVC_OK;
  ProfScope ps(c, TVC_PH_ME_SEARCH);
  static int variant = -1;       // tuning knob: resident blocks per SM the kernel is compiled for
  if (variant < 0) { const char* e = getenv("TVC_SEARCH_MINB"); variant = e ? atoi(e) : 4; }
  const uint16_t* tb = use_tables ? c->me_tables : nullptr;
  const int nctu = c->num_ctus_x * c->num_ctus_y;
  if (variant >= 8) k_me_search<8><<<(n + 3) / 4, 128, 0, c->stream>>>(c->planes, cur_slot, n, jobs_dev, out_dev, tb, c->me_centers, nctu, c->num_ctus_x, c->bi, rast, sweep, stats);
  else if (variant == 5) k_me_search<5><<<(n + 3) / 4, 128, 0, c->stream>>>(c->planes, cur_slot, n, jobs_dev, out_dev, tb, c->me_centers, nctu, c->num_ctus_x, c->bi, rast, sweep, stats);
  else if (variant >= 6) k_me_search<6><<<(n + 3) / 4, 128, 0, c->stream>>>(c->planes, cur_slot, n, jobs_dev, out_dev, tb, c->me_centers, nctu, c->num_ctus_x, c->bi, rast, sweep, stats);
  else if (variant >= 4) k_me_search<4><<<(n + 3) / 4, 128, 0, c->stream>>>(c->planes, cur_slot, n, jobs_dev, out_dev, tb, c->me_centers, nctu, c->num_ctus_x, c->bi, rast, sweep, stats);
  else k_me_search<1><<<(n + 3) / 4, 128, 0, c->stream>>>(c->planes, cur_slot, n, jobs_dev, out_dev, tb, c->me_centers, nctu, c->num_ctus_x, c->bi, rast, sweep, stats);
  TVC_LAUNCH_CHECK(c);
  return TVC_OK;
}

int tvc_me_search_batch_dev(tvc_ctx* c, int cur_slot, int use_tables, int n, const tvc_me_job* jobs_dev, tvc_me_result* out_dev)
{
  return launch_search(c, cur_slot, use_tables, n, jobs_dev, out_dev, nullptr, nullptr, nullptr);
}

int tvc_me_search_batch(tvc_ctx* c, int cur_slot, int use_tables, int n, const tvc_me_job* jobs, tvc_me_result* out)
{
  if (!c || !valid_slot(c, cur_slot) || n < 0 || (n && (!jobs || !out))) return set_err(c, TVC_ERR_ARG, "tvc_me_search_batch: bad argument");
  if (n == 0) return TVC_OK;
  const Pic& p = c->pics[cur_slot];
  for (int i = 0; i < n; i++) {
    const tvc_me_job& j = jobs[i];
    bool ok = valid_slot(c, j.ref_slot) && j.w > 0 && j.h > 0 && j.w <= 64 && j.h <= 64 && !(j.w & 3) && !(j.h & 3) && !(j.x & 3) &&
              !(j.y & 3) && j.x >= 0 && j.y >= 0 && (j.x & 63) + j.w <= 64 && (j.y & 63) + j.h <= 64 &&
              j.x < c->num_ctus_x * 64 && j.y < c->num_ctus_y * 64 && j.lx <= j.rx && j.ty <= j.by &&
              (j.mode == TVC_ME_FULL || j.mode == TVC_ME_TZ) && j.search_range >= 1 && j.search_range <= TVC_ME_RANGE &&
              (!use_tables || (j.ref_index >= 0 && j.ref_index < c->me_num_refs && c->me_ref_slots[j.ref_index] == j.ref_slot));
    // every candidate the search may touch must read inside the padded plane
    if (ok) {
      int minx = j.lx < 0 ? j.lx : 0, maxx = j.rx > 0 ? j.rx : 0, miny = j.ty < 0 ? j.ty : 0, maxy = j.by > 0 ? j.by : 0;
      if (j.mode == TVC_ME_TZ) { minx = minx < j.startx ? minx : j.startx; maxx = maxx > j.startx ? maxx : j.startx;
                                 miny = miny < j.starty ? miny : j.starty; maxy = maxy > j.starty ? maxy : j.starty; }
      ok = j.x + minx >= -p.mx[0] && j.y + miny >= -p.my[0] && j.x + j.w + maxx <= p.w[0] + p.mx[0] && j.y + j.h + maxy <= p.h[0] + p.my[0];
    }
    if (!ok) return set_err(c, TVC_ERR_ARG, "tvc_me_search_batch: job %d invalid", i);
  }
  int r;
  if ((r = ensure_scratch(c, c->in, (size_t)n * sizeof(tvc_me_job)))) return r;
  if ((r = ensure_scratch(c, c->out, (size_t)n * sizeof(tvc_me_result)))) return r;
  memcpy(c->in.host, jobs, (size_t)n * sizeof(tvc_me_job));
  TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, c->in.host, (size_t)n * sizeof(tvc_me_job), cudaMemcpyHostToDevice, c->stream));
  if ((r = tvc_me_search_batch_dev(c, cur_slot, use_tables, n, (const tvc_me_job*)c->in.dev, (tvc_me_result*)c->out.dev))) return r;
  TVC_CUDA(c, cudaMemcpyAsync(c->out.host, c->out.dev, (size_t)n * sizeof(tvc_me_result), cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(out, c->out.host, (size_t)n * sizeof(tvc_me_result));
  return TVC_OK;
}

int tvc_me_frac_batch_dev(tvc_ctx* c, int cur_slot, int n, const tvc_frac_job* jobs_dev, tvc_frac_result* out_dev)
{
  if (!c || !valid_slot(c, cur_slot) || n < 0 || (n && (!jobs_dev || !out_dev))) return set_err(c, TVC_ERR_ARG, "tvc_me_frac_batch_dev: bad argument");
  if (n == 0) return TVC_OK;
  return launch_frac(c, cur_slot, n, jobs_dev, out_dev, false);
}

int tvc_me_frac_batch(tvc_ctx* c, int cur_slot, int n, const tvc_frac_job* jobs, tvc_frac_result* out)
{
  if (!c || !valid_slot(c, cur_slot) || n < 0 || (n && (!jobs || !out))) return set_err(c, TVC_ERR_ARG, "tvc_me_frac_batch: bad argument");
  if (n == 0) return TVC_OK;
  const Pic& p = c->pics[cur_slot];
  for (int i = 0; i < n; i++) {
    const tvc_frac_job& j = jobs[i];
    bool ok = valid_slot(c, j.ref_slot) && j.w > 0 && j.h > 0 && j.w <= 64 && j.h <= 64 && !(j.w & 3) && !(j.h & 3) && j.x >= 0 && j.y >= 0 &&
              j.x + j.w <= p.w[0] + p.mx[0] && j.y + j.h <= p.h[0] + p.my[0];
    if (ok) {
      int ix = j.x + j.imvx, iy = j.y + j.imvy;   // 8-tap reach: -4 .. +4 around the block (+-1 for the half-pel shift)
      ok = ix - 5 >= -p.mx[0] && iy - 5 >= -p.my[0] && ix + j.w + 5 <= p.w[0] + p.mx[0] && iy + j.h + 5 <= p.h[0] + p.my[0];
    }
    if (!ok) return set_err(c, TVC_ERR_ARG, "tvc_me_frac_batch: job %d invalid or reaches outside the padded picture", i);
  }
  int r;
  if ((r = ensure_scratch(c, c->in, (size_t)n * sizeof(tvc_frac_job)))) return r;
  if ((r = ensure_scratch(c, c->out, (size_t)n * sizeof(tvc_frac_result)))) return r;
  memcpy(c->in.host, jobs, (size_t)n * sizeof(tvc_frac_job));
  TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, c->in.host, (size_t)n * sizeof(tvc_frac_job), cudaMemcpyHostToDevice, c->stream));
  if ((r = tvc_me_frac_batch_dev(c, cur_slot, n, (const tvc_frac_job*)c->in.dev, (tvc_frac_result*)c->out.dev))) return r;
  TVC_CUDA(c, cudaMemcpyAsync(c->out.host, c->out.dev, (size_t)n * sizeof(tvc_frac_result), cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(out, c->out.host, (size_t)n * sizeof(tvc_frac_result));
  return TVC_OK;
}


int tvc_me_bipred(tvc_ctx* c, int target_slot, const int16_t* target, int target_stride, const tvc_me_job* job, int hadamard,
                  tvc_me_result* int_out, tvc_frac_result* frac_out)
{
  if (!c || !valid_slot(c, target_slot) || !target || !job || !int_out || !frac_out || target_stride < job->w)
    return set_err(c, TVC_ERR_ARG, "tvc_me_bipred: bad argument");
  const tvc_me_job& j = *job;
  const Pic& p = c->pics[target_slot];
  bool ok = valid_slot(c, j.ref_slot) && j.ref_slot != target_slot && j.w > 0 && j.h > 0 && j.w <= 64 && j.h <= 64 && !(j.w & 3) && !(j.h & 3) &&
            j.x >= 0 && j.y >= 0 && j.x + j.w <= p.w[0] + p.mx[0] && j.y + j.h <= p.h[0] + p.my[0] && j.lx <= j.rx && j.ty <= j.by &&
            j.rx - j.lx <= 2 * TVC_ME_RANGE && j.by - j.ty <= 2 * TVC_ME_RANGE && j.mode == TVC_ME_FULL;
  // every candidate, with the 8-tap reach of the fractional stage around it, must read inside the padded reference plane
  if (ok) ok = j.x + j.lx - 5 >= -p.mx[0] && j.y + j.ty - 5 >= -p.my[0] && j.x + j.w + j.rx + 5 <= p.w[0] + p.mx[0] && j.y + j.h + j.by + 5 <= p.h[0] + p.my[0];
  if (!ok) return set_err(c, TVC_ERR_ARG, "tvc_me_bipred: job invalid or reaches outside the padded picture");
  if (!c->bi_buf) {
    TVC_CUDA(c, cudaMalloc(&c->bi_buf, sizeof(tvc_me_job) + sizeof(tvc_me_result) + sizeof(tvc_frac_job) + sizeof(tvc_frac_result) + 64));
    TVC_CUDA(c, cudaHostAlloc(&c->bi_host, 64 * 64 * sizeof(int16_t) + sizeof(tvc_me_job) + sizeof(tvc_me_result) + sizeof(tvc_frac_result) + 64, cudaHostAllocDefault));
  }
  // device: [job][int result][frac result][frac job]; host (pinned): [target block w x h][job][int result][frac result]
  tvc_me_job* d_job = (tvc_me_job*)c->bi_buf;
  tvc_me_result* d_int = (tvc_me_result*)(d_job + 1);
  tvc_frac_result* d_frac = (tvc_frac_result*)(d_int + 1);
  tvc_frac_job* d_fjob = (tvc_frac_job*)(d_frac + 1);
  int16_t* h_blk = (int16_t*)c->bi_host;
  tvc_me_job* h_job = (tvc_me_job*)(h_blk + 64 * 64);
  tvc_me_result* h_int = (tvc_me_result*)(h_job + 1);
  for (int r = 0; r < j.h; r++) memcpy(h_blk + (size_t)r * j.w, target + (ptrdiff_t)r * target_stride, sizeof(int16_t) * j.w);
  *h_job = j;
  TVC_CUDA(c, cudaMemcpy2DAsync(p.org[0] + (ptrdiff_t)j.y * p.stride[0] + j.x, sizeof(int16_t) * p.stride[0], h_blk, sizeof(int16_t) * j.w,
                                sizeof(int16_t) * j.w, j.h, cudaMemcpyHostToDevice, c->stream));
  TVC_CUDA(c, cudaMemcpyAsync(d_job, h_job, sizeof(tvc_me_job), cudaMemcpyHostToDevice, c->stream));
  int r;
  if ((r = launch_search(c, target_slot, 0, 1, d_job, d_int, nullptr, nullptr, nullptr))) return r;
  k_me_frame_frac_jobs<<<1, 32, 0, c->stream>>>(1, d_job, d_int, hadamard, d_fjob);
  TVC_LAUNCH_CHECK(c);
  if ((r = launch_frac(c, target_slot, 1, d_fjob, d_frac, false))) return r;
  TVC_CUDA(c, cudaMemcpyAsync(h_int, d_int, sizeof(tvc_me_result) + sizeof(tvc_frac_result), cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  *int_out = *h_int;
  *frac_out = *(const tvc_frac_result*)(h_int + 1);
  return TVC_OK;
}

int tvc_me_census(tvc_census_pu* out)
{
  if (!out) return TVC_ERR_ARG;
  build_census(out);
  return TVC_OK;
}

static int ensure_census(tvc_ctx* c)
{
  static int census_dev = -1;
  if (census_dev != c->cfg.device) {
    tvc_census_pu h[TVC_ME_CENSUS];
    build_census(h);
    TVC_CUDA(c, cudaMemcpyToSymbol(c_census, h, sizeof(h)));
    census_dev = c->cfg.device;
  }
  return TVC_OK;
}

int tvc_me_frame_dev(tvc_ctx* c, int cur_slot, int num_refs, const int* ref_slots, const tvc_me_center* pred_qpel,
                     const tvc_me_frame_cfg* cfg, tvc_me_result** int_dev, tvc_frac_result** frac_dev)
{
  if (!c || !valid_slot(c, cur_slot) || num_refs <= 0 || num_refs > 8 || !ref_slots || !cfg || cfg->search_range < 1 ||
      cfg->search_range > TVC_ME_RANGE)
    return set_err(c, TVC_ERR_ARG, "tvc_me_frame: bad argument");
  for (int r = 0; r < num_refs; r++)
    if (!valid_slot(c, ref_slots[r])) return set_err(c, TVC_ERR_ARG, "tvc_me_frame: bad reference slot");
  const int nctu = c->num_ctus_x * c->num_ctus_y;
  const size_t n = (size_t)num_refs * nctu * TVC_ME_CENSUS;
  int r0;
  if ((r0 = ensure_census(c))) return r0;
  if (n > c->fr_cap) {
    if (c->fr_jobs) cudaFree(c->fr_jobs);
    if (c->fr_int) cudaFree(c->fr_int);
    if (c->fr_fjobs) cudaFree(c->fr_fjobs);
    if (c->fr_frac) cudaFree(c->fr_frac);
    if (c->fr_rast) cudaFree(c->fr_rast);
    if (c->fr_sweep) cudaFree(c->fr_sweep);
    c->fr_rast = nullptr; c->fr_sweep = nullptr;
    c->fr_jobs = nullptr; c->fr_int = nullptr; c->fr_fjobs = nullptr; c->fr_frac = nullptr; c->fr_cap = 0;
    TVC_CUDA(c, cudaMalloc(&c->fr_jobs, n * sizeof(tvc_me_job)));
    TVC_CUDA(c, cudaMalloc(&c->fr_int, n * sizeof(tvc_me_result)));
    TVC_CUDA(c, cudaMalloc(&c->fr_fjobs, n * sizeof(tvc_frac_job)));
    TVC_CUDA(c, cudaMalloc(&c->fr_frac, n * sizeof(tvc_frac_result)));
    TVC_CUDA(c, cudaMalloc(&c->fr_rast, n * 8));
    TVC_CUDA(c, cudaMalloc(&c->fr_sweep, n * sizeof(SweepState)));
    c->fr_cap = n;
  }
  // predictors (quarter pels) and table centres (CTU-level clipMv, integer pels) on the host: tiny
  const size_t np = (size_t)num_refs * nctu;
  int r;
  if ((r = stage_acquire(c, c->fr_stage, c->fr_ev, np * sizeof(tvc_me_center) + 8 * sizeof(int)))) return r;
  std::vector<tvc_me_center> centers(np);
  tvc_me_center* hp = (tvc_me_center*)c->fr_stage.host;
  int* hslots = (int*)((char*)c->fr_stage.host + np * sizeof(tvc_me_center));
  const int pw = c->cfg.width, ph = c->cfg.height;
  for (size_t i = 0; i < np; i++) {
    tvc_me_center p = pred_qpel ? pred_qpel[i] : tvc_me_center{0, 0};
    hp[i] = p;
    int ctu = (int)(i % nctu), x0 = (ctu % c->num_ctus_x) * 64, y0 = (ctu / c->num_ctus_x) * 64;
    int hmax = (pw + 8 - x0 - 1) * 4, hmin = (-64 - 8 - x0 + 1) * 4, vmax = (ph + 8 - y0 - 1) * 4, vmin = (-64 - 8 - y0 + 1) * 4;
    int x = p.cx < hmin ? hmin : (p.cx > hmax ? hmax : p.cx), y = p.cy < vmin ? vmin : (p.cy > vmax ? vmax : p.cy);
    centers[i].cx = x >> 2; centers[i].cy = y >> 2;
  }
  for (int k = 0; k < num_refs; k++) hslots[k] = ref_slots[k];
  // Pipelined form (TVC_ME_PIPE=<chunks of CTU rows per reference>, off by default): the SAD tables of chunk k+1 (c->stream) run beside
  // the raster + search of chunk k (pipe[0]) and the fractional search of chunk k-1 (pipe[1]; side streams at higher priority).
  // Measured on B200 (1080p, 4 references): serial 19.8 ms, 1 / 2 / 4 chunks 20.4 / 20.9 / 21.8 ms with identical results -- the
  // table kernel needs ~70 % of the integer issue slots to keep HBM saturated and fills the register file (2 x 256 x 128), so a
  // co-resident CTA only displaces table work and the chunk tails add up.  Kept as a knob; per-phase profiling needs the serial form.
  static int pipe_chunks = -1;
  if (pipe_chunks < 0) { const char* e = getenv("TVC_ME_PIPE"); pipe_chunks = e ? atoi(e) : 0; if (pipe_chunks > kMaxPipeChunks) pipe_chunks = kMaxPipeChunks; }
  // use_tables selects the fast integer stage.  Default: the group kernel (tvc_me_group.cu: one CTA per (CTU, reference), SADs on
  // demand from the staged window, nothing written to HBM but the results); TVC_ME_FUSED=0: the round-1 form (full SAD tables
  // in HBM, shared raster / first-sweep stage, per-PU search reading the tables).  10-bit content has no u8 planes: per-PU search.
  const bool fused = cfg->use_tables && me_fused_enabled(c);
  const bool legacy_tables = cfg->use_tables && !fused && c->cfg.bit_depth == 8;
  const bool piped = legacy_tables && pipe_chunks > 0 && !c->prof_on;
  c->fr_piped_last = piped;
  if (legacy_tables && (r = prepass_prepare(c, cur_slot, num_refs, ref_slots, centers.data()))) return r;
  TVC_CUDA(c, cudaMemcpyAsync(c->fr_stage.dev, c->fr_stage.host, np * sizeof(tvc_me_center) + num_refs * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  TVC_CUDA(c, cudaEventRecord(c->fr_ev, c->stream));
  const tvc_me_center* dpred = (const tvc_me_center*)c->fr_stage.dev;
  const int* dslots = (const int*)((char*)c->fr_stage.dev + np * sizeof(tvc_me_center));
  {
    ProfScope ps(c, TVC_PH_OTHER);
    k_me_frame_jobs<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(pw, ph, nctu, c->num_ctus_x, num_refs, dslots, dpred, *cfg, c->fr_jobs);
    TVC_LAUNCH_CHECK(c);
  }
  if (!c->fr_stats) TVC_CUDA(c, cudaMalloc(&c->fr_stats, 3 * sizeof(unsigned long long)));
  TVC_CUDA(c, cudaMemsetAsync(c->fr_stats, 0, 3 * sizeof(unsigned long long), c->stream));
  static int use_rast = -1, use_sweep = -1;      // tuning knobs (TVC_ME_RASTER=0 / TVC_ME_SWEEP=0: every PU on its own)
  if (use_rast < 0) { const char* e = getenv("TVC_ME_RASTER"); use_rast = e ? atoi(e) : 1; }
  if (use_sweep < 0) { const char* e = getenv("TVC_ME_SWEEP"); use_sweep = e ? atoi(e) : 1; }
  // raster + first sweep + TZ search of the jobs of CTUs [ctu0, ctu0 + nct) x references [ref0, ref0 + nrf) on c->stream
  auto search_part = [&](int ctu0, int nct, int ref0) -> int {
    const RasterBest* rast = nullptr;
    const SweepState* sweep = nullptr;
    if (cfg->use_tables && use_rast) {
      dim3 grd(nct, 1);
      if (use_sweep) {
        k_me_raster<true><<<grd, kRastThreads, 0, c->stream>>>(c->fr_jobs, c->me_tables, c->me_centers, nctu, 5, c->bi, (RasterBest*)c->fr_rast,
                                                               (SweepState*)c->fr_sweep, c->fr_stats, ctu0, ref0);
        TVC_LAUNCH_CHECK(c);
      }
      k_me_raster<false><<<grd, kRastThreads, 0, c->stream>>>(c->fr_jobs, c->me_tables, c->me_centers, nctu, 5, c->bi, (RasterBest*)c->fr_rast,
                                                              nullptr, c->fr_stats, ctu0, ref0);
      TVC_LAUNCH_CHECK(c);
      rast = (const RasterBest*)c->fr_rast;
      if (use_sweep) sweep = (const SweepState*)c->fr_sweep;
    }
    const size_t off = ((size_t)ref0 * nctu + ctu0) * TVC_ME_CENSUS;      // jobs of one reference's CTU range are contiguous
    return launch_search(c, cur_slot, cfg->use_tables, nct * TVC_ME_CENSUS, c->fr_jobs + off, c->fr_int + off,
                         rast ? rast + off : nullptr, sweep ? sweep + off : nullptr, c->fr_stats);
  };
  auto frac_part = [&](size_t off, size_t cnt) -> int {
    {
      ProfScope ps(c, TVC_PH_OTHER);
      k_me_frame_frac_jobs<<<(unsigned)((cnt + 255) / 256), 256, 0, c->stream>>>((int)cnt, c->fr_jobs + off, c->fr_int + off, cfg->hadamard, c->fr_fjobs + off);
      TVC_LAUNCH_CHECK(c);
    }
    return launch_frac(c, cur_slot, (int)cnt, c->fr_fjobs + off, c->fr_frac + off, true);
  };
  if (fused) {
    if ((r = tvc_launch_me_group(c, cur_slot, num_refs * nctu, c->fr_jobs, c->fr_int, num_refs, ref_slots, -1, c->fr_stats))) return r;
    if (c->fr_int_ready) TVC_CUDA(c, cudaEventRecord(c->fr_int_ready, c->stream));
    if (cfg->do_frac && (r = frac_part(0, n))) return r;
  } else if (!piped) {
    if (legacy_tables && (r = launch_tables(c, 0, nctu, 0, num_refs))) return r;
    if (legacy_tables && use_rast) {
      ProfScope ps(c, TVC_PH_ME_RASTER);
      dim3 grd(nctu, num_refs);
      if (use_sweep) {
        k_me_raster<true><<<grd, kRastThreads, 0, c->stream>>>(c->fr_jobs, c->me_tables, c->me_centers, nctu, 5, c->bi, (RasterBest*)c->fr_rast,
                                                               (SweepState*)c->fr_sweep, c->fr_stats, 0, 0);
        TVC_LAUNCH_CHECK(c);
      }
      k_me_raster<false><<<grd, kRastThreads, 0, c->stream>>>(c->fr_jobs, c->me_tables, c->me_centers, nctu, 5, c->bi, (RasterBest*)c->fr_rast,
                                                              nullptr, c->fr_stats, 0, 0);
      TVC_LAUNCH_CHECK(c);
    }
    const bool shared = legacy_tables && use_rast;
    if ((r = launch_search(c, cur_slot, legacy_tables ? 1 : 0, (int)n, c->fr_jobs, c->fr_int, shared ? (const RasterBest*)c->fr_rast : nullptr,
                           shared && use_sweep ? (const SweepState*)c->fr_sweep : nullptr, c->fr_stats))) return r;
    if (c->fr_int_ready) TVC_CUDA(c, cudaEventRecord(c->fr_int_ready, c->stream));      // tvc_me_frame copies the integer results from here
    if (cfg->do_frac && (r = frac_part(0, n))) return r;
  } else {
    if ((r = ensure_pipe(c))) return r;
    cudaStream_t main_stream = c->stream;
    struct Restore { tvc_ctx* c; cudaStream_t s; ~Restore() { c->stream = s; } } restore{c, main_stream};
    TVC_CUDA(c, cudaEventRecord(c->pipe_ev[0], main_stream));            // jobs, centres, counters ready
    TVC_CUDA(c, cudaStreamWaitEvent(c->pipe[0], c->pipe_ev[0], 0));
    const int rows = c->num_ctus_y, parts = pipe_chunks < rows ? pipe_chunks : rows;
    int k = 0;
    for (int rf = 0; rf < num_refs; rf++)
      for (int pt = 0; pt < parts; pt++, k++) {
        const int y0 = rows * pt / parts, y1 = rows * (pt + 1) / parts;
        const int ctu0 = y0 * c->num_ctus_x, nct = (y1 - y0) * c->num_ctus_x;
        if (nct <= 0) continue;
        cudaEvent_t ev_t = c->pipe_ev[1 + 2 * k], ev_s = c->pipe_ev[2 + 2 * k];
        c->stream = main_stream;
        if ((r = launch_tables(c, ctu0, nct, rf, 1))) return r;
        TVC_CUDA(c, cudaEventRecord(ev_t, main_stream));
        TVC_CUDA(c, cudaStreamWaitEvent(c->pipe[0], ev_t, 0));
        c->stream = c->pipe[0];
        if ((r = search_part(ctu0, nct, rf))) return r;
        if (cfg->do_frac) {
          TVC_CUDA(c, cudaEventRecord(ev_s, c->pipe[0]));
          TVC_CUDA(c, cudaStreamWaitEvent(c->pipe[1], ev_s, 0));
          c->stream = c->pipe[1];
          if ((r = frac_part(((size_t)rf * nctu + ctu0) * TVC_ME_CENSUS, (size_t)nct * TVC_ME_CENSUS))) return r;
        }
      }
    c->stream = main_stream;
    cudaStream_t last = cfg->do_frac ? c->pipe[1] : c->pipe[0];
    TVC_CUDA(c, cudaEventRecord(c->pipe_ev[0], last));                    // the side streams are in order: their last event covers all
    TVC_CUDA(c, cudaStreamWaitEvent(main_stream, c->pipe_ev[0], 0));
  }
  if (int_dev) *int_dev = c->fr_int;
  if (frac_dev) *frac_dev = cfg->do_frac ? c->fr_frac : nullptr;
  return TVC_OK;
}

int tvc_me_frame_stats(tvc_ctx* c, uint64_t stats[3])
{
  if (!c || !stats) return set_err(c, TVC_ERR_ARG, "tvc_me_frame_stats: bad argument");
  if (!c->fr_stats) return set_err(c, TVC_ERR_STATE, "tvc_me_frame_stats: no frame pre-pass has run");
  unsigned long long h[3];
  TVC_CUDA(c, cudaMemcpyAsync(h, c->fr_stats, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 3; i++) stats[i] = h[i];
  return TVC_OK;
}

int tvc_me_frame(tvc_ctx* c, int cur_slot, int num_refs, const int* ref_slots, const tvc_me_center* pred_qpel,
                 const tvc_me_frame_cfg* cfg, tvc_me_result* int_out, tvc_frac_result* frac_out)
{
  tvc_me_result* di = nullptr;
  tvc_frac_result* df = nullptr;
  // the integer results (16 B per job: 19 MB at 1080p x 4 references) leave on a side stream while the fractional search still runs
  const bool early = c && cfg && cfg->do_frac && int_out && is_pinned(int_out) && ensure_pipe(c) == TVC_OK;
  if (early && !c->fr_int_ready) TVC_CUDA(c, cudaEventCreateWithFlags(&c->fr_int_ready, cudaEventDisableTiming));
  int r = tvc_me_frame_dev(c, cur_slot, num_refs, ref_slots, pred_qpel, cfg, &di, &df);
  if (r) return r;
  const size_t n = (size_t)num_refs * c->num_ctus_x * c->num_ctus_y * TVC_ME_CENSUS;
  bool side = false;
  if (int_out) {
    if (early && !c->fr_piped_last) {
      TVC_CUDA(c, cudaStreamWaitEvent(c->pipe[0], c->fr_int_ready, 0));
      TVC_CUDA(c, cudaMemcpyAsync(int_out, di, n * sizeof(tvc_me_result), cudaMemcpyDeviceToHost, c->pipe[0]));
      side = true;
    } else
      TVC_CUDA(c, cudaMemcpyAsync(int_out, di, n * sizeof(tvc_me_result), cudaMemcpyDeviceToHost, c->stream));
  }
  if (frac_out && df) TVC_CUDA(c, cudaMemcpyAsync(frac_out, df, n * sizeof(tvc_frac_result), cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  if (side) TVC_CUDA(c, cudaStreamSynchronize(c->pipe[0]));
  return TVC_OK;
}

int tvc_me_frame_packed(tvc_ctx* c, int cur_slot, int num_refs, const int* ref_slots, const tvc_me_center* pred_qpel,
                        const tvc_me_frame_cfg* cfg, tvc_me_packed* out)
{
  if (!c || !cfg || !cfg->do_frac || !out) return set_err(c, TVC_ERR_ARG, "tvc_me_frame_packed: needs both stages (cfg->do_frac) and an output array");
  tvc_me_result* di = nullptr;
  tvc_frac_result* df = nullptr;
  int r = tvc_me_frame_dev(c, cur_slot, num_refs, ref_slots, pred_qpel, cfg, &di, &df);
  if (r) return r;
  const size_t n = (size_t)num_refs * c->num_ctus_x * c->num_ctus_y * TVC_ME_CENSUS;
  if (n > c->fr_packed_cap) {
    if (c->fr_packed) cudaFree(c->fr_packed);
    c->fr_packed = nullptr; c->fr_packed_cap = 0;
    TVC_CUDA(c, cudaMalloc(&c->fr_packed, n * sizeof(tvc_me_packed)));
    c->fr_packed_cap = n;
  }
  k_me_pack<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((int)n, di, df, (tvc_me_packed*)c->fr_packed);
  TVC_LAUNCH_CHECK(c);
  TVC_CUDA(c, cudaMemcpyAsync(out, c->fr_packed, n * sizeof(tvc_me_packed), cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  return TVC_OK;
}

// one census group (CTU, reference, predictor) queued on c->stream into the buffers of `t`: job records, integer stage, fractional
// stage, results to the ticket's pinned host copy; no synchronisation
static int me_ctu_enqueue(tvc_ctx* c, CtuTicket& t, int cur_slot, int ref_index, int ref_slot, int ctu, tvc_me_center pred_qpel,
                          const tvc_me_frame_cfg* cfg, bool fused)
{
  constexpr size_t N = TVC_ME_CENSUS;
  int r;
  if (!t.dev) {
    TVC_CUDA(c, cudaMalloc(&t.dev, N * (sizeof(tvc_me_job) + sizeof(tvc_me_result) + sizeof(tvc_frac_job) + sizeof(tvc_frac_result)) + N));
    TVC_CUDA(c, cudaHostAlloc(&t.host, N * (sizeof(tvc_me_result) + sizeof(tvc_frac_result)), cudaHostAllocDefault));
    TVC_CUDA(c, cudaEventCreateWithFlags(&t.ev, cudaEventDisableTiming));
  }
  // device layout: [int results][frac results][jobs][frac jobs][done flags] -- the results are contiguous: one copy back
  tvc_me_result* d_int = (tvc_me_result*)t.dev;
  tvc_frac_result* d_frac = (tvc_frac_result*)(d_int + N);
  tvc_me_job* d_jobs = (tvc_me_job*)(d_frac + N);
  tvc_frac_job* d_fjobs = (tvc_frac_job*)(d_jobs + N);
  uint8_t* d_done = (uint8_t*)(d_fjobs + N);
  k_me_ctu_jobs<<<(int)((N + 127) / 128), 128, 0, c->stream>>>(c->cfg.width, c->cfg.height, c->num_ctus_x, ctu, ref_index, ref_slot, pred_qpel,
                                                           *cfg, d_jobs);
  TVC_LAUNCH_CHECK(c);
  // no shared raster stage here: one block would walk the 729 raster candidates alone; the PUs that need the
  // raster walk it themselves, in parallel
  if (fused) {
    if ((r = tvc_launch_me_group(c, cur_slot, 1, d_jobs, d_int, 1, &ref_slot, 0, nullptr))) return r;
  } else if ((r = launch_search(c, cur_slot, cfg->use_tables, (int)N, d_jobs, d_int, nullptr, nullptr, nullptr))) return r;
  if (cfg->do_frac) {
    k_me_frame_frac_jobs<<<(int)((N + 127) / 128), 128, 0, c->stream>>>((int)N, d_jobs, d_int, cfg->hadamard, d_fjobs);
    TVC_LAUNCH_CHECK(c);
    if ((r = launch_frac(c, cur_slot, (int)N, d_fjobs, d_frac, true, d_done))) return r;
  }
  const size_t back = N * sizeof(tvc_me_result) + (cfg->do_frac ? N * sizeof(tvc_frac_result) : 0);
  TVC_CUDA(c, cudaMemcpyAsync(t.host, t.dev, back, cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaEventRecord(t.ev, c->stream));
  t.busy = true; t.do_frac = cfg->do_frac != 0;
  return TVC_OK;
}

static int me_ctu_check(tvc_ctx* c, int cur_slot, int ref_index, int ref_slot, int ctu, const tvc_me_frame_cfg* cfg, bool* fused)
{
  if (!c || !valid_slot(c, cur_slot) || !valid_slot(c, ref_slot) || !cfg || cfg->search_range < 1 || cfg->search_range > TVC_ME_RANGE ||
      ctu < 0 || ctu >= c->num_ctus_x * c->num_ctus_y)
    return set_err(c, TVC_ERR_ARG, "tvc_me_ctu: bad argument");
  *fused = cfg->use_tables && me_fused_enabled(c);
  if (cfg->use_tables && !*fused && (!c->me_tables || c->me_cur_slot != cur_slot || ref_index < 0 || ref_index >= c->me_num_refs ||
                                     c->me_ref_slots[ref_index] != ref_slot))
    return set_err(c, TVC_ERR_STATE, "tvc_me_ctu: tables requested but tvc_me_prepass has not run for this picture / reference");
  return ensure_census(c);
}

int tvc_me_ctu(tvc_ctx* c, int cur_slot, int ref_index, int ref_slot, int ctu, tvc_me_center pred_qpel, const tvc_me_frame_cfg* cfg,
               tvc_me_result* int_out, tvc_frac_result* frac_out)
{
  bool fused = false;
  int r = me_ctu_check(c, cur_slot, ref_index, ref_slot, ctu, cfg, &fused);
  if (r) return r;
  if (!int_out || (cfg->do_frac && !frac_out)) return set_err(c, TVC_ERR_ARG, "tvc_me_ctu: bad argument");
  constexpr size_t N = TVC_ME_CENSUS;
  CtuTicket& t = c->ctu_tickets[TVC_ME_CTU_TICKETS];       // the synchronous call's own buffers
  if ((r = me_ctu_enqueue(c, t, cur_slot, ref_index, ref_slot, ctu, pred_qpel, cfg, fused))) return r;
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  t.busy = false;
  memcpy(int_out, t.host, N * sizeof(tvc_me_result));
  if (cfg->do_frac) memcpy(frac_out, (char*)t.host + N * sizeof(tvc_me_result), N * sizeof(tvc_frac_result));
  return TVC_OK;
}

int tvc_me_ctu_async(tvc_ctx* c, int ticket, int cur_slot, int ref_index, int ref_slot, int ctu, tvc_me_center pred_qpel,
                     const tvc_me_frame_cfg* cfg)
{
  bool fused = false;
  int r = me_ctu_check(c, cur_slot, ref_index, ref_slot, ctu, cfg, &fused);
  if (r) return r;
  if (ticket < 0 || ticket >= TVC_ME_CTU_TICKETS) return set_err(c, TVC_ERR_ARG, "tvc_me_ctu_async: ticket 0..%d", TVC_ME_CTU_TICKETS - 1);
  CtuTicket& t = c->ctu_tickets[ticket];
  if (!c->spec_stream) TVC_CUDA(c, cudaStreamCreateWithFlags(&c->spec_stream, cudaStreamNonBlocking));
  if (t.busy) { TVC_CUDA(c, cudaEventSynchronize(t.ev)); t.busy = false; }      // a result nobody fetched: its buffers are reused
  // the picture slots the group reads were uploaded on c->stream: the side stream starts behind everything queued there so far
  if (!c->spec_ev) TVC_CUDA(c, cudaEventCreateWithFlags(&c->spec_ev, cudaEventDisableTiming));
  TVC_CUDA(c, cudaEventRecord(c->spec_ev, c->stream));
  TVC_CUDA(c, cudaStreamWaitEvent(c->spec_stream, c->spec_ev, 0));
  cudaStream_t main_stream = c->stream;
  struct Restore { tvc_ctx* c; cudaStream_t s; ~Restore() { c->stream = s; } } restore{c, main_stream};
  c->stream = c->spec_stream;
  return me_ctu_enqueue(c, t, cur_slot, ref_index, ref_slot, ctu, pred_qpel, cfg, fused);
}

int tvc_me_ctu_fetch(tvc_ctx* c, int ticket, tvc_me_result* int_out, tvc_frac_result* frac_out)
{
  if (!c || ticket < 0 || ticket >= TVC_ME_CTU_TICKETS || !int_out) return set_err(c, TVC_ERR_ARG, "tvc_me_ctu_fetch: bad argument");
  CtuTicket& t = c->ctu_tickets[ticket];
  if (!t.busy) return set_err(c, TVC_ERR_STATE, "tvc_me_ctu_fetch: nothing in flight on ticket %d", ticket);
  if (t.do_frac && !frac_out) return set_err(c, TVC_ERR_ARG, "tvc_me_ctu_fetch: the call had a fractional stage");
  TVC_CUDA(c, cudaEventSynchronize(t.ev));
  t.busy = false;
  constexpr size_t N = TVC_ME_CENSUS;
  memcpy(int_out, t.host, N * sizeof(tvc_me_result));
  if (t.do_frac) memcpy(frac_out, (char*)t.host + N * sizeof(tvc_me_result), N * sizeof(tvc_frac_result));
  return TVC_OK;
}

int tvc_ubench(tvc_ctx* c, int which, double* ginstr_per_s)
{
  if (!c || !ginstr_per_s || which < 0 || which > 5) return set_err(c, TVC_ERR_ARG, "tvc_ubench: bad argument");
  if (which == TVC_UB_HBM_WRITE) {
    // write-only bandwidth: what a kernel that only stores (the SAD tables) can reach at best
    const size_t bytes = (size_t)8 << 30;
    uint4* buf = nullptr;
    TVC_CUDA(c, cudaMalloc(&buf, bytes));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0, c->stream);
      static int pat = -1;
      if (pat < 0) { const char* ev = getenv("TVC_UB_WRITE_SEG16"); pat = ev ? atoi(ev) : 0; }
      if (pat > 0) k_ub_hbm_write_pat<<<kNumSM * 16, 256, 0, c->stream>>>(buf, bytes / 16, (uint32_t)rep, pat);
      else k_ub_hbm_write<<<kNumSM * 16, 256, 0, c->stream>>>(buf, bytes / 16, (uint32_t)rep);
      cudaEventRecord(e1, c->stream);
      cudaError_t e = cudaEventSynchronize(e1);
      if (e != cudaSuccess) { cudaFree(buf); return check_cuda(c, e, "tvc_ubench"); }
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
      c->launches++;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf);
    *ginstr_per_s = (double)bytes / (best * 1e-3) / 1e9;
    return TVC_OK;
  }
  const int blocks = kNumSM * 8, threads = 256, iters = 4096;
  uint32_t* d = nullptr;
  TVC_CUDA(c, cudaMalloc(&d, (size_t)blocks * threads * 4));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double per_thread = 0;
  for (int rep = 0; rep < 3; rep++) {
    if (rep == 2) cudaEventRecord(e0, c->stream);
    switch (which) {
      case TVC_UB_VABSDIFF4: k_ub_vabsdiff4<<<blocks, threads, 0, c->stream>>>(d, iters, 0x01020304u, 0x11213141u); per_thread = 8.0 * iters; break;
      case TVC_UB_IADD3: k_ub_iadd3<<<blocks, threads, 0, c->stream>>>(d, iters, 0x01020304u, 0x11213141u); per_thread = 16.0 * iters; break;
      case TVC_UB_IMAD: k_ub_imad<<<blocks, threads, 0, c->stream>>>(d, iters, 0x01020304u, 0x11213141u); per_thread = 16.0 * iters; break;
      case TVC_UB_DP2A: k_ub_dp2a<<<blocks, threads, 0, c->stream>>>(d, iters, 0x01020304u, 0x11213141u); per_thread = 16.0 * iters; break;
      default: k_ub_lds128<<<blocks, threads, 0, c->stream>>>(d, iters); per_thread = 8.0 * iters; break;
    }
    c->launches++;
  }
  cudaEventRecord(e1, c->stream);
  cudaError_t e = cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  if (e != cudaSuccess) return check_cuda(c, e, "tvc_ubench");
  *ginstr_per_s = per_thread * blocks * threads / (ms * 1e-3) / 1e9;
  return TVC_OK;
}

}  // extern "C"
