// tvc_me_group.cu -- integer motion search of a whole (CTU, reference) group in ONE CTA, SADs computed on demand.
//
// Round 1 wrote the full +-64 SAD tables of every (CTU, reference) to HBM (17 MB each, 34.8 GB per 1080p picture with four
// references) and only 18 % of those bytes were ever read back.  This kernel replaces table + raster + search for the census
// groups of the frame pre-pass and of tvc_me_ctu: the 208x192 u8 search window and the 64x64 u8 CTU are staged into shared
// memory once (TMA), and a SAD is only ever computed for a candidate some PU actually visits.
//
// All 593 PUs of the census search the same window around the same predictor (xSetSearchRange, TEncSearch.cpp:4209-4225), so the
// CTA works CANDIDATE-major in lock-step rounds of the TZ state machine (xTZSearch, :4302-4474):
//   1. every PU thread names the REQUEST its next step makes -- (start + zero vector), a diamond sweep of <= 76 points around a
//      centre, the two points of xTZ2PointSearch around a best point, the raster grid -- and the distinct requests of the round
//      are collected in a shared-memory hash table: PUs that make the same request visit the same candidates;
//   2. the candidates of the requests become slots (request, visiting index, position) processed in chunks: one warp per slot
//      turns the 64x64 absolute differences (VABSDIFF4, window rows from shared memory, conflict-free 128-bit loads) into the
//      even-row / odd-row SADs of the 256 4x4 blocks (the FEN row sub-sampling of TEncSearch.cpp:324-330 needs them apart) and
//      integrates them into two 17x17 integral images; every PU thread reads the SAD of the slots of ITS request with four loads;
//   3. every PU thread replays the reference's sequential strict-'<' update on its own costs (ordered arg-min: cost and visiting
//      index packed into one word) and moves its state machine on.
// Candidates outside the staged window (zero vector of a far predictor, PUs at the picture border whose clipMv differs from
// the CTU's) take the same route with their reference rows read from the u8 plane in global memory.  Results are those of
// xTZSearch for every PU: MV, ruiSAD and the number of SADs evaluated (tests/test_gpu_parity.py compares all three with the oracle).
#include "tvc_internal.cuh"
#include "tvc_me.cuh"

namespace tvc {

constexpr int kGT = 640;                       // threads: one per census PU (593), whole warps
constexpr int kGWarps = kGT / 32;
constexpr int kGChunk = 40;                    // candidate slots per chunk (two per warp)
constexpr int kGWinW = 208, kGWinH = 192;      // staged window: 64 + 2 * 64 columns + 16 bytes of TMA alignment slack
constexpr int kGCurP = 80;                     // pitch of the staged CTU (80 = 20 words: rows 0..7 start in distinct bank groups)
constexpr int kGHash = 1024;                   // hash slots of the request table (<= 593 distinct requests per round)
constexpr int kGSlots = 4096;                  // candidate slots per pass
constexpr int kGClsPerPass = kGSlots / 76;     // 53 sweep requests fill one pass
constexpr int kGMaxRounds = 64;
constexpr uint32_t kNone = 0xFFFFFFFFu;

enum { PH_INIT = 0, PH_FIRST = 1, PH_TWO_F = 2, PH_RASTER = 3, PH_REFINE = 4, PH_TWO_R = 5, PH_DONE = 6, PH_FALLBACK = 7 };
enum { RQ_INIT = 1, RQ_SWEEP = 2, RQ_TWO = 3, RQ_RASTER = 4 };       // request kinds (0 = empty hash slot)

struct GroupMaps {
  CUtensorMap cur;                 // u8 luma of the current picture, box 80 x 64
  CUtensorMap ref[8];              // u8 luma of each reference, box 208 x 192
  const uint8_t* ref8[8];          // the same planes, pel (0,0)
  int stride8;
};

// one request = one distinct (kind, anchor, window) among the PUs of a round: every PU that makes it visits the same candidates
struct Request { int16_t kind, aux, ax, ay, lx, ty, rx, by; };
struct Slot { int16_t x, y; uint16_t req, idx; };      // candidate (absolute integer MV) of request `req`, visiting index idx

struct GroupSmem {
  alignas(128) uint8_t win[kGWinW * kGWinH];
  alignas(128) uint8_t cur[kGCurP * 64];
  alignas(16) uint32_t IE[kGChunk][17 * 17];   // integral of the even-row block SADs
  uint32_t IA[kGChunk][17 * 17];               // integral of even + odd
  alignas(8) unsigned long long hkey[kGHash];  // request keys (open addressing)
  unsigned long long rcnt[kGT];                // per request: valid candidates per sweep round (7 x 5 bits)
  Request req[kGT];
  uint16_t hreq[kGHash];                       // hash slot -> request index
  Slot slot[kGSlots];
  uint32_t mvc[kGChunk];
  int nreq, nslot;
  alignas(8) uint64_t bar;
};

struct Win { int lx, ty, rx, by; };

// the two points of xTZ2PointSearch (TEncSearch.cpp:351-476) around the best point, with the border tests named there
__device__ __forceinline__ void two_points(const Win& w, int bx, int by, int point_nr, int (&x)[2], int (&y)[2], bool (&v)[2])
{
  const bool up = (by - 1) >= w.ty, dn = (by + 1) <= w.by, lf = (bx - 1) >= w.lx, rt = (bx + 1) <= w.rx;
  x[0] = x[1] = bx; y[0] = y[1] = by; v[0] = v[1] = false;
  switch (point_nr) {
    case 1: x[0] = bx - 1; v[0] = lf; y[1] = by - 1; v[1] = up; break;
    case 2: x[0] = bx - 1; y[0] = by - 1; v[0] = up && lf; x[1] = bx + 1; y[1] = by - 1; v[1] = up && rt; break;
    case 3: y[0] = by - 1; v[0] = up; x[1] = bx + 1; v[1] = rt; break;
    case 4: x[0] = bx - 1; y[0] = by + 1; v[0] = lf && dn; x[1] = bx - 1; y[1] = by - 1; v[1] = lf && up; break;
    case 5: x[0] = bx + 1; y[0] = by - 1; v[0] = rt && up; x[1] = bx + 1; y[1] = by + 1; v[1] = rt && dn; break;
    case 6: x[0] = bx - 1; v[0] = lf; y[1] = by + 1; v[1] = dn; break;
    case 7: x[0] = bx - 1; y[0] = by + 1; v[0] = dn && lf; x[1] = bx + 1; y[1] = by + 1; v[1] = dn && rt; break;
    case 8: x[0] = bx + 1; v[0] = rt; y[1] = by + 1; v[1] = dn; break;
    default: break;
  }
}

// even-row / odd-row SADs of the 256 4x4 blocks of one candidate -> the two images (not yet integrated).  One warp.
// Lane = (q: 16-column quarter, g: row phase); the lane takes rows g, g + 8, ..., g + 56: row pitches of 208 and 80 bytes put
// the eight phases of a quarter-warp into eight distinct bank groups, so every 128-bit load is conflict-free.  Rows r and r + 2
// (same 4x4 block, same parity) sit on lanes g and g ^ 2, the other parity on g ^ 1: two shuffle steps finish a block row.
template <int WO>
__device__ __forceinline__ void grid_rows_window(const uint8_t* __restrict__ wb, const uint8_t* __restrict__ cb, int sh, uint32_t (&v)[8][2])
{
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const uint4 lo = *reinterpret_cast<const uint4*>(wb + k * 8 * kGWinW);
    const uint4 hi = *reinterpret_cast<const uint4*>(wb + k * 8 * kGWinW + 16);
    const uint4 c4 = *reinterpret_cast<const uint4*>(cb + k * 8 * kGCurP);
    const uint32_t W[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    uint32_t R[4];
#pragma unroll
    for (int m = 0; m < 4; m++) R[m] = __funnelshift_r(W[WO + m], W[WO + m + 1], sh);
    v[k][0] = vsad4_acc(c4.x, R[0], 0u) | (vsad4_acc(c4.y, R[1], 0u) << 16);
    v[k][1] = vsad4_acc(c4.z, R[2], 0u) | (vsad4_acc(c4.w, R[3], 0u) << 16);
  }
}

__device__ __forceinline__ void grid_rows_global(const uint8_t* __restrict__ rp, int stride8, const uint8_t* __restrict__ cb, uint32_t (&v)[8][2])
{
  const int a4 = (int)((uintptr_t)rp & 3), sh = a4 * 8;
  const uint8_t* base = rp - a4;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(base + (ptrdiff_t)k * 8 * stride8);
    const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2), w3 = __ldg(p + 3), w4 = a4 ? __ldg(p + 4) : 0u;
    const uint4 c4 = *reinterpret_cast<const uint4*>(cb + k * 8 * kGCurP);
    const uint32_t R0 = __funnelshift_r(w0, w1, sh), R1 = __funnelshift_r(w1, w2, sh), R2 = __funnelshift_r(w2, w3, sh),
                   R3 = __funnelshift_r(w3, w4, sh);
    v[k][0] = vsad4_acc(c4.x, R0, 0u) | (vsad4_acc(c4.y, R1, 0u) << 16);
    v[k][1] = vsad4_acc(c4.z, R2, 0u) | (vsad4_acc(c4.w, R3, 0u) << 16);
  }
}

__device__ __forceinline__ void grid_store(uint32_t (&v)[8][2], int lane, uint32_t* __restrict__ IEc, uint32_t* __restrict__ IAc)
{
  const int q = lane >> 3, g = lane & 7;
#pragma unroll
  for (int k = 0; k < 8; k++) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      uint32_t same = v[k][h] + __shfl_xor_sync(0xffffffffu, v[k][h], 2);       // both lines of this parity (halves <= 2040: no carry)
      const uint32_t other = __shfl_xor_sync(0xffffffffu, same, 1);             // the other parity
      if ((g & 3) == 0) {                                                       // g = 0 / 4: even parity in `same`, odd in `other`
        const int by = 2 * k + (g >> 2), bx = 4 * q + 2 * h;
        uint32_t* e = IEc + (by + 1) * 17 + bx + 1;
        uint32_t* a = IAc + (by + 1) * 17 + bx + 1;
        const uint32_t e0 = same & 0xffffu, e1 = same >> 16;
        e[0] = e0; e[1] = e1;
        a[0] = e0 + (other & 0xffffu); a[1] = e1 + (other >> 16);
      }
    }
  }
}

// raw block SADs of one candidate -> integral images, by the warp that computed them (no block barrier): row prefix sums
// (lane = image x row), then column prefix sums (lane = image x column)
__device__ __forceinline__ void grid_integrate(uint32_t* __restrict__ IEc, uint32_t* __restrict__ IAc, int lane)
{
  __syncwarp();
  {
    uint32_t* p = ((lane & 16) ? IAc : IEc) + ((lane & 15) + 1) * 17 + 1;
    uint32_t s = 0;
#pragma unroll
    for (int x = 0; x < 16; x++) { s += p[x]; p[x] = s; }
  }
  __syncwarp();
  {
    uint32_t* p = ((lane & 16) ? IAc : IEc) + 17 + (lane & 15) + 1;
    uint32_t s = 0;
#pragma unroll
    for (int r = 0; r < 16; r++) { s += p[r * 17]; p[r * 17] = s; }
  }
}

__global__ void __launch_bounds__(kGT, 1)
k_me_group(const __grid_constant__ GroupMaps maps, const tvc_me_job* __restrict__ jobs, tvc_me_result* __restrict__ out,
           const tvc_census_pu* __restrict__ census, int pic_w, int pic_h, int mx, int my,
           int ref_index_fixed, int* __restrict__ fb_list, int* __restrict__ fb_count, unsigned long long* __restrict__ stats)
{
  extern __shared__ __align__(128) uint8_t g_smem[];
  GroupSmem& S = *reinterpret_cast<GroupSmem*>(g_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t gbase = (size_t)blockIdx.x * TVC_ME_CENSUS;
  const tvc_me_job j0 = jobs[gbase];                       // the 64x64 PU: its window is the group's
  const int x0 = j0.x, y0 = j0.y;                          // CTU origin (census PU 0 sits at it)
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
  const int wx = mx + x0 + cenx - kMeR, e16 = wx & 15;
  if (tid == 0) {
    mbar_init(&S.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&S.bar, kGWinW * kGWinH + kGCurP * 64);
    tma_load_2d(S.win, &maps.ref[ref], wx - e16, my + y0 + ceny - kMeR, &S.bar);
    tma_load_2d(S.cur, &maps.cur, mx + x0, my + y0, &S.bar);
  }
  // while the copies fly: per-PU set-up, request table and integral borders cleared
  for (int i = tid; i < kGHash; i += kGT) S.hkey[i] = 0ull;
  for (int i = tid; i < kGChunk * 17; i += kGT) {
    const int c = i / 17, k = i % 17;
    S.IE[c][k] = 0; S.IA[c][k] = 0; S.IE[c][k * 17] = 0; S.IA[c][k * 17] = 0;
  }
  if (tid == 0) { S.nreq = 0; S.nslot = 0; }

  const bool is_pu = tid < TVC_ME_CENSUS;
  const tvc_me_job jb = jobs[gbase + (is_pu ? tid : 0)];
  const tvc_census_pu cp = census[is_pu ? tid : 0];
  const Win win = {jb.lx, jb.ty, jb.rx, jb.by};
  const bool same = jb.lx == j0.lx && jb.ty == j0.ty && jb.rx == j0.rx && jb.by == j0.by;
  const int sub = (jb.fen && cp.h > 8) ? 1 : 0;
  const int c00 = (cp.y >> 2) * 17 + (cp.x >> 2), c01 = (cp.y >> 2) * 17 + ((cp.x + cp.w) >> 2);
  const int c10 = ((cp.y + cp.h) >> 2) * 17 + (cp.x >> 2), c11 = ((cp.y + cp.h) >> 2) * 17 + ((cp.x + cp.w) >> 2);
  const int dmax = j0.search_range, raster = 5;              // the census shares range, predictor and lambda
  int nsweep = 0;
  for (int d = 1; d <= dmax; d <<= 1) nsweep += round_size(d);
  const int rnx = (win.rx - win.lx) / raster + 1, rny = (win.by - win.ty) / raster + 1;

  int phase = (is_pu && jb.w > 0 && jb.mode == TVC_ME_TZ) ? PH_INIT : PH_DONE;
  if (is_pu && jb.w > 0 && jb.mode != TVC_ME_TZ) phase = PH_FALLBACK;
  uint32_t best = kNone, best_dist = 0, n_sads = 0;
  int best_x = 0, best_y = 0, point_nr = 0, sx = 0, sy = 0;
  uint32_t rmin[7];
  uint32_t acc = kNone;
  unsigned long long st_win = 0, st_glob = 0, st_rounds = 0;

  mbar_wait(&S.bar, 0);
  __syncthreads();

  for (int round = 0; round < kGMaxRounds; round++) {
    // ------------------------------------------------------------------ 1. requests: distinct (kind, anchor, window) of this round
    acc = kNone;
    int kind = 0, aux = 0, ax = 0, ay = 0;
    if (phase == PH_INIT) { kind = RQ_INIT; ax = jb.startx; ay = jb.starty; rmin[0] = rmin[1] = kNone; }
    else if (phase == PH_FIRST || phase == PH_REFINE) {
      kind = RQ_SWEEP; ax = sx; ay = sy;
      if (phase == PH_FIRST) {
#pragma unroll
        for (int r = 0; r < 7; r++) rmin[r] = kNone;
      }
    } else if (phase == PH_TWO_F || phase == PH_TWO_R) { kind = RQ_TWO; ax = best_x; ay = best_y; aux = point_nr; }
    else if (phase == PH_RASTER) kind = RQ_RASTER;
    if (!__syncthreads_or(kind != 0)) break;
    int hslot = 0;
    bool owner = false;
    if (kind) {
      // kind 3 bits | aux 4 bits | window id 10 bits (0: the group's window, else the PU's own) | ax 16 bits | ay 16 bits
      const unsigned long long key = (unsigned long long)kind | ((unsigned long long)(aux & 15) << 3) |
                                     ((unsigned long long)(same ? 0 : tid + 1) << 7) | ((unsigned long long)(uint16_t)ax << 17) |
                                     ((unsigned long long)(uint16_t)ay << 33);
      uint32_t h = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 54);
      for (;;) {
        const unsigned long long old = atomicCAS(&S.hkey[h], 0ull, key);
        if (old == 0ull) { owner = true; break; }
        if (old == key) break;
        h = (h + 1) & (kGHash - 1);
      }
      hslot = (int)h;
    }
    __syncthreads();
    if (owner) {
      const int r = atomicAdd(&S.nreq, 1);
      S.hreq[hslot] = (uint16_t)r;
      S.req[r] = Request{(int16_t)kind, (int16_t)aux, (int16_t)ax, (int16_t)ay, (int16_t)win.lx, (int16_t)win.ty, (int16_t)win.rx, (int16_t)win.by};
      S.rcnt[r] = 0ull;
    }
    __syncthreads();
    const int my_req = kind ? (int)S.hreq[hslot] : -1;
    const int nreq = S.nreq;
    st_rounds++;

    // ------------------------------------------------------------------ 2. passes: candidate slots of a run of requests, chunk by chunk
    for (int r0 = 0; r0 < nreq;) {
      int r1 = r0, budget = 0;
      while (r1 < nreq) {
        const int b = S.req[r1].kind == RQ_RASTER ? 1024 : 76;
        if (budget + b > kGSlots) break;
        budget += b; r1++;
      }
      if (tid == 0) S.nslot = 0;
      __syncthreads();
      for (int it = tid; it < (r1 - r0) * 76; it += kGT) {
        const int r = r0 + it / 76, c = it % 76;
        const Request rq = S.req[r];
        const Win rw = {rq.lx, rq.ty, rq.rx, rq.by};
        int x = 0, y = 0;
        bool v = false;
        if (rq.kind == RQ_SWEEP) {
          if (c < nsweep) {
            int d, i, pt;
            uint32_t dist;
            sweep_slot(c, dmax, d, i);
            v = diamond_cand(rw, rq.ax, rq.ay, d, i, x, y, pt, dist);
            if (v) atomicAdd(&S.rcnt[r], 1ull << (5 * (c < 4 ? 0 : (c < 28 ? 1 + ((c - 4) >> 3) : 4 + ((c - 28) >> 4)))));
          }
        } else if (rq.kind == RQ_INIT) {
          v = c < 2;                             // start point, then the zero vector (TEncSearch.cpp:4320, 4336-4339)
          if (c == 0) { x = rq.ax; y = rq.ay; }
        } else if (rq.kind == RQ_TWO) {
          if (c < 2) {
            int tx[2], ty2[2];
            bool tv[2];
            two_points(rw, rq.ax, rq.ay, rq.aux, tx, ty2, tv);
            x = tx[c]; y = ty2[c]; v = tv[c];
            if (v) atomicAdd(&S.rcnt[r], 1ull);
          }
        }
        if (v) {
          const int k = atomicAdd(&S.nslot, 1);
          S.slot[k] = Slot{(int16_t)x, (int16_t)y, (uint16_t)r, (uint16_t)c};
        }
      }
      for (int r = r0; r < r1; r++) {
        const Request rq = S.req[r];
        if (rq.kind != RQ_RASTER) continue;      // uniform: every thread reads the same record
        const int nx = (rq.rx - rq.lx) / raster + 1, ny = (rq.by - rq.ty) / raster + 1;
        for (int i = tid; i < nx * ny; i += kGT) {
          const int k = atomicAdd(&S.nslot, 1);
          S.slot[k] = Slot{(int16_t)(rq.lx + (i % nx) * raster), (int16_t)(rq.ty + (i / nx) * raster), (uint16_t)r, (uint16_t)i};
        }
      }
      __syncthreads();
      const int nslot = S.nslot;

      for (int cb0 = 0; cb0 < nslot; cb0 += kGChunk) {
        const int nc = min(kGChunk, nslot - cb0);
        // ---- 2a. one warp per candidate: block SAD grid, then its two integral images
        for (int ci = warp; ci < nc; ci += kGWarps) {
          const Slot sl = S.slot[cb0 + ci];
          const int x = sl.x, y = sl.y, dx = x - cenx, dy = y - ceny;
          if (lane == 0) S.mvc[ci] = mv_cost(j0.lambda_cost, x, y, 2, j0.predx, j0.predy);
          const int q = lane >> 3, g = lane & 7;
          const uint8_t* cbp = S.cur + g * kGCurP + 16 * q;
          uint32_t v[8][2];
          if (dx >= -kMeR && dx <= kMeR && dy >= -kMeR && dy <= kMeR) {
            const int col = dx + kMeR + e16, row = dy + kMeR;
            const int a = col & 15, sh = (a & 3) * 8;
            const uint8_t* wb = S.win + (row + g) * kGWinW + (col & ~15) + 16 * q;
            switch (a >> 2) {
              case 0: grid_rows_window<0>(wb, cbp, sh, v); break;
              case 1: grid_rows_window<1>(wb, cbp, sh, v); break;
              case 2: grid_rows_window<2>(wb, cbp, sh, v); break;
              default: grid_rows_window<3>(wb, cbp, sh, v); break;
            }
            if (lane == 0) st_win++;
          } else {
            // beyond the staged window (zero vector of a far predictor, border PUs clipped differently): rows from the u8 plane
            const uint8_t* rp = maps.ref8[ref] + (ptrdiff_t)(y0 + y + g) * maps.stride8 + (x0 + x + 16 * q);
            grid_rows_global(rp, maps.stride8, cbp, v);
            if (lane == 0) st_glob++;
          }
          grid_store(v, lane, S.IE[ci], S.IA[ci]);
          grid_integrate(S.IE[ci], S.IA[ci], lane);
        }
        __syncthreads();
        // ---- 3. every PU: cost of the chunk's candidates that belong to its request, ordered arg-min
        if (my_req >= 0) {
          for (int ci = 0; ci < nc; ci++) {
            const Slot sl = S.slot[cb0 + ci];
            if ((int)sl.req != my_req) continue;
            const uint32_t* I = sub ? S.IE[ci] : S.IA[ci];
            const uint32_t sad = I[c11] - I[c01] - I[c10] + I[c00];
            const uint32_t cost = (sad << sub) + S.mvc[ci];
            const uint32_t c = sl.idx;
            if (phase == PH_INIT) {
              if (c == 0) rmin[0] = cost;
              else rmin[1] = cost;
            } else if (phase == PH_FIRST) {
              const int rr = c < 4 ? 0 : (c < 28 ? 1 + (int)((c - 4) >> 3) : 4 + (int)((c - 28) >> 4));
              const uint32_t i = c < 4 ? c : (c < 28 ? ((c - 4) & 7u) : ((c - 28) & 15u));
              const uint32_t pk = (cost << 4) | i;
#pragma unroll
              for (int r = 0; r < 7; r++)
                if (r == rr) rmin[r] = min(rmin[r], pk);
            } else if (phase == PH_REFINE) acc = min(acc, (cost << 7) | c);
            else if (phase == PH_RASTER) acc = min(acc, (cost << 10) | c);
            else acc = min(acc, (cost << 1) | c);
          }
        }
        __syncthreads();
      }
      r0 = r1;
    }

    // ------------------------------------------------------------------ 4. replay the sequential update, move the state machine
    const unsigned long long cnt = my_req >= 0 ? S.rcnt[my_req] : 0ull;
    bool to_after_first = false, to_refine_entry = false;
    if (phase == PH_INIT) {
      best = rmin[0]; best_x = jb.startx; best_y = jb.starty;
      if (rmin[1] < rmin[0]) { best = rmin[1]; best_x = 0; best_y = 0; }
      best_dist = 0; point_nr = 0; n_sads = 2;
      sx = best_x; sy = best_y;
      phase = PH_FIRST;
    } else if (phase == PH_FIRST) {
      int best_round = 0, r = 0;
      for (int d = 1; d <= dmax; d <<= 1, r++) {               // xTZSearch first search (:4346-4361)
        uint32_t pk = kNone;
#pragma unroll
        for (int k = 0; k < 7; k++)
          if (k == r) pk = rmin[k];
        best_round += 1;
        n_sads += (uint32_t)((cnt >> (5 * r)) & 31ull);
        if (pk != kNone && (pk >> 4) < best) {
          int pt;
          diamond_cand(win, sx, sy, d, (int)(pk & 15u), best_x, best_y, pt, best_dist);
          best = pk >> 4; point_nr = pt; best_round = 0;
        }
        if (best_round >= 3) break;                            // bFirstSearchStop, uiFirstSearchRounds = 3
      }
      if (best_dist == 1) { best_dist = 0; phase = PH_TWO_F; }  // :4382-4386
      else to_after_first = true;
    } else if (phase == PH_TWO_F || phase == PH_TWO_R) {
      n_sads += (uint32_t)(cnt & 31ull);
      if (acc != kNone && (acc >> 1) < best) {
        int x[2], y[2];
        bool v[2];
        two_points(win, best_x, best_y, point_nr, x, y, v);
        best = acc >> 1; best_x = x[acc & 1u]; best_y = y[acc & 1u];
        best_dist = 2; point_nr = 0;
      }
      if (phase == PH_TWO_F) to_after_first = true;
      else to_refine_entry = true;
    } else if (phase == PH_RASTER) {
      n_sads += (uint32_t)(rnx * rny);
      if (acc != kNone && (acc >> 10) < best) {
        const int idx = (int)(acc & 1023u);
        best = acc >> 10;
        best_x = win.lx + (idx % rnx) * raster; best_y = win.ty + (idx / rnx) * raster;
        best_dist = raster; point_nr = 0;
      }
      to_refine_entry = true;
    } else if (phase == PH_REFINE) {
#pragma unroll
      for (int r = 0; r < 7; r++) n_sads += (uint32_t)((cnt >> (5 * r)) & 31ull);
      if (acc != kNone && (acc >> 7) < best) {
        int d, i, pt;
        sweep_slot((int)(acc & 127u), dmax, d, i);
        diamond_cand(win, sx, sy, d, i, best_x, best_y, pt, best_dist);
        best = acc >> 7; point_nr = pt;
      }
      if (best_dist == 1) {
        best_dist = 0;
        if (point_nr != 0) phase = PH_TWO_R;
        else to_refine_entry = true;
      } else to_refine_entry = true;
    }
    if (to_after_first) {                                       // raster stage (:4389-4400)
      if ((int)best_dist > raster) { best_dist = raster; phase = PH_RASTER; }
      else to_refine_entry = true;
    }
    if (to_refine_entry) {                                      // star refinement (:4435-4468)
      if (best_dist > 0) { sx = best_x; sy = best_y; best_dist = 0; point_nr = 0; phase = PH_REFINE; }
      else phase = PH_DONE;
    }
    // request table of the next round
    for (int i = tid; i < kGHash; i += kGT) S.hkey[i] = 0ull;
    if (tid == 0) S.nreq = 0;
    __syncthreads();
  }

  if (is_pu) {
    if (phase == PH_DONE) {
      tvc_me_result r;
      if (jb.w > 0) { r.mvx = best_x; r.mvy = best_y; r.sad = best - mv_cost(jb.lambda_cost, best_x, best_y, 2, jb.predx, jb.predy); r.n_sads = n_sads; }
      else { r.mvx = 0; r.mvy = 0; r.sad = 0; r.n_sads = 0; }
      out[gbase + tid] = r;
    } else {
      // handed back: the per-PU kernel finishes it from the pictures (a mode other than TZ, round cap)
      const int k = atomicAdd(fb_count, 1);
      fb_list[k] = (int)(gbase + tid);
    }
  }
  if (stats && lane == 0) {
    if (st_win) atomicAdd(&stats[0], st_win);
    if (st_glob) atomicAdd(&stats[1], st_glob);
    if (tid == 0) atomicAdd(&stats[2], st_rounds);
  }
}

}  // namespace tvc

using namespace tvc;

// group kernel over `ngroups` census groups (jobs laid out [group][593]); PUs it hands back are finished by the per-PU kernel.
// ref_index_fixed >= 0: every group searches maps.ref[ref_index_fixed] (tvc_me_ctu), else the group's own job.ref_index.
int tvc_launch_me_group(tvc_ctx* c, int cur_slot, int ngroups, const tvc_me_job* jobs_dev, tvc_me_result* out_dev, int num_refs,
                        const int* ref_slots, int ref_index_fixed, unsigned long long* stats)
{
  if (!c->pics[cur_slot].has_tmap || c->cfg.bit_depth != 8) return set_err(c, TVC_ERR_STATE, "group search: needs the 8-bit u8 planes and tensor maps");
  if (!c->grp_census) {
    tvc_census_pu cen[TVC_ME_CENSUS];
    tvc_me_census(cen);
    TVC_CUDA(c, cudaMalloc(&c->grp_census, sizeof(cen)));
    TVC_CUDA(c, cudaMalloc(&c->grp_fb_count, sizeof(int)));
    TVC_CUDA(c, cudaMemcpyAsync(c->grp_census, cen, sizeof(cen), cudaMemcpyHostToDevice, c->stream));
    TVC_CUDA(c, cudaStreamSynchronize(c->stream));         // the source is a stack buffer
    TVC_CUDA(c, cudaFuncSetAttribute(k_me_group, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GroupSmem)));
  }
  const size_t njobs = (size_t)ngroups * TVC_ME_CENSUS;
  if (njobs > c->grp_fb_cap) {
    if (c->grp_fb_list) cudaFree(c->grp_fb_list);
    c->grp_fb_list = nullptr; c->grp_fb_cap = 0;
    TVC_CUDA(c, cudaMalloc(&c->grp_fb_list, njobs * sizeof(int)));
    c->grp_fb_cap = njobs;
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
  TVC_CUDA(c, cudaMemsetAsync(c->grp_fb_count, 0, sizeof(int), c->stream));
  {
    ProfScope ps(c, TVC_PH_ME_SEARCH);
    k_me_group<<<ngroups, kGT, sizeof(GroupSmem), c->stream>>>(maps, jobs_dev, out_dev, (const tvc_census_pu*)c->grp_census,
                                                               c->cfg.width, c->cfg.height, p.mx[0], p.my[0], ref_index_fixed, c->grp_fb_list,
                                                               c->grp_fb_count, stats);
    TVC_LAUNCH_CHECK(c);
    int r = tvc_launch_me_search_list(c, cur_slot, c->grp_fb_list, c->grp_fb_count, jobs_dev, out_dev);
    if (r) return r;
  }
  return TVC_OK;
}
