// tvc_internal.cuh -- shared definitions of the TLibCuda implementation (not part of the ABI).
#pragma once

#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/thevc_cuda.h"

namespace tvc {

constexpr int kNumSM = 148;          // B200
constexpr int kMaxSlots = 32;
constexpr int kMaxRefs = 16;
constexpr int kMaxPipeChunks = 8;   // CTU-row chunks per reference of the pipelined frame pre-pass

// Device-resident TComPicYuv: int16 planes with margins (TComPicYuv.cpp:71-127) and, for 8-bit
// content, a packed u8 copy of luma for the integer-ME SIMD path.
struct Pic {
  int16_t* buf[3] = {nullptr, nullptr, nullptr};   // allocation base (top-left of the margin)
  int16_t* org[3] = {nullptr, nullptr, nullptr};   // pel (0,0)
  int stride[3] = {0, 0, 0};                       // elements
  int w[3] = {0, 0, 0}, h[3] = {0, 0, 0};
  int mx[3] = {0, 0, 0}, my[3] = {0, 0, 0};        // margins
  uint8_t* buf8 = nullptr;                         // packed luma copy (base incl. margin)
  uint8_t* org8 = nullptr;
  int stride8 = 0;
  CUtensorMap tmap_cur;                            // u8 luma, box 64x64
  CUtensorMap tmap_cur80;                          // u8 luma, box 80x64 (the group search stages the CTU at pitch 80)
  CUtensorMap tmap_ref;                            // u8 luma, box 208x192 (192 + 16 alignment slack)
  bool has_tmap = false;
};

// plane table handed to kernels by value
struct PlaneTable {
  int16_t* org[kMaxSlots][3];
  int stride[3];
};

// buffers of one census-group call (tvc_me_ctu / tvc_me_ctu_async): device records + results, pinned host copy of the results
struct CtuTicket {
  void* dev = nullptr;
  void* host = nullptr;
  cudaEvent_t ev = nullptr;
  bool busy = false, do_frac = false;
};

struct Scratch {
  void* dev = nullptr;
  void* host = nullptr;     // pinned
  size_t bytes = 0;
};

}  // namespace tvc

struct tvc_ctx {
  tvc_config cfg;
  int bi;                         // bitIncrement = bit_depth - 8
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::vector<tvc::Pic> pics;
  tvc::PlaneTable planes;
  tvc::Scratch in, out;           // staging for host-pointer entry points
  tvc::CtuTicket ctu_tickets[TVC_ME_CTU_TICKETS + 1];   // tvc_me_ctu_async tickets; the last one serves the synchronous tvc_me_ctu
  cudaStream_t spec_stream = nullptr;                   // side stream of the asynchronous group calls
  cudaEvent_t spec_ev = nullptr;
  void* rdoq_scratch = nullptr;   // per-coefficient RDOQ working arrays (device)
  size_t rdoq_scratch_elems = 0;
  std::string err;
  uint64_t launches = 0;
  int num_ctus_x = 0, num_ctus_y = 0;
  // ME tables
  uint16_t* me_tables = nullptr;
  size_t me_table_bytes = 0;
  tvc_me_center* me_centers = nullptr;      // device, num_refs * num_ctus
  int me_num_refs = 0;
  int me_ref_slots[tvc::kMaxRefs];
  int me_cur_slot = -1;
  // frame-level ME pre-pass buffers (device): jobs, integer results, frac jobs, frac results
  tvc_me_job* fr_jobs = nullptr;
  tvc_me_result* fr_int = nullptr;
  tvc_frac_job* fr_fjobs = nullptr;
  tvc_frac_result* fr_frac = nullptr;
  void* fr_rast = nullptr;        // shared raster-stage results (RasterBest per job)
  void* fr_sweep = nullptr;       // shared first-sweep results (SweepState per job)
  unsigned long long* fr_stats = nullptr;   // device: 3 work counters (tvc_me_frame_stats)
  size_t fr_cap = 0;              // entries
  void* frac_done = nullptr;      // per census job: served by the CU-level fractional kernel (device)
  size_t frac_done_cap = 0;
  void* grp_cost = nullptr;       // k_me_group: duration of each group's CTA in the last picture-level call (device u32[4096])
  void* grp_order = nullptr;      // the groups by descending cost (device int[4096]); valid for grp_order_n groups
  int grp_order_n = 0;
  void* frac_list = nullptr;      // four counters + the lists of census jobs left to the per-PU fractional kernels (device ints)
  size_t frac_list_cap = 0;
  void* fr_packed = nullptr;      // tvc_me_frame_packed: 16-byte results (device)
  size_t fr_packed_cap = 0;
  void* bi_buf = nullptr;         // tvc_me_bipred: job / results of one refinement search (device) and its pinned staging
  void* bi_host = nullptr;
  int me_fused = -1;              // -1: TVC_ME_FUSED (default on), 0 / 1: set by tvc_me_set_fused
  // dedicated pinned staging of the asynchronous ME entry points (an event guards host reuse)
  tvc::Scratch me_stage, fr_stage;
  cudaEvent_t me_ev = nullptr, fr_ev = nullptr;
  // pipelined frame pre-pass: search stream, fractional-search stream, per-chunk events
  cudaStream_t pipe[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> pipe_ev;
  cudaEvent_t fr_int_ready = nullptr;       // recorded after the integer search of the frame pre-pass (serial form)
  bool fr_piped_last = false;
  // per-phase timing (tvc_prof_*)
  bool prof_on = false;
  struct ProfPair { int phase; cudaEvent_t a, b; };
  std::vector<ProfPair> prof_live;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[TVC_PH_COUNT] = {};
  uint64_t prof_n[TVC_PH_COUNT] = {};
  // driver entry point for tensor maps
  void* encode_tiled = nullptr;
};

namespace tvc {

int set_err(tvc_ctx* c, int code, const char* fmt, ...);
int check_cuda(tvc_ctx* c, cudaError_t e, const char* what);
int ensure_scratch(tvc_ctx* c, Scratch& s, size_t bytes);
// wait until the previous asynchronous H2D copy out of a staging buffer has executed, (re)size it
int stage_acquire(tvc_ctx* c, Scratch& s, cudaEvent_t& ev, size_t bytes);
// true when p is page-locked host memory CUDA knows about (cudaHostAlloc / cudaHostRegister): the
// host-pointer entry points then copy straight from / into the caller's buffer instead of staging
bool is_pinned(const void* p);
inline bool valid_slot(const tvc_ctx* c, int s) { return s >= 0 && s < (int)c->pics.size(); }

#define TVC_CUDA(c, expr)                                                       \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) return tvc::check_cuda((c), _e, #expr);              \
  } while (0)

#define TVC_LAUNCH_CHECK(c)                                                     \
  do {                                                                          \
    (c)->launches++;                                                            \
    cudaError_t _e = cudaGetLastError();                                        \
    if (_e != cudaSuccess) return tvc::check_cuda((c), _e, "kernel launch");    \
  } while (0)

// RAII timer of one kernel group on the context stream (no-op unless tvc_prof_enable(ctx, 1))
struct ProfScope {
  tvc_ctx* c; cudaEvent_t b = nullptr;
  ProfScope(tvc_ctx* ctx, int phase) : c(ctx)
  {
    if (!c || !c->prof_on) return;
    cudaEvent_t a = take(); b = take();
    cudaEventRecord(a, c->stream);
    c->prof_live.push_back({phase, a, b});
  }
  ~ProfScope() { if (b) cudaEventRecord(b, c->stream); }
  cudaEvent_t take()
  {
    cudaEvent_t e = nullptr;
    if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
  }
};

// coding scans (initSigLastScan, TComRom.cpp:564-690) on the device: [scan_idx 0 diag,1 hor,2 ver][log2-2]
struct ScanTables { const uint16_t* s[3][4]; };
int ensure_scans(tvc_ctx* c, ScanTables& st);          // tvc_tq.cu
// range-checks a host TU list (grouped by ascending log2_size) and counts TUs per size
int validate_tus(tvc_ctx* c, int plane_slot, int n, const tvc_tu* tus, size_t coef_elems, int counts[4]);   // tvc_tq.cu
// the per-record test of validate_tus (ordering of the list is checked by the caller)
inline bool tu_record_ok(const Pic& p, const tvc_tu& t, size_t coef_elems)
{
  const int N = 1 << (t.log2_size & 7);
  return !(t.log2_size < 2 || t.log2_size > 5 || t.plane < 0 || t.plane > 2 || t.scan_idx < 0 || t.scan_idx > 2 ||
           t.x < 0 || t.y < 0 || t.x + N > p.w[t.plane] + p.mx[t.plane] || t.y + N > p.h[t.plane] + p.my[t.plane] ||
           t.qp_rem < 0 || t.qp_rem > 5 || t.qp_per < 0 || t.qp_per > 12 || t.base_per < 0 || t.base_per > 12 ||
           t.coef_offset < 0 || (size_t)t.coef_offset + (size_t)N * N > coef_elems ||
           ((t.flags & (TVC_TU_DST | TVC_TU_SKIP)) && t.log2_size != 2));
}
// dequant (optional) + inverse transform of a device TU list, one launch per size; recon = Clip(pred + resi) when pred_slot >= 0
int launch_inv(tvc_ctx* c, int resi_slot, int pred_slot, int recon_slot, const int counts[4], const tvc_tu* tus_dev,
               const int32_t* levels_dev, int dequant);                                                          // tvc_tq.cu

}  // namespace tvc
// group search over census groups ([group][593] jobs) and the per-PU kernel over a device-resident job list (tvc_me.cu)
int tvc_launch_me_group(tvc_ctx* c, int cur_slot, int ngroups, const tvc_me_job* jobs_dev, tvc_me_result* out_dev, int num_refs,
                        const int* ref_slots, int ref_index_fixed, unsigned long long* stats);
int tvc_launch_me_search_list(tvc_ctx* c, int cur_slot, const int* list_dev, const int* count_dev, const tvc_me_job* jobs_dev,
                              tvc_me_result* out_dev);
namespace tvc {
// true unless TVC_ME_FUSED=0: the census entry points (tvc_me_frame, tvc_me_ctu) search with the group kernel instead of SAD tables
bool me_fused_enabled(const tvc_ctx* c);

// ---- table layout of the ME pre-pass (shared by producer and consumers) ----------------------
// T[ref][ctu][cand 129*129][by 16][q 4][par 2][bx 4] uint16: one candidate = 1 KB contiguous
// (cand = (dy+64)*129 + (dx+64)); inside it block row `by`, then the 16-column quarter `q`, then one
// 16-byte granule = even-row SADs of the quarter's four 4x4 blocks followed by their odd-row SADs.
// Producer: the four lanes of one dx write 64 contiguous bytes per block row (full sectors);
// consumers: a PU's granules of one candidate all lie inside that candidate's 1 KB, a full-candidate
// read (raster pre-pass) is one coalesced 1 KB load.
constexpr int kMeR = TVC_ME_RANGE;
constexpr int kMeC = TVC_ME_CAND;                       // 129
constexpr int kMeCands = kMeC * kMeC;                   // 16641
constexpr size_t kMeGranule = 8;                        // uint16 per (cand, by, q): par 2 x bx 4
constexpr int kMeCandGranules = 64;                     // granules per candidate (16 block rows x 4 quarters)
constexpr size_t kMeCtuElems = (size_t)kMeCands * kMeCandGranules * kMeGranule;   // uint16 per (ref, ctu)

// granule index of (candidate, block row, quarter) inside one (ref, ctu) table
__host__ __device__ inline uint32_t me_granule(int dy, int dx, int by, int q)
{
  return (uint32_t)(((dy + kMeR) * kMeC + (dx + kMeR)) * kMeCandGranules + by * 4 + q);
}

}  // namespace tvc
