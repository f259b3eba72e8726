// tvc_tq.cu -- forward/inverse core transforms fused with (de)quantisation, batched over TUs.
//
// Replaces TComTrQuant::transformNxN / invtransformNxN and the functions under them
// (TComTrQuant.cpp:417-972 partialButterfly*, fastForwardDst/fastInverseDst, xTrMxN/xITrMxN;
// :977-1100 signBitHidingHDQ; :1102-1270 xQuant non-RDOQ branch; :1272-1355 xDeQuant flat branch;
// :1542-1704 xT/xIT/xTransformSkip/xITransformSkip).
//
// Mapping: one TU uses N lanes of a warp (N = transform size), so a warp carries 32/N TUs.  Lane j
// owns row j of the block for the first 1-D pass and row j of the intermediate for the second; the
// transform matrix is read from constant memory with warp-uniform addresses (immediate operands
// after unrolling), the data rows live in registers.  The butterflies of the reference are exact
// integer arithmetic, so the plain dot products below give the same numbers; the reference's
// (short) truncation (forward) and Clip3(-32768,32767) (inverse) are applied at the same points.
#include "tvc_internal.cuh"

namespace tvc {

// ---- transform matrices ------------------------------------------------------------------------
// HEVC core transform (H.265 8.6.4.2; reference tables TComRom.cpp:303-377): T32[k][j] is
// +-mag[(2j+1)k folded into a quarter period]; smaller sizes are even-row sub-sampled.
struct T32Table { int8_t v[32][32]; };
constexpr int kCosMag[33] = {64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                             61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0};
constexpr int t32_coef(int k, int j)
{
  int m = ((2 * j + 1) * k) & 127;
  return m <= 32 ? kCosMag[m] : (m <= 64 ? -kCosMag[64 - m] : (m <= 96 ? -kCosMag[m - 64] : kCosMag[128 - m]));
}
constexpr T32Table make_t32()
{
  T32Table t{};
  for (int k = 0; k < 32; k++)
    for (int j = 0; j < 32; j++) t.v[k][j] = (int8_t)t32_coef(k, j);
  return t;
}
static __constant__ T32Table c_T32 = make_t32();
// 4x4 DST-VII (fastForwardDst/fastInverseDst written as a matrix, TComTrQuant.cpp:443-479)
static __constant__ int8_t c_DST4[4][4] = {{29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29}};
static __constant__ int c_quantScales[6] = {26214, 23302, 20560, 18396, 16384, 14564};   // g_quantScales
static __constant__ int c_invQuantScales[6] = {40, 45, 51, 57, 64, 72};                  // g_invQuantScales

template <int N>
__device__ __forceinline__ int tcoef(int k, int n, bool dst)
{
  if (N == 4 && dst) return c_DST4[k][n];
  return c_T32.v[k * (32 / N)][n];
}

__device__ __forceinline__ int clip16(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }


template <int LOG2> struct TuSmem {
  static constexpr int N = 1 << LOG2;
  static constexpr int P = N + 2;                  // padded row pitch (conflict-free row reads)
  int16_t a[N * P];
  int16_t b[N * P];
  int16_t lev[N * N];
  int16_t du[N * N];
};

// signBitHidingHDQ (TComTrQuant.cpp:977-1100) for ONE coefficient group; `is_last_cg` is the
// reference's lastCG==1 state (the highest-scan-position group holding a non-zero level).
__device__ void sbh_group(int16_t* lev, const int16_t* coef, int coef_pitch_log2, const int16_t* du,
                          const uint16_t* scan, int subPos, bool is_last_cg, int P)
{
  int firstNZ = 16, lastNZ = -1, absSum = 0;
  for (int n = 15; n >= 0; --n) if (lev[scan[n + subPos]]) { lastNZ = n; break; }
  for (int n = 0; n < 16; n++) if (lev[scan[n + subPos]]) { firstNZ = n; break; }
  for (int n = firstNZ; n <= lastNZ; n++) absSum += lev[scan[n + subPos]];
  if (lastNZ - firstNZ < 4) return;                                   // SBH_THRESHOLD
  unsigned signbit = (lev[scan[subPos + firstNZ]] > 0) ? 0u : 1u;
  if (signbit == (unsigned)(absSum & 1)) return;
  int minCostInc = 2147483647, minPos = -1, finalChange = 0, curCost = 2147483647, curChange = 0;
  for (int n = (is_last_cg ? lastNZ : 15); n >= 0; --n) {
    int blkPos = scan[n + subPos];
    int cpos = ((blkPos >> coef_pitch_log2) * P) + (blkPos & ((1 << coef_pitch_log2) - 1));
    if (lev[blkPos] != 0) {
      if (du[blkPos] > 0) { curCost = -du[blkPos]; curChange = 1; }
      else if (n == firstNZ && abs((int)lev[blkPos]) == 1) curCost = 2147483647;
      else { curCost = du[blkPos]; curChange = -1; }
    } else if (n < firstNZ) {
      unsigned thisSign = (coef[cpos] >= 0) ? 0u : 1u;
      if (thisSign != signbit) curCost = 2147483647;
      else { curCost = -du[blkPos]; curChange = 1; }
    } else { curCost = -du[blkPos]; curChange = 1; }
    if (curCost < minCostInc) { minCostInc = curCost; finalChange = curChange; minPos = blkPos; }
  }
  int lv = lev[minPos];
  if (lv == 32767 || lv == -32768) finalChange = -1;
  int cpos = ((minPos >> coef_pitch_log2) * P) + (minPos & ((1 << coef_pitch_log2) - 1));
  // int16 storage is safe: at +-32767/-32768 the reference forces finalChange = -1 (:1078-1081), so
  // the adjusted level never leaves [-32768, 32767].
  lev[minPos] = (int16_t)((coef[cpos] >= 0) ? lv + finalChange : lv - finalChange);
}

// ---------------------------------------------------------------------------------------- forward
template <int LOG2, bool QUANT>
__global__ void __launch_bounds__(128)
k_fwd_tq(PlaneTable pt, int resi_slot, int n_tus, const tvc_tu* __restrict__ tus, tvc_quant_cfg qc, int bd,
         ScanTables scans, int32_t* __restrict__ out_coef, int32_t* __restrict__ out_arl, uint32_t* __restrict__ abs_sum)
{
  constexpr int N = 1 << LOG2, TPW = 32 / N, P = N + 2;
  __shared__ TuSmem<LOG2> sm[4 * TPW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / N, j = lane % N;
  const int tu_idx = (blockIdx.x * 4 + warp) * TPW + sub;
  const bool active = tu_idx < n_tus;
  TuSmem<LOG2>& S = sm[warp * TPW + sub];
  tvc_tu tu;
  if (active) tu = tus[tu_idx]; else { tu = tus[0]; }
  const int bi = bd - 8;
  const int stride = pt.stride[tu.plane];
  const int16_t* resi = pt.org[resi_slot][tu.plane] + (ptrdiff_t)tu.y * stride + tu.x;
  const bool use_dst = (tu.flags & TVC_TU_DST) != 0;
  const bool skip = (tu.flags & TVC_TU_SKIP) != 0;
  const bool bypass = (tu.flags & TVC_TU_BYPASS) != 0;

  // stage the residual block: lanes of the TU sweep rows, coalesced along x
  if (active) {
    for (int i = j; i < N * N; i += N) {
      int r = i >> LOG2, x = i & (N - 1);
      S.a[r * P + x] = resi[(ptrdiff_t)r * stride + x];
    }
  }
  __syncwarp();

  // Every lane runs both 1-D passes (no divergent barriers); transform-skip / bypass TUs then
  // replace the result from the staged block.
  int c[N];   // this lane's N final coefficients: raster index k*N + j, k = 0..N-1
  {
    const int s1 = LOG2 - 1 + bi, s2 = LOG2 + 6;
    int x[N];
#pragma unroll
    for (int n = 0; n < N; n++) x[n] = S.a[j * P + n];
    // first pass: tmp[k][j] = (short)((sum_n T[k][n] * block[j][n] + add) >> s1)
#pragma unroll
    for (int k = 0; k < N; k++) {
      int s = 0;
#pragma unroll
      for (int n = 0; n < N; n++) s += tcoef<N>(k, n, use_dst) * x[n];
      S.b[k * P + j] = (int16_t)((s + (1 << (s1 - 1))) >> s1);
    }
    __syncwarp();
#pragma unroll
    for (int n = 0; n < N; n++) x[n] = S.b[j * P + n];
    // second pass: coeff[k*N + j] = (short)((sum_n T[k][n] * tmp[j][n] + add) >> s2)
#pragma unroll
    for (int k = 0; k < N; k++) {
      int s = 0;
#pragma unroll
      for (int n = 0; n < N; n++) s += tcoef<N>(k, n, use_dst) * x[n];
      c[k] = (int)(int16_t)((s + (1 << (s2 - 1))) >> s2);
    }
  }
  if (skip) {
    // xTransformSkip :1622-1660 (psCoeff[j*height+k] = resi[j][k] << shift)
    const int shift = 15 - bd - LOG2;
#pragma unroll
    for (int k = 0; k < N; k++) {
      int v = S.a[k * P + j];
      c[k] = shift >= 0 ? v * (1 << shift) : ((v + (1 << (-shift - 1))) >> (-shift));
    }
  }
  if (bypass) {
    // transformNxN :1390-1402  coeff = residual, uiAbsSum = sum |resi|, no quantisation
#pragma unroll
    for (int k = 0; k < N; k++) c[k] = S.a[k * P + j];
  }

  if (!QUANT) {
    if (active) {
#pragma unroll
      for (int k = 0; k < N; k++) out_coef[tu.coef_offset + k * N + j] = c[k];
    }
    return;
  }

  // ---- xQuant non-RDOQ (:1185-1258), flat list: level = (|c|*Q[rem] + add) >> qbits with the
  // slice base QP's per (ADAPTIVE_QP_SELECTION), deltaU for sign hiding, ARL side output.
  const int tshift = 15 - bd - LOG2;
  const int qscale = c_quantScales[tu.qp_rem];
  const int qbits = 14 + tu.base_per + tshift;
  const int add = (qc.is_intra_slice ? 171 : 85) << (qbits - 9);
  const int qbitsC = qbits - 7;
  const int addC = 1 << (qbitsC - 1);
  uint32_t acs = 0;
  __syncwarp();
#pragma unroll
  for (int k = 0; k < N; k++) {
    int v = c[k];
    int level, du = 0;
    if (bypass) {
      level = v;
      acs += (uint32_t)abs(v);
    } else {
      long long t = (long long)abs(v) * qscale;
      if (qc.use_arl && out_arl && active) out_arl[tu.coef_offset + k * N + j] = (int)((t + addC) >> qbitsC);
      level = (int)((t + add) >> qbits);
      du = (int)((t - (long long)(int)((unsigned)level << qbits)) >> (qbits - 8));
      acs += (uint32_t)level;
      level = clip16(v < 0 ? -level : level);
    }
    c[k] = level;
    S.lev[k * N + j] = (int16_t)level;
    S.du[k * N + j] = (int16_t)du;
    S.a[k * P + j] = (int16_t)v;              // coefficient sign source for sign hiding
  }
#pragma unroll
  for (int o = N / 2; o > 0; o >>= 1) acs += __shfl_xor_sync(0xffffffffu, acs, o);
  __syncwarp();
  const bool do_sbh = qc.sign_hide && acs >= 2 && !bypass;
  const uint16_t* scan = scans.s[tu.scan_idx][LOG2 - 2];
  constexpr int NCG = N * N / 16;
  // which group is the last one (in scan order) holding a non-zero level
  int last_cg = -1;
  if (do_sbh) {
    for (int g = j; g < NCG; g += N) {
      bool nz = false;
      for (int n = 0; n < 16; n++) nz |= S.lev[scan[g * 16 + n]] != 0;
      if (nz) last_cg = g;
    }
  }
#pragma unroll
  for (int o = N / 2; o > 0; o >>= 1) last_cg = max(last_cg, __shfl_xor_sync(0xffffffffu, last_cg, o));
  __syncwarp();
  if (do_sbh) {
    for (int g = j; g < NCG; g += N) sbh_group(S.lev, S.a, LOG2, S.du, scan, g * 16, g == last_cg, P);
  }
  __syncwarp();
  if (active) {
    if (bypass) {
#pragma unroll
      for (int k = 0; k < N; k++) out_coef[tu.coef_offset + k * N + j] = c[k];    // unclipped residual copy
    } else {
      for (int i = j; i < N * N; i += N) out_coef[tu.coef_offset + i] = S.lev[i];
    }
    if (j == 0 && abs_sum) abs_sum[tu_idx] = acs;
  }
}

// ---------------------------------------------------------------------------------------- inverse
template <int LOG2>
__global__ void __launch_bounds__(128)
k_inv_tq(PlaneTable pt, int resi_slot, int pred_slot, int recon_slot, int n_tus, const tvc_tu* __restrict__ tus,
         int bd, const int32_t* __restrict__ levels, int dequant)
{
  constexpr int N = 1 << LOG2, TPW = 32 / N, P = N + 2;
  __shared__ int16_t sm_a[4 * TPW][N * P];
  __shared__ int16_t sm_b[4 * TPW][N * P];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / N, j = lane % N;
  const int tu_idx = (blockIdx.x * 4 + warp) * TPW + sub;
  const bool active = tu_idx < n_tus;
  int16_t* A = sm_a[warp * TPW + sub];
  int16_t* B = sm_b[warp * TPW + sub];
  tvc_tu tu = tus[active ? tu_idx : 0];
  const int bi = bd - 8;
  const int stride = pt.stride[tu.plane];
  const bool use_dst = (tu.flags & TVC_TU_DST) != 0;
  const bool skip = (tu.flags & TVC_TU_SKIP) != 0;
  const bool bypass = (tu.flags & TVC_TU_BYPASS) != 0;
  const int32_t* q = levels + tu.coef_offset;

  // lane j reads column j of the level raster (coalesced), dequantises (xDeQuant flat :1343-1353)
  const int tshift = 15 - bd - LOG2;
  const int dshift = 20 - 14 - tshift;
  const int dadd = 1 << (dshift - 1);
  const int dscale = c_invQuantScales[tu.qp_rem] << tu.qp_per;
  int x[N];
#pragma unroll
  for (int k = 0; k < N; k++) {
    int v = active ? q[k * N + j] : 0;
    if (dequant && !bypass) {
      int cq = clip16(v);
      v = clip16((int)((unsigned)cq * (unsigned)dscale + (unsigned)dadd) >> dshift);
    }
    x[k] = v;
  }
  // Every lane runs both passes (no divergent barriers); skip / bypass TUs overwrite the block
  // afterwards.  A[r*P + col] holds residual (row r, col) for the copy-out.
  int xs[N];
#pragma unroll
  for (int k = 0; k < N; k++) xs[k] = x[k];
  // xIT :1599-1603 casts the Int coefficients to short first
#pragma unroll
  for (int k = 0; k < N; k++) x[k] = (int)(int16_t)x[k];
  // first pass (shift 7): tmp[j][n] = Clip16((sum_k T[k][n] * coef[k][j] + 64) >> 7)
#pragma unroll
  for (int n = 0; n < N; n++) {
    int s = 0;
#pragma unroll
    for (int k = 0; k < N; k++) s += tcoef<N>(k, n, use_dst) * x[k];
    B[j * P + n] = (int16_t)clip16((s + 64) >> 7);
  }
  __syncwarp();
  // second pass (shift 12-bi): block[j][n] = Clip16((sum_k T[k][n] * tmp[k][j] + add) >> s2)
  const int s2 = 12 - bi;
#pragma unroll
  for (int k = 0; k < N; k++) x[k] = B[k * P + j];
  if (!skip && !bypass) {
#pragma unroll
    for (int n = 0; n < N; n++) {
      int s = 0;
#pragma unroll
      for (int k = 0; k < N; k++) s += tcoef<N>(k, n, use_dst) * x[k];
      A[j * P + n] = (int16_t)clip16((s + (1 << (s2 - 1))) >> s2);
    }
  } else if (bypass) {
    // invtransformNxN :1430-1440  rpcResidual[k*stride+j] = pcCoeff[k*w+j]
#pragma unroll
    for (int k = 0; k < N; k++) A[k * P + j] = (int16_t)xs[k];
  } else {
    // xITransformSkip :1668-1704
    const int shift = tshift;
#pragma unroll
    for (int k = 0; k < N; k++)
      A[k * P + j] = (int16_t)(shift > 0 ? ((xs[k] + (1 << (shift - 1))) >> shift) : (xs[k] * (1 << (-shift))));
  }
  __syncwarp();
  if (!active) return;
  int16_t* resi = pt.org[resi_slot][tu.plane] + (ptrdiff_t)tu.y * stride + tu.x;
  const int maxv = (1 << bd) - 1;
  for (int i = j; i < N * N; i += N) {
    int r = i >> LOG2, xx = i & (N - 1);
    int v = A[r * P + xx];
    resi[(ptrdiff_t)r * stride + xx] = (int16_t)v;
    if (pred_slot >= 0) {
      ptrdiff_t o = (ptrdiff_t)(tu.y + r) * stride + tu.x + xx;
      int rec = pt.org[pred_slot][tu.plane][o] + v;          // TComYuv::addClip (TComYuv.cpp:407-429)
      rec = rec < 0 ? 0 : (rec > maxv ? maxv : rec);
      pt.org[recon_slot][tu.plane][o] = (int16_t)rec;
    }
  }
}

// dequant only (xDeQuant drop-in)
__global__ void k_dequant(const int32_t* __restrict__ q, int32_t* __restrict__ out, int n, int log2, int per, int rem, int bd)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int tshift = 15 - bd - log2, shift = 6 - tshift, add = 1 << (shift - 1);
  int scale = c_invQuantScales[rem] << per;
  int cq = clip16(q[i]);
  out[i] = clip16((int)((unsigned)cq * (unsigned)scale + (unsigned)add) >> shift);
}

// ---- host side ---------------------------------------------------------------------------------
static uint16_t* g_scan_dev[3][4] = {};
static int g_scan_device = -1;

static void build_scan(int scan_idx, int log2, std::vector<uint16_t>& out)
{
  // coding scans of initSigLastScan (TComRom.cpp:564-690): 4x4 groups visited along up-right
  // diagonals (or row/column-wise), up-right diagonal (or row / column) inside a group
  int n = 1 << log2;
  out.resize((size_t)n * n);
  auto diag = [](int m, std::vector<int>& ys, std::vector<int>& xs) {
    ys.clear(); xs.clear();
    for (int line = 0; (int)ys.size() < m * m; line++) {
      int prim = line, scnd = 0;
      while (prim >= m) { scnd++; prim--; }
      while (prim >= 0 && scnd < m) { ys.push_back(prim); xs.push_back(scnd); scnd++; prim--; }
    }
  };
  int cnt = 0;
  if (scan_idx == 0) {
    std::vector<int> iy, ix, gy, gx;
    diag(4, iy, ix);
    if (n == 4) { for (int i = 0; i < 16; i++) out[i] = (uint16_t)(iy[i] * 4 + ix[i]); return; }
    int nb = n >> 2;
    diag(nb, gy, gx);
    for (int b = 0; b < nb * nb; b++)
      for (int i = 0; i < 16; i++) out[16 * b + i] = (uint16_t)(iy[i] * n + ix[i] + 4 * (gx[b] + gy[b] * n));
  } else if (scan_idx == 1) {
    int nb = n >> 2;
    for (int by = 0; by < nb; by++) for (int bx = 0; bx < nb; bx++)
      for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) out[cnt++] = (uint16_t)((by * 4 + y) * n + bx * 4 + x);
  } else {
    int nb = n >> 2;
    for (int bx = 0; bx < nb; bx++) for (int by = 0; by < nb; by++)
      for (int x = 0; x < 4; x++) for (int y = 0; y < 4; y++) out[cnt++] = (uint16_t)((by * 4 + y) * n + bx * 4 + x);
  }
}

int ensure_scans(tvc_ctx* c, ScanTables& st)
{
  if (g_scan_device != c->cfg.device) {
    for (int s = 0; s < 3; s++)
      for (int l = 2; l <= 5; l++) {
        std::vector<uint16_t> v;
        build_scan(s, l, v);
        uint16_t* d = nullptr;
        TVC_CUDA(c, cudaMalloc(&d, v.size() * 2));
        TVC_CUDA(c, cudaMemcpy(d, v.data(), v.size() * 2, cudaMemcpyHostToDevice));
        g_scan_dev[s][l - 2] = d;
      }
    g_scan_device = c->cfg.device;
  }
  for (int s = 0; s < 3; s++) for (int l = 0; l < 4; l++) st.s[s][l] = g_scan_dev[s][l];
  return TVC_OK;
}

int validate_tus(tvc_ctx* c, int plane_slot, int n, const tvc_tu* tus, size_t coef_elems, int counts[4])
{
  const Pic& p = c->pics[plane_slot];
  counts[0] = counts[1] = counts[2] = counts[3] = 0;
  int prev = 2;
  for (int i = 0; i < n; i++) {
    const tvc_tu& t = tus[i];
    if (!tu_record_ok(p, t, coef_elems))
      return set_err(c, TVC_ERR_ARG, "TU %d invalid", i);
    if (t.log2_size < prev) return set_err(c, TVC_ERR_ARG, "TU list must be grouped by ascending log2_size (TU %d)", i);
    prev = t.log2_size;
    counts[t.log2_size - 2]++;
  }
  return TVC_OK;
}

template <bool QUANT>
static int launch_fwd(tvc_ctx* c, int resi_slot, const int counts[4], const tvc_tu* tus_dev, const tvc_quant_cfg& qc,
                      int32_t* coef_dev, int32_t* arl_dev, uint32_t* abs_dev)
{
  ScanTables st;
  int r = ensure_scans(c, st);
  if (r) return r;
  int off = 0, bd = c->cfg.bit_depth;
  ProfScope ps(c, TVC_PH_FWD_TQ);
#define TVC_FWD(L)                                                                                              \
  if (counts[L - 2] > 0) {                                                                                      \
    int per_cta = 4 * (32 >> L), n = counts[L - 2];                                                             \
    k_fwd_tq<L, QUANT><<<(n + per_cta - 1) / per_cta, 128, 0, c->stream>>>(c->planes, resi_slot, n, tus_dev + off, qc, bd, st, \
                                                                coef_dev, arl_dev, abs_dev ? abs_dev + off : nullptr); \
    TVC_LAUNCH_CHECK(c);                                                                                        \
    off += n;                                                                                                   \
  }
  TVC_FWD(2) TVC_FWD(3) TVC_FWD(4) TVC_FWD(5)
#undef TVC_FWD
  return TVC_OK;
}

int launch_inv(tvc_ctx* c, int resi_slot, int pred_slot, int recon_slot, const int counts[4], const tvc_tu* tus_dev,
               const int32_t* levels_dev, int dequant)
{
  int off = 0, bd = c->cfg.bit_depth;
  ProfScope ps(c, TVC_PH_INV_TQ);
#define TVC_INV(L)                                                                                              \
  if (counts[L - 2] > 0) {                                                                                      \
    int per_cta = 4 * (32 >> L), n = counts[L - 2];                                                             \
    k_inv_tq<L><<<(n + per_cta - 1) / per_cta, 128, 0, c->stream>>>(c->planes, resi_slot, pred_slot, recon_slot, n, \
                                                                    tus_dev + off, bd, levels_dev, dequant);   \
    TVC_LAUNCH_CHECK(c);                                                                                        \
    off += n;                                                                                                   \
  }
  TVC_INV(2) TVC_INV(3) TVC_INV(4) TVC_INV(5)
#undef TVC_INV
  return TVC_OK;
}

}  // namespace tvc

using namespace tvc;

extern "C" {

// device-resident variants: the TU list must be grouped by ascending log2_size and the caller
// passes the per-size counts (no copies, no synchronisation here).
static int check_counts(int n, const int32_t* counts)
{
  if (!counts) return TVC_ERR_ARG;
  long long t = 0;
  for (int i = 0; i < 4; i++) { if (counts[i] < 0) return TVC_ERR_ARG; t += counts[i]; }
  return t == n ? TVC_OK : TVC_ERR_ARG;
}

int tvc_fwd_tq_batch_dev(tvc_ctx* c, int resi_slot, int n, const tvc_tu* tus_dev, const int32_t* counts,
                         const tvc_quant_cfg* qc, int32_t* levels_dev, int32_t* arl_dev, uint32_t* abs_sum_dev)
{
  if (!c || !valid_slot(c, resi_slot) || n < 0 || !qc || (n && (!tus_dev || !levels_dev)) || check_counts(n, counts))
    return set_err(c, TVC_ERR_ARG, "tvc_fwd_tq_batch_dev: bad argument");
  if (n == 0) return TVC_OK;
  int cn[4] = {counts[0], counts[1], counts[2], counts[3]};
  return launch_fwd<true>(c, resi_slot, cn, tus_dev, *qc, levels_dev, arl_dev, abs_sum_dev);
}

int tvc_fwd_transform_batch_dev(tvc_ctx* c, int resi_slot, int n, const tvc_tu* tus_dev, const int32_t* counts, int32_t* coef_dev)
{
  if (!c || !valid_slot(c, resi_slot) || n < 0 || (n && (!tus_dev || !coef_dev)) || check_counts(n, counts))
    return set_err(c, TVC_ERR_ARG, "tvc_fwd_transform_batch_dev: bad argument");
  if (n == 0) return TVC_OK;
  int cn[4] = {counts[0], counts[1], counts[2], counts[3]};
  tvc_quant_cfg q0 = {0, 0, 0};
  return launch_fwd<false>(c, resi_slot, cn, tus_dev, q0, coef_dev, nullptr, nullptr);
}

int tvc_inv_tq_batch_dev(tvc_ctx* c, int resi_slot, int pred_slot, int recon_slot, int n, const tvc_tu* tus_dev,
                         const int32_t* counts, const int32_t* levels_dev)
{
  if (!c || !valid_slot(c, resi_slot) || n < 0 || (n && (!tus_dev || !levels_dev)) || check_counts(n, counts) ||
      (pred_slot >= 0 && (!valid_slot(c, pred_slot) || !valid_slot(c, recon_slot))))
    return set_err(c, TVC_ERR_ARG, "tvc_inv_tq_batch_dev: bad argument");
  if (n == 0) return TVC_OK;
  int cn[4] = {counts[0], counts[1], counts[2], counts[3]};
  return launch_inv(c, resi_slot, pred_slot, recon_slot, cn, tus_dev, levels_dev, 1);
}

static int fwd_host(tvc_ctx* c, bool quant, int resi_slot, int n, const tvc_tu* tus, const tvc_quant_cfg* qc,
                    int32_t* coef, int32_t* arl, size_t coef_elems, uint32_t* abs_sum)
{
  if (!c || !valid_slot(c, resi_slot) || n < 0 || (n && (!tus || !coef))) return set_err(c, TVC_ERR_ARG, "fwd tq: bad argument");
  if (n == 0) return TVC_OK;
  int counts[4], r;
  if ((r = validate_tus(c, resi_slot, n, tus, coef_elems, counts))) return r;
  size_t tu_bytes = ((size_t)n * sizeof(tvc_tu) + 255) & ~(size_t)255;
  size_t coef_bytes = ((coef_elems * 4) + 255) & ~(size_t)255;
  size_t abs_bytes = (((size_t)n * 4) + 255) & ~(size_t)255;
  const bool want_arl = quant && arl && qc->use_arl;
  // page-locked caller buffers are copied directly; pageable ones go through the pinned staging area
  const bool pin_tus = is_pinned(tus), pin_out = is_pinned(coef) && (!want_arl || is_pinned(arl)) && (!abs_sum || is_pinned(abs_sum));
  if ((r = ensure_scratch(c, c->in, tu_bytes))) return r;
  if ((r = ensure_scratch(c, c->out, 2 * coef_bytes + abs_bytes))) return r;
  const void* tu_src = tus;
  if (!pin_tus) { memcpy(c->in.host, tus, (size_t)n * sizeof(tvc_tu)); tu_src = c->in.host; }
  TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, tu_src, (size_t)n * sizeof(tvc_tu), cudaMemcpyHostToDevice, c->stream));
  // device layout: [coefficients][abs sums][ARL coefficients]; only what was produced is copied back
  int32_t* d_coef = (int32_t*)c->out.dev;
  uint32_t* d_abs = (uint32_t*)((char*)c->out.dev + coef_bytes);
  int32_t* d_arl = (int32_t*)((char*)c->out.dev + coef_bytes + abs_bytes);
  tvc_quant_cfg q0 = {0, 0, 0};
  if (quant) r = launch_fwd<true>(c, resi_slot, counts, (const tvc_tu*)c->in.dev, *qc, d_coef, want_arl ? d_arl : nullptr, d_abs);
  else r = launch_fwd<false>(c, resi_slot, counts, (const tvc_tu*)c->in.dev, q0, d_coef, nullptr, nullptr);
  if (r) return r;
  if (pin_out) {
    TVC_CUDA(c, cudaMemcpyAsync(coef, d_coef, coef_elems * 4, cudaMemcpyDeviceToHost, c->stream));
    if (want_arl) TVC_CUDA(c, cudaMemcpyAsync(arl, d_arl, coef_elems * 4, cudaMemcpyDeviceToHost, c->stream));
    if (quant && abs_sum) TVC_CUDA(c, cudaMemcpyAsync(abs_sum, d_abs, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    TVC_CUDA(c, cudaStreamSynchronize(c->stream));
    return TVC_OK;
  }
  size_t back = coef_bytes + (quant ? abs_bytes : 0) + (want_arl ? coef_bytes : 0);
  TVC_CUDA(c, cudaMemcpyAsync(c->out.host, c->out.dev, back, cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(coef, c->out.host, coef_elems * 4);
  if (want_arl) memcpy(arl, (char*)c->out.host + coef_bytes + abs_bytes, coef_elems * 4);
  if (quant && abs_sum) memcpy(abs_sum, (char*)c->out.host + coef_bytes, (size_t)n * 4);
  return TVC_OK;
}

int tvc_fwd_transform_batch(tvc_ctx* c, int resi_slot, int n, const tvc_tu* tus, int32_t* coef, size_t coef_elems)
{
  return fwd_host(c, false, resi_slot, n, tus, nullptr, coef, nullptr, coef_elems, nullptr);
}

int tvc_fwd_tq_batch(tvc_ctx* c, int resi_slot, int n, const tvc_tu* tus, const tvc_quant_cfg* qc, int32_t* levels,
                     int32_t* arl, size_t coef_elems, uint32_t* abs_sum)
{
  if (!qc) return set_err(c, TVC_ERR_ARG, "tvc_fwd_tq_batch: null quant cfg");
  return fwd_host(c, true, resi_slot, n, tus, qc, levels, arl, coef_elems, abs_sum);
}

int tvc_inv_tq_batch(tvc_ctx* c, int resi_slot, int pred_slot, int recon_slot, int n, const tvc_tu* tus,
                     const int32_t* levels, size_t coef_elems)
{
  if (!c || !valid_slot(c, resi_slot) || n < 0 || (n && (!tus || !levels)) ||
      (pred_slot >= 0 && (!valid_slot(c, pred_slot) || !valid_slot(c, recon_slot))))
    return set_err(c, TVC_ERR_ARG, "tvc_inv_tq_batch: bad argument");
  if (n == 0) return TVC_OK;
  int counts[4], r;
  if ((r = validate_tus(c, resi_slot, n, tus, coef_elems, counts))) return r;
  size_t tu_bytes = ((size_t)n * sizeof(tvc_tu) + 255) & ~(size_t)255;
  if ((r = ensure_scratch(c, c->in, tu_bytes + coef_elems * 4))) return r;
  const bool pin_tus = is_pinned(tus), pin_lev = is_pinned(levels);
  const void* tu_src = tus;
  const void* lev_src = levels;
  if (!pin_tus) { memcpy(c->in.host, tus, (size_t)n * sizeof(tvc_tu)); tu_src = c->in.host; }
  if (!pin_lev) { memcpy((char*)c->in.host + tu_bytes, levels, coef_elems * 4); lev_src = (char*)c->in.host + tu_bytes; }
  TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, tu_src, (size_t)n * sizeof(tvc_tu), cudaMemcpyHostToDevice, c->stream));
  TVC_CUDA(c, cudaMemcpyAsync((char*)c->in.dev + tu_bytes, lev_src, coef_elems * 4, cudaMemcpyHostToDevice, c->stream));
  if ((r = launch_inv(c, resi_slot, pred_slot, recon_slot, counts, (const tvc_tu*)c->in.dev,
                      (const int32_t*)((char*)c->in.dev + tu_bytes), 1)))
    return r;
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  return TVC_OK;
}

// ---- single-TU drop-ins on host blocks (TComTrQuant::xT / xIT / xDeQuant) ------------------------
// They run through the same kernels, using picture slot 0's luma plane is NOT touched: a private
// staging plane inside the scratch buffer is addressed through a one-entry plane table.
static int single_tu(tvc_ctx* c, bool forward, int use_dst, const int16_t* resi_in, int16_t* resi_out, int stride,
                     const int32_t* coef_in, int32_t* coef_out, int w, int h)
{
  if (!c || w != h || (w != 4 && w != 8 && w != 16 && w != 32)) return set_err(c, TVC_ERR_ARG, "tvc_xT/xIT: unsupported size");
  int log2 = w == 4 ? 2 : w == 8 ? 3 : w == 16 ? 4 : 5;
  size_t nn = (size_t)w * h;
  size_t tu_bytes = 256, plane_bytes = ((nn * 2) + 255) & ~(size_t)255, coef_bytes = nn * 4;
  int r;
  if ((r = ensure_scratch(c, c->in, tu_bytes + plane_bytes + coef_bytes))) return r;
  if ((r = ensure_scratch(c, c->out, plane_bytes + coef_bytes))) return r;
  tvc_tu tu;
  memset(&tu, 0, sizeof(tu));
  tu.log2_size = log2;
  tu.flags = (use_dst && w == 4) ? TVC_TU_DST : 0;
  char* hp = (char*)c->in.host;
  memcpy(hp, &tu, sizeof(tu));
  PlaneTable pt;
  memset(&pt, 0, sizeof(pt));
  pt.stride[0] = w;
  int counts[4] = {0, 0, 0, 0};
  counts[log2 - 2] = 1;
  if (forward) {
    int16_t* hb = (int16_t*)(hp + tu_bytes);
    for (int y = 0; y < h; y++) memcpy(hb + (size_t)y * w, resi_in + (ptrdiff_t)y * stride, (size_t)w * 2);
    TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, hp, tu_bytes + plane_bytes, cudaMemcpyHostToDevice, c->stream));
    pt.org[0][0] = (int16_t*)((char*)c->in.dev + tu_bytes);
    PlaneTable saved = c->planes;
    c->planes = pt;
    tvc_quant_cfg q0 = {0, 0, 0};
    r = launch_fwd<false>(c, 0, counts, (const tvc_tu*)c->in.dev, q0, (int32_t*)c->out.dev, nullptr, nullptr);
    c->planes = saved;
    if (r) return r;
    TVC_CUDA(c, cudaMemcpyAsync(c->out.host, c->out.dev, coef_bytes, cudaMemcpyDeviceToHost, c->stream));
    TVC_CUDA(c, cudaStreamSynchronize(c->stream));
    memcpy(coef_out, c->out.host, coef_bytes);
  } else {
    memcpy(hp + tu_bytes, coef_in, coef_bytes);
    TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, hp, tu_bytes + coef_bytes, cudaMemcpyHostToDevice, c->stream));
    pt.org[0][0] = (int16_t*)c->out.dev;
    PlaneTable saved = c->planes;
    c->planes = pt;
    r = launch_inv(c, 0, -1, -1, counts, (const tvc_tu*)c->in.dev, (const int32_t*)((char*)c->in.dev + tu_bytes), 0);
    c->planes = saved;
    if (r) return r;
    TVC_CUDA(c, cudaMemcpyAsync(c->out.host, c->out.dev, nn * 2, cudaMemcpyDeviceToHost, c->stream));
    TVC_CUDA(c, cudaStreamSynchronize(c->stream));
    const int16_t* ob = (const int16_t*)c->out.host;
    for (int y = 0; y < h; y++) memcpy(resi_out + (ptrdiff_t)y * stride, ob + (size_t)y * w, (size_t)w * 2);
  }
  return TVC_OK;
}

int tvc_xT(tvc_ctx* c, int use_dst, const int16_t* resi, int stride, int32_t* coef, int w, int h)
{
  if (!resi || !coef) return set_err(c, TVC_ERR_ARG, "tvc_xT: null pointer");
  return single_tu(c, true, use_dst, resi, nullptr, stride, nullptr, coef, w, h);
}

int tvc_xIT(tvc_ctx* c, int use_dst, const int32_t* coef, int16_t* resi, int stride, int w, int h)
{
  if (!resi || !coef) return set_err(c, TVC_ERR_ARG, "tvc_xIT: null pointer");
  return single_tu(c, false, use_dst, nullptr, resi, stride, coef, nullptr, w, h);
}

int tvc_xDeQuant(tvc_ctx* c, const int32_t* qcoef, int32_t* coef, int w, int h, int per, int rem)
{
  if (!c || !qcoef || !coef || w != h || (w != 4 && w != 8 && w != 16 && w != 32) || rem < 0 || rem > 5 || per < 0 || per > 12)
    return set_err(c, TVC_ERR_ARG, "tvc_xDeQuant: bad argument");
  int log2 = w == 4 ? 2 : w == 8 ? 3 : w == 16 ? 4 : 5;
  size_t nn = (size_t)w * h;
  int r;
  if ((r = ensure_scratch(c, c->in, nn * 4))) return r;
  if ((r = ensure_scratch(c, c->out, nn * 4))) return r;
  memcpy(c->in.host, qcoef, nn * 4);
  TVC_CUDA(c, cudaMemcpyAsync(c->in.dev, c->in.host, nn * 4, cudaMemcpyHostToDevice, c->stream));
  k_dequant<<<(int)((nn + 255) / 256), 256, 0, c->stream>>>((const int32_t*)c->in.dev, (int32_t*)c->out.dev, (int)nn, log2, per, rem, c->cfg.bit_depth);
  TVC_LAUNCH_CHECK(c);
  TVC_CUDA(c, cudaMemcpyAsync(c->out.host, c->out.dev, nn * 4, cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(coef, c->out.host, nn * 4);
  return TVC_OK;
}

}  // extern "C"
