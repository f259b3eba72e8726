// tvc_dist.cuh -- device-side Hadamard / reduction helpers of the distortion functions (TComRdCost.cpp:1663-1872),
// shared by the distortion drop-ins (tvc_dist.cu) and the prediction-cost kernel (tvc_interp.cu).
#pragma once
#include <stdint.h>

namespace tvc {

__device__ __forceinline__ uint32_t warp_sum(uint32_t v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// in-register 1-D Hadamard butterflies over a strided view of d[]
template <int N, int STRIDE>
__device__ __forceinline__ void hadamard1d(int* d)
{
#pragma unroll
  for (int len = 1; len < N; len <<= 1) {
#pragma unroll
    for (int i = 0; i < N; i += 2 * len) {
#pragma unroll
      for (int j = i; j < i + len; j++) {
        int a = d[j * STRIDE], b = d[(j + len) * STRIDE];
        d[j * STRIDE] = a + b;
        d[(j + len) * STRIDE] = a - b;
      }
    }
  }
}

// xCalcHADs8x8 / 4x4 / 2x2 (TComRdCost.cpp:1663-1872): full 2-D Hadamard, sum of magnitudes,
// per-tile rounding (sum+2)>>2 for 8x8, (sum+1)>>1 for 4x4, none for 2x2.
template <int N>
__device__ __forceinline__ uint32_t had_tile(const int16_t* __restrict__ o, int so, const int16_t* __restrict__ c, int sc)
{
  int d[N * N];
#pragma unroll
  for (int y = 0; y < N; y++)
#pragma unroll
    for (int x = 0; x < N; x++) d[y * N + x] = (int)o[y * so + x] - (int)c[y * sc + x];
#pragma unroll
  for (int y = 0; y < N; y++) hadamard1d<N, 1>(d + y * N);
#pragma unroll
  for (int x = 0; x < N; x++) hadamard1d<N, N>(d + x);
  int s = 0;
#pragma unroll
  for (int k = 0; k < N * N; k++) s += abs(d[k]);
  if (N == 8) s = (s + 2) >> 2;
  else if (N == 4) s = (s + 1) >> 1;
  return (uint32_t)s;
}

}  // namespace tvc
