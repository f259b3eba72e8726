// tvc_interp.cuh -- device-side sample arithmetic of TComInterpolationFilter
// (TComInterpolationFilter.cpp:55-244).  Shared by the filter drop-ins, MC and fractional ME.
#pragma once
#include <stdint.h>

namespace tvc {

// HEVC interpolation taps (H.265 8.5.3.3.3; reference: TComInterpolationFilter.cpp:55-73)
static __constant__ __align__(8) int8_t c_luma_taps[4][8] = {
  {  0, 0,   0, 64,  0,   0, 0,  0 },
  { -1, 4, -10, 58, 17,  -5, 1,  0 },
  { -1, 4, -11, 40, 40, -11, 4, -1 },
  {  0, 1,  -5, 17, 58, -10, 4, -1 }
};
static __constant__ int8_t c_chroma_taps[8][4] = {
  {  0, 64,  0,  0 }, { -2, 58, 10, -2 }, { -4, 54, 16, -2 }, { -6, 46, 28, -4 },
  { -4, 36, 36, -4 }, { -4, 28, 46, -6 }, { -2, 16, 54, -4 }, { -2, 10, 58, -2 }
};

constexpr int kIfPrec = 14;            // IF_INTERNAL_PREC
constexpr int kIfFilt = 6;             // IF_FILTER_PREC
constexpr int kIfOffs = 1 << (kIfPrec - 1);   // IF_INTERNAL_OFFS = 8192

// filterCopy (TComInterpolationFilter.cpp:91-145) for one sample
__device__ __forceinline__ int16_t if_copy(int v, bool isFirst, bool isLast, int bd)
{
  if (isFirst == isLast) return (int16_t)v;
  int shift = kIfPrec - bd;
  if (isFirst) {
    int16_t val = (int16_t)(v << shift);
    return (int16_t)(val - (int16_t)kIfOffs);
  }
  int16_t offset = (int16_t)(kIfOffs + (shift ? (1 << (shift - 1)) : 0));
  int16_t val = (int16_t)(((int)(int16_t)v + offset) >> shift);
  int16_t maxv = (int16_t)((1 << bd) - 1);
  if (val < 0) val = 0;
  if (val > maxv) val = maxv;
  return val;
}

// rounding of filter<N,..> (TComInterpolationFilter.cpp:190-238): (Short)((sum+offset)>>shift), clip if last
__device__ __forceinline__ int16_t if_round(int sum, bool isFirst, bool isLast, int bd)
{
  int headRoom = kIfPrec - bd;
  int shift = kIfFilt, offset;
  if (isLast) {
    shift += isFirst ? 0 : headRoom;
    offset = 1 << (shift - 1);
    offset += isFirst ? 0 : (kIfOffs << kIfFilt);
  } else {
    shift -= isFirst ? headRoom : 0;
    offset = isFirst ? -(kIfOffs << shift) : 0;
  }
  int16_t val = (int16_t)((sum + offset) >> shift);
  if (isLast) {
    int16_t maxv = (int16_t)((1 << bd) - 1);
    if (val < 0) val = 0;
    if (val > maxv) val = maxv;
  }
  return val;
}

// one output sample of filterHor/VerLuma/Chroma.  p points at the output's source position,
// cs is the element step along the filter direction.
template <int N>
__device__ __forceinline__ int16_t if_sample(const int16_t* __restrict__ p, int cs, int frac, bool isFirst, bool isLast, int bd)
{
  if (frac == 0) return if_copy(p[0], isFirst, isLast, bd);
  const int8_t* c = (N == 8) ? c_luma_taps[frac] : c_chroma_taps[frac];
  p -= (N / 2 - 1) * cs;
  int sum = 0;
#pragma unroll
  for (int t = 0; t < N; t++) sum += (int)p[t * cs] * (int)c[t];
  return if_round(sum, isFirst, isLast, bd);
}

}  // namespace tvc
