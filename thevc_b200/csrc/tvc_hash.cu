// tvc_hash.cu -- picture hashes and PSNR sums on device-resident pictures.
//
// Reference: TLibCommon/TComPicYuvMD5.cpp (calcMD5 :175-200 with md5_plane :66-86, calcCRC :119-134 with compCRC :88-117,
// calcChecksum :167-173 with compChecksum :136-165) as called for the decoded-picture-hash SEI (TLibEncoder/TEncGOP.cpp:1150-1172,
// TLibDecoder/TDecGop.cpp:340-370), and the three sums of TEncGOP::xCalculateAddPSNR (TLibEncoder/TEncGOP.cpp:1582-1641).
//
// checksum and SSD are plain reductions.  The CRC is linear over GF(2): one thread per row computes the row's CRC from a zero state,
// a single thread then folds the rows in order, state <- state * x^(bits per row) mod P xor row CRC (P = x^16 + 0x1021), which is
// the reference's bit-serial walk regrouped.  MD5 is a serial chain by construction: one thread per plane walks it (three planes in
// parallel); it is here so that a reconstruction that never leaves the device can still be hashed, not because it is fast.
#include "tvc_internal.cuh"

namespace tvc {

__global__ void k_checksum_plane(const int16_t* __restrict__ p, int stride, int w, int h, int bd, uint32_t* __restrict__ out)
{
  uint32_t acc = 0;
  const int n = w * h;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int y = i / w, x = i - y * w;
    const uint32_t mask = (uint32_t)((x & 0xff) ^ (y & 0xff) ^ (x >> 8) ^ (y >> 8)) & 0xffu;     // unsigned char xor_mask (:141,147)
    const int v = p[(ptrdiff_t)y * stride + x];
    acc += (uint32_t)((v & 0xff) ^ (int)mask);
    if (bd > 8) acc += (uint32_t)((v >> 8) ^ (int)mask);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

__global__ void k_ssd_plane(const int16_t* __restrict__ a, const int16_t* __restrict__ b, int stride, int w, int h,
                            unsigned long long* __restrict__ out)
{
  unsigned long long acc = 0;
  const int n = w * h;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int y = i / w, x = i - y * w;
    const int d = (int)a[(ptrdiff_t)y * stride + x] - (int)b[(ptrdiff_t)y * stride + x];
    acc += (unsigned long long)(long long)(d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

__device__ __forceinline__ uint32_t crc_step(uint32_t crc, uint32_t bit)
{
  return (((crc << 1) + bit) & 0xffffu) ^ (((crc >> 15) & 1u) * 0x1021u);
}

// CRC of every row from a zero state (compCRC's inner loops, :98-108)
__global__ void k_crc_rows(const int16_t* __restrict__ p, int stride, int w, int h, int bd, uint16_t* __restrict__ row_crc)
{
  const int y = blockIdx.x * blockDim.x + threadIdx.x;
  if (y >= h) return;
  const uint32_t msb = (uint32_t)bd - 1;
  uint32_t crc = 0;
  const int16_t* r = p + (ptrdiff_t)y * stride;
  for (int x = 0; x < w; x++) {
    const uint32_t v = (uint32_t)(int)r[x];
    for (uint32_t b = 0; b < (uint32_t)bd; b++) crc = crc_step(crc, (v >> (msb - (b & msb))) & 1u);      // bit positions as written (:104)
  }
  row_crc[y] = (uint16_t)crc;
}

__global__ void k_crc_fold(const uint16_t* __restrict__ row_crc, int w, int h, int bd, uint8_t* __restrict__ digest)
{
  if (threadIdx.x || blockIdx.x) return;
  uint32_t xl = 1;                                   // x^(w * bd) mod P
  for (int i = 0; i < w * bd; i++) xl = crc_step(xl, 0);
  uint32_t s = 0xffff;
  for (int y = 0; y < h; y++) {
    uint32_t r = 0;                                  // s * xl mod P
    for (int i = 15; i >= 0; i--) { r = crc_step(r, 0); if ((xl >> i) & 1u) r ^= s; }
    s = r ^ row_crc[y];
  }
  for (int i = 0; i < 16; i++) s = crc_step(s, 0);   // the 16 flushing bits (:109-113)
  digest[0] = (uint8_t)(s >> 8);
  digest[1] = (uint8_t)s;
}

// ---- MD5 (RFC 1321), one thread per plane
__constant__ uint32_t c_md5_k[64] = {
  0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1,
  0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453,
  0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942,
  0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05,
  0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d,
  0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};

struct Md5 { uint32_t a, b, c, d; };

__device__ __forceinline__ uint32_t rotl(uint32_t v, int s) { return (v << s) | (v >> (32 - s)); }

__device__ void md5_block(Md5& m, const uint32_t (&w)[16])
{
  uint32_t a = m.a, b = m.b, c = m.c, d = m.d;
#pragma unroll
  for (int i = 0; i < 64; i++) {
    uint32_t f; int g, s;
    if (i < 16) { f = (b & c) | (~b & d); g = i; s = (i & 3) == 0 ? 7 : (i & 3) == 1 ? 12 : (i & 3) == 2 ? 17 : 22; }
    else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; s = (i & 3) == 0 ? 5 : (i & 3) == 1 ? 9 : (i & 3) == 2 ? 14 : 20; }
    else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; s = (i & 3) == 0 ? 4 : (i & 3) == 1 ? 11 : (i & 3) == 2 ? 16 : 23; }
    else { f = c ^ (b | ~d); g = (7 * i) & 15; s = (i & 3) == 0 ? 6 : (i & 3) == 1 ? 10 : (i & 3) == 2 ? 15 : 21; }
    const uint32_t t = a + f + c_md5_k[i] + w[g];
    a = d; d = c; c = b;
    b = b + rotl(t, s);
  }
  m.a += a; m.b += b; m.c += c; m.d += d;
}

struct MdPlane { const int16_t* p; int stride, w, h; };

// samples packed little endian, 1 byte for bit depth <= 8 else 2, raster order (md5_plane / md5_block, :46-86).  One warp per plane:
// the 32 lanes pack 4 KB of the byte stream into shared memory, lane 0 walks the 64 MD5 blocks of the chunk out of it (message words
// at compile-time indices, so they live in registers), and so on; padding and the bit length are appended in shared memory.
constexpr int kMd5ChunkWords = 1024;
__global__ void __launch_bounds__(32) k_md5_planes(MdPlane p0, MdPlane p1, MdPlane p2, int bd, uint8_t* __restrict__ digest /* 3 x 16 */)
{
  __shared__ uint32_t sbuf[kMd5ChunkWords + 32];
  const MdPlane P = blockIdx.x == 0 ? p0 : (blockIdx.x == 1 ? p1 : p2);
  const int lane = threadIdx.x;
  const int bps = bd > 8 ? 2 : 1;
  const unsigned long long total = (unsigned long long)P.w * P.h * bps;
  Md5 m = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
  for (unsigned long long base = 0; base < total || base == 0; base += 4ull * kMd5ChunkWords) {
    const unsigned long long left = total - base;
    const unsigned n = left < 4ull * kMd5ChunkWords ? (unsigned)left : 4u * kMd5ChunkWords;
    for (int k = lane; k < kMd5ChunkWords + 32; k += 32) {
      uint32_t word = 0;
      if (k < kMd5ChunkWords) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const unsigned long long t = base + 4ull * k + j;
          if (t < total) {
            const unsigned long long smp = bps == 2 ? t >> 1 : t;
            const int y = (int)(smp / (unsigned)P.w), x = (int)(smp - (unsigned long long)y * (unsigned)P.w);
            const uint32_t v = (uint32_t)(int)P.p[(ptrdiff_t)y * P.stride + x];
            word |= ((bps == 2 && (t & 1)) ? (v >> 8) & 0xffu : v & 0xffu) << (8 * j);
          }
        }
      }
      sbuf[k] = word;
    }
    __syncwarp();
    if (lane == 0) {
      unsigned blocks = n / 64;
      if (base + n >= total) {                       // last chunk: 0x80, zeros to 56 mod 64, bit length little endian
        const unsigned rem = n;                      // bytes of this chunk (the words behind them are zero)
        sbuf[rem >> 2] |= 0x80u << (8 * (rem & 3));
        blocks = (rem + 1 + 8 + 63) / 64;
        const unsigned long long bits = total * 8ull;
        sbuf[blocks * 16 - 2] = (uint32_t)bits;
        sbuf[blocks * 16 - 1] = (uint32_t)(bits >> 32);
      }
      for (unsigned bk = 0; bk < blocks; bk++) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = sbuf[bk * 16 + i];
        md5_block(m, w);
      }
    }
    __syncwarp();
    if (total == 0) break;
  }
  if (lane == 0) {
    uint8_t* d = digest + 16 * blockIdx.x;
    const uint32_t v[4] = {m.a, m.b, m.c, m.d};
    for (int i = 0; i < 16; i++) d[i] = (uint8_t)(v[i >> 2] >> (8 * (i & 3)));
  }
}

}  // namespace tvc

using namespace tvc;

extern "C" {

int tvc_pic_hash(tvc_ctx* c, int slot, int method, uint8_t* digest)
{
  if (!c || !valid_slot(c, slot) || !digest || method < TVC_HASH_MD5 || method > TVC_HASH_CHECKSUM)
    return set_err(c, TVC_ERR_ARG, "tvc_pic_hash: bad argument (method 1 MD5, 2 CRC, 3 checksum)");
  const Pic& p = c->pics[slot];
  const int bd = c->cfg.bit_depth;
  int r;
  // device scratch: [digest 48][checksums 3 x u32][row CRCs]
  const size_t rows_off = 256, need = rows_off + (size_t)p.h[0] * 2 * 3;
  if ((r = ensure_scratch(c, c->out, need))) return r;
  uint8_t* d = (uint8_t*)c->out.dev;
  TVC_CUDA(c, cudaMemsetAsync(d, 0, rows_off, c->stream));
  ProfScope ps(c, TVC_PH_OTHER);
  if (method == TVC_HASH_MD5) {
    MdPlane m[3];
    for (int pl = 0; pl < 3; pl++) m[pl] = MdPlane{p.org[pl], p.stride[pl], p.w[pl], p.h[pl]};
    k_md5_planes<<<3, 32, 0, c->stream>>>(m[0], m[1], m[2], bd, d);
    TVC_LAUNCH_CHECK(c);
  } else if (method == TVC_HASH_CRC) {
    for (int pl = 0; pl < 3; pl++) {
      uint16_t* rows = (uint16_t*)(d + rows_off) + (size_t)pl * p.h[0];
      k_crc_rows<<<(p.h[pl] + 63) / 64, 64, 0, c->stream>>>(p.org[pl], p.stride[pl], p.w[pl], p.h[pl], bd, rows);
      TVC_LAUNCH_CHECK(c);
      k_crc_fold<<<1, 32, 0, c->stream>>>(rows, p.w[pl], p.h[pl], bd, d + 16 * pl);
      TVC_LAUNCH_CHECK(c);
    }
  } else {
    uint32_t* sums = (uint32_t*)(d + 64);
    for (int pl = 0; pl < 3; pl++) {
      k_checksum_plane<<<kNumSM * 4, 256, 0, c->stream>>>(p.org[pl], p.stride[pl], p.w[pl], p.h[pl], bd, sums + pl);
      TVC_LAUNCH_CHECK(c);
    }
  }
  TVC_CUDA(c, cudaMemcpyAsync(c->out.host, d, rows_off, cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  const uint8_t* ho = (const uint8_t*)c->out.host;
  memset(digest, 0, 48);
  if (method == TVC_HASH_CHECKSUM) {
    const uint32_t* sums = (const uint32_t*)(ho + 64);
    for (int pl = 0; pl < 3; pl++) {            // digest bytes big endian (:160-163)
      digest[16 * pl + 0] = (uint8_t)(sums[pl] >> 24); digest[16 * pl + 1] = (uint8_t)(sums[pl] >> 16);
      digest[16 * pl + 2] = (uint8_t)(sums[pl] >> 8); digest[16 * pl + 3] = (uint8_t)sums[pl];
    }
  } else
    memcpy(digest, ho, 48);
  return TVC_OK;
}

int tvc_pic_ssd(tvc_ctx* c, int slot_a, int slot_b, uint64_t* ssd)
{
  if (!c || !valid_slot(c, slot_a) || !valid_slot(c, slot_b) || !ssd) return set_err(c, TVC_ERR_ARG, "tvc_pic_ssd: bad argument");
  const Pic& a = c->pics[slot_a];
  const Pic& b = c->pics[slot_b];
  int r;
  if ((r = ensure_scratch(c, c->out, 256))) return r;
  unsigned long long* d = (unsigned long long*)c->out.dev;
  TVC_CUDA(c, cudaMemsetAsync(d, 0, 24, c->stream));
  {
    ProfScope ps(c, TVC_PH_OTHER);
    for (int pl = 0; pl < 3; pl++) {
      k_ssd_plane<<<kNumSM * 4, 256, 0, c->stream>>>(a.org[pl], b.org[pl], a.stride[pl], a.w[pl], a.h[pl], d + pl);
      TVC_LAUNCH_CHECK(c);
    }
  }
  TVC_CUDA(c, cudaMemcpyAsync(c->out.host, d, 24, cudaMemcpyDeviceToHost, c->stream));
  TVC_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(ssd, c->out.host, 24);
  return TVC_OK;
}

}  // extern "C"
