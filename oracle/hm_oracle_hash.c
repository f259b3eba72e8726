/*
 * hm_oracle_hash.c -- CPU restatement of the picture hashes and the PSNR sums.
 *
 * TEST INFRASTRUCTURE ONLY (see hm_oracle.h).  Follows TLibCommon/TComPicYuvMD5.cpp: md5_plane :66-86 (samples packed little
 * endian, 1 byte for bit depth <= 8 else 2, rows in raster order; the digest itself is RFC 1321, which the reference takes from
 * libmd5), compCRC :88-117 (CRC-16, polynomial 0x1021, initial value 0xffff, 16 flushing zero bits, bits of a sample taken at
 * positions dataMsbIdx - (bitIdx & dataMsbIdx) -- for 10-bit data that is 9,8,9,8,9,8,9,8,1,0, restated as written), compChecksum
 * :136-165 (byte sum with the (x, y) xor mask, modulo 2^32), and TEncGOP::xCalculateAddPSNR TLibEncoder/TEncGOP.cpp:1582-1641
 * (the three UInt64 sums of squared differences; the log10 stays with the caller).
 */
#include <string.h>
#include "hm_oracle.h"

/* ---- RFC 1321 */
typedef struct { uint32_t a, b, c, d; uint64_t len; unsigned char buf[64]; unsigned fill; } orc_md5;
static const uint32_t K[64] = {
  0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1,
  0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453,
  0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942,
  0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05,
  0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d,
  0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
static const unsigned char R[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20,
                                    5, 9, 14, 20, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21,
                                    6, 10, 15, 21, 6, 10, 15, 21};

static void md5_block(orc_md5* m, const unsigned char* p)
{
  uint32_t w[16], a = m->a, b = m->b, c = m->c, d = m->d;
  for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) | ((uint32_t)p[4 * i + 3] << 24);
  for (int i = 0; i < 64; i++) {
    uint32_t f; int g;
    if (i < 16) { f = (b & c) | (~b & d); g = i; }
    else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
    else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
    else { f = c ^ (b | ~d); g = (7 * i) & 15; }
    const uint32_t t = a + f + K[i] + w[g];
    a = d; d = c; c = b;
    b = b + ((t << R[i]) | (t >> (32 - R[i])));
  }
  m->a += a; m->b += b; m->c += c; m->d += d;
}

static void md5_init(orc_md5* m) { m->a = 0x67452301; m->b = 0xefcdab89; m->c = 0x98badcfe; m->d = 0x10325476; m->len = 0; m->fill = 0; }

static void md5_update(orc_md5* m, const unsigned char* p, unsigned n)
{
  m->len += n;
  while (n) {
    unsigned k = 64 - m->fill;
    if (k > n) k = n;
    memcpy(m->buf + m->fill, p, k);
    m->fill += k; p += k; n -= k;
    if (m->fill == 64) { md5_block(m, m->buf); m->fill = 0; }
  }
}

static void md5_final(orc_md5* m, unsigned char out[16])
{
  const uint64_t bits = m->len * 8;
  unsigned char pad[72] = {0x80};
  const unsigned padn = (m->fill < 56 ? 56 : 120) - m->fill;
  md5_update(m, pad, padn);
  unsigned char lenb[8];
  for (int i = 0; i < 8; i++) lenb[i] = (unsigned char)(bits >> (8 * i));
  md5_update(m, lenb, 8);
  const uint32_t v[4] = {m->a, m->b, m->c, m->d};
  for (int i = 0; i < 16; i++) out[i] = (unsigned char)(v[i >> 2] >> (8 * (i & 3)));
}

/* TComPicYuvMD5.cpp:46-86 + calcMD5 :175-200, one plane */
void orc_md5_plane(const Pel* plane, int w, int h, int stride, int bd, unsigned char digest[16])
{
  orc_md5 m;
  md5_init(&m);
  unsigned char buf[128];
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x += 32) {
      const int n = w - x < 32 ? w - x : 32;
      int k = 0;
      for (int i = 0; i < n; i++) {
        const Pel p = plane[(size_t)y * stride + x + i];
        buf[k++] = (unsigned char)p;
        if (bd > 8) buf[k++] = (unsigned char)(p >> 8);
      }
      md5_update(&m, buf, (unsigned)k);
    }
  md5_final(&m, digest);
}

/* compCRC, TComPicYuvMD5.cpp:88-117 */
void orc_crc_plane(const Pel* plane, int w, int h, int stride, int bd, unsigned char digest[16])
{
  const unsigned msb = (unsigned)bd - 1;
  unsigned crc = 0xffff;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
      for (unsigned b = 0; b < (unsigned)bd; b++) {
        const unsigned top = (crc >> 15) & 1;
        const unsigned bit = ((unsigned)plane[(size_t)y * stride + x] >> (msb - (b & msb))) & 1;
        crc = (((crc << 1) + bit) & 0xffff) ^ (top * 0x1021);
      }
  for (int b = 0; b < 16; b++) {
    const unsigned top = (crc >> 15) & 1;
    crc = ((crc << 1) & 0xffff) ^ (top * 0x1021);
  }
  memset(digest, 0, 16);
  digest[0] = (unsigned char)(crc >> 8);
  digest[1] = (unsigned char)crc;
}

/* compChecksum, TComPicYuvMD5.cpp:136-165 */
void orc_checksum_plane(const Pel* plane, int w, int h, int stride, int bd, unsigned char digest[16])
{
  uint32_t sum = 0;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const unsigned char mask = (unsigned char)((x & 0xff) ^ (y & 0xff) ^ (x >> 8) ^ (y >> 8));
      const int p = plane[(size_t)y * stride + x];
      sum += (uint32_t)((p & 0xff) ^ mask);
      if (bd > 8) sum += (uint32_t)((p >> 8) ^ mask);
    }
  memset(digest, 0, 16);
  digest[0] = (unsigned char)(sum >> 24); digest[1] = (unsigned char)(sum >> 16); digest[2] = (unsigned char)(sum >> 8); digest[3] = (unsigned char)sum;
}

/* one sum of xCalculateAddPSNR, TEncGOP.cpp:1606-1614 */
uint64_t orc_ssd_plane(const Pel* a, int sa, const Pel* b, int sb, int w, int h)
{
  uint64_t s = 0;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const int d = (int)a[(size_t)y * sa + x] - (int)b[(size_t)y * sb + x];
      s += (uint64_t)(int64_t)(d * d);
    }
  return s;
}
