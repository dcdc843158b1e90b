/*
 * oracle/hashes.c — CPU ORACLE (test infrastructure, not product code).
 * SHA-256 (FIPS 180-4), BLAKE2b-512 (RFC 7693, unkeyed) and XXH64 — the legacy bitrot
 * algorithms of cmd/bitrot.go:47-64 and the digest erasureSelfTest uses
 * (cmd/erasure-coding.go:176-185).  Cross-checked against Python hashlib / xxhash in
 * tests/test_oracle_goldens.py and pinned by cmd/bitrot.go:226-227.
 */
#include <string.h>
#include "oracle.h"

/* ---------------- SHA-256 ---------------- */
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
#define ROR32(x, n) (((x) >> (n)) | ((x) << (32 - (n))))
static void sha256_block(uint32_t h[8], const uint8_t *p) {
  uint32_t w[64];
  for (int i = 0; i < 16; i++)
    w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
  for (int i = 16; i < 64; i++) {
    uint32_t s0 = ROR32(w[i - 15], 7) ^ ROR32(w[i - 15], 18) ^ (w[i - 15] >> 3);
    uint32_t s1 = ROR32(w[i - 2], 17) ^ ROR32(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
  for (int i = 0; i < 64; i++) {
    uint32_t S1 = ROR32(e, 6) ^ ROR32(e, 11) ^ ROR32(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = hh + S1 + ch + K256[i] + w[i];
    uint32_t S0 = ROR32(a, 2) ^ ROR32(a, 13) ^ ROR32(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
void orc_sha256_init(orc_sha256_ctx *c) {
  static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                                 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  memcpy(c->h, iv, sizeof iv);
  c->len = 0;
  c->nbuf = 0;
}
void orc_sha256_write(orc_sha256_ctx *c, const uint8_t *p, size_t n) {
  c->len += n;
  if (c->nbuf) {
    size_t take = 64 - c->nbuf;
    if (take > n) take = n;
    memcpy(c->buf + c->nbuf, p, take);
    c->nbuf += (uint32_t)take; p += take; n -= take;
    if (c->nbuf == 64) { sha256_block(c->h, c->buf); c->nbuf = 0; }
  }
  while (n >= 64) { sha256_block(c->h, p); p += 64; n -= 64; }
  if (n) { memcpy(c->buf, p, n); c->nbuf = (uint32_t)n; }
}
void orc_sha256_sum(const orc_sha256_ctx *in, uint8_t out[32]) {
  orc_sha256_ctx c = *in;
  uint8_t pad[72] = {0x80};
  uint64_t bits = c.len * 8;
  size_t padlen = (c.nbuf < 56) ? 56 - c.nbuf : 120 - c.nbuf;
  uint8_t lenb[8];
  for (int i = 0; i < 8; i++) lenb[i] = (uint8_t)(bits >> (56 - 8 * i));
  orc_sha256_write(&c, pad, padlen);
  orc_sha256_write(&c, lenb, 8);
  for (int i = 0; i < 8; i++) {
    out[4 * i] = (uint8_t)(c.h[i] >> 24); out[4 * i + 1] = (uint8_t)(c.h[i] >> 16);
    out[4 * i + 2] = (uint8_t)(c.h[i] >> 8); out[4 * i + 3] = (uint8_t)c.h[i];
  }
}
void orc_sha256(const uint8_t *p, size_t n, uint8_t out[32]) {
  orc_sha256_ctx c;
  orc_sha256_init(&c);
  orc_sha256_write(&c, p, n);
  orc_sha256_sum(&c, out);
}

/* ---------------- BLAKE2b-512 (unkeyed) ---------------- */
static const uint64_t B2IV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull,
                                 0xa54ff53a5f1d36f1ull, 0x510e527fade682d1ull, 0x9b05688c2b3e6c1full,
                                 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
static const uint8_t B2S[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
#define ROR64(x, n) (((x) >> (n)) | ((x) << (64 - (n))))
#define B2G(a, b, c, d, x, y) \
  do { a = a + b + x; d = ROR64(d ^ a, 32); c = c + d; b = ROR64(b ^ c, 24); \
       a = a + b + y; d = ROR64(d ^ a, 16); c = c + d; b = ROR64(b ^ c, 63); } while (0)
static void b2_compress(orc_blake2b_ctx *c, const uint8_t *blk, int last) {
  uint64_t m[16], v[16];
  for (int i = 0; i < 16; i++) {
    uint64_t w = 0;
    for (int b = 7; b >= 0; b--) w = (w << 8) | blk[8 * i + b];
    m[i] = w;
  }
  for (int i = 0; i < 8; i++) { v[i] = c->h[i]; v[i + 8] = B2IV[i]; }
  v[12] ^= c->t[0];
  v[13] ^= c->t[1];
  if (last) v[14] = ~v[14];
  for (int r = 0; r < 12; r++) {
    const uint8_t *s = B2S[r];
    B2G(v[0], v[4], v[8], v[12], m[s[0]], m[s[1]]);
    B2G(v[1], v[5], v[9], v[13], m[s[2]], m[s[3]]);
    B2G(v[2], v[6], v[10], v[14], m[s[4]], m[s[5]]);
    B2G(v[3], v[7], v[11], v[15], m[s[6]], m[s[7]]);
    B2G(v[0], v[5], v[10], v[15], m[s[8]], m[s[9]]);
    B2G(v[1], v[6], v[11], v[12], m[s[10]], m[s[11]]);
    B2G(v[2], v[7], v[8], v[13], m[s[12]], m[s[13]]);
    B2G(v[3], v[4], v[9], v[14], m[s[14]], m[s[15]]);
  }
  for (int i = 0; i < 8; i++) c->h[i] ^= v[i] ^ v[i + 8];
}
void orc_blake2b512_init(orc_blake2b_ctx *c) {
  memcpy(c->h, B2IV, sizeof B2IV);
  c->h[0] ^= 0x01010000ull ^ 64; /* digest length 64, no key, fanout=depth=1 */
  c->t[0] = c->t[1] = 0;
  c->nbuf = 0;
}
void orc_blake2b512_write(orc_blake2b_ctx *c, const uint8_t *p, size_t n) {
  while (n) {
    if (c->nbuf == 128) { /* buffer full and more data follows: compress as a non-final block */
      c->t[0] += 128;
      if (c->t[0] < 128) c->t[1]++;
      b2_compress(c, c->buf, 0);
      c->nbuf = 0;
    }
    size_t take = 128 - c->nbuf;
    if (take > n) take = n;
    memcpy(c->buf + c->nbuf, p, take);
    c->nbuf += (uint32_t)take; p += take; n -= take;
  }
}
void orc_blake2b512_sum(const orc_blake2b_ctx *in, uint8_t out[64]) {
  orc_blake2b_ctx c = *in;
  c.t[0] += c.nbuf;
  if (c.t[0] < c.nbuf) c.t[1]++;
  memset(c.buf + c.nbuf, 0, 128 - c.nbuf);
  b2_compress(&c, c.buf, 1);
  for (int i = 0; i < 8; i++)
    for (int b = 0; b < 8; b++) out[8 * i + b] = (uint8_t)(c.h[i] >> (8 * b));
}
void orc_blake2b512(const uint8_t *p, size_t n, uint8_t out[64]) {
  orc_blake2b_ctx c;
  orc_blake2b512_init(&c);
  orc_blake2b512_write(&c, p, n);
  orc_blake2b512_sum(&c, out);
}

/* ---------------- XXH64 ---------------- */
#define XP1 11400714785074694791ull
#define XP2 14029467366897019727ull
#define XP3 1609587929392839161ull
#define XP4 9650029242287828579ull
#define XP5 2870177450012600261ull
#define ROL64(x, n) (((x) << (n)) | ((x) >> (64 - (n))))
static uint64_t xrd64(const uint8_t *p) { uint64_t v = 0; for (int i = 7; i >= 0; i--) v = (v << 8) | p[i]; return v; }
static uint64_t xrd32(const uint8_t *p) { return (uint64_t)p[0] | ((uint64_t)p[1] << 8) | ((uint64_t)p[2] << 16) | ((uint64_t)p[3] << 24); }
static uint64_t xround(uint64_t acc, uint64_t in) { acc += in * XP2; acc = ROL64(acc, 31); return acc * XP1; }
static uint64_t xmerge(uint64_t acc, uint64_t v) { acc ^= xround(0, v); return acc * XP1 + XP4; }
uint64_t orc_xxh64(const uint8_t *p, size_t n, uint64_t seed) {
  const uint8_t *end = p + n;
  uint64_t h;
  if (n >= 32) {
    uint64_t v1 = seed + XP1 + XP2, v2 = seed + XP2, v3 = seed, v4 = seed - XP1;
    do {
      v1 = xround(v1, xrd64(p)); v2 = xround(v2, xrd64(p + 8));
      v3 = xround(v3, xrd64(p + 16)); v4 = xround(v4, xrd64(p + 24));
      p += 32;
    } while (p + 32 <= end);
    h = ROL64(v1, 1) + ROL64(v2, 7) + ROL64(v3, 12) + ROL64(v4, 18);
    h = xmerge(h, v1); h = xmerge(h, v2); h = xmerge(h, v3); h = xmerge(h, v4);
  } else {
    h = seed + XP5;
  }
  h += (uint64_t)n;
  while (p + 8 <= end) { h ^= xround(0, xrd64(p)); h = ROL64(h, 27) * XP1 + XP4; p += 8; }
  if (p + 4 <= end) { h ^= xrd32(p) * XP1; h = ROL64(h, 23) * XP2 + XP3; p += 4; }
  while (p < end) { h ^= (*p) * XP5; h = ROL64(h, 11) * XP1; p++; }
  h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
  return h;
}

/* ---------------- generic dispatch ---------------- */
int orc_bitrot_digest_size(int algo) {
  switch (algo) {
    case ORC_SHA256: case ORC_HIGHWAYHASH256: case ORC_HIGHWAYHASH256S: return 32;
    case ORC_BLAKE2B512: return 64;
  }
  return -1;
}
int orc_bitrot_hash(int algo, const uint8_t *p, size_t n, uint8_t *out) {
  switch (algo) {
    case ORC_SHA256: orc_sha256(p, n, out); return 32;
    case ORC_BLAKE2B512: orc_blake2b512(p, n, out); return 64;
    case ORC_HIGHWAYHASH256: case ORC_HIGHWAYHASH256S: orc_hh256(orc_magic_hh_key, p, n, out); return 32;
  }
  return -1;
}

/* ---- object checksums (internal/hash/checksum.go:64-73): CRC32 (IEEE), CRC32C (Castagnoli), CRC64NVME ------------------
 * Go's hash/crc32 and hash/crc64 (crc64.MakeTable(bits.Reverse64(0xad93d23594c93659)), internal/hash/crc.go:74-76): reflected,
 * init and xorout all ones.  Bit-at-a-time restatement; pinned by the catalogue check values of "123456789"
 * (0xCBF43926, 0xE3069283, 0xAE8B14860A799888) and zlib in tests/test_checksums.py. */
static uint64_t crc_bits(const uint8_t *p, size_t n, uint64_t poly, int bits) {
  const uint64_t mask = bits == 64 ? ~0ull : ((1ull << bits) - 1);
  uint64_t c = mask;
  for (size_t i = 0; i < n; i++) {
    c ^= p[i];
    for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ poly : c >> 1;
  }
  return (c ^ mask) & mask;
}
uint64_t orc_crc32(const uint8_t *p, size_t n) { return crc_bits(p, n, 0xEDB88320ull, 32); }
uint64_t orc_crc32c(const uint8_t *p, size_t n) { return crc_bits(p, n, 0x82F63B78ull, 32); }
uint64_t orc_crc64nvme(const uint8_t *p, size_t n) { return crc_bits(p, n, 0x9A6C9329AC4BC9B5ull, 64); }
