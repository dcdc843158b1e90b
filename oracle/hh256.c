/*
 * oracle/hh256.c — CPU ORACLE (test infrastructure, not product code).
 * HighwayHash-256 as computed by github.com/minio/highwayhash v1.0.3 (absent from
 * /root/reference; go.mod:58), restated from the published portable algorithm
 * (google/highwayhash c/highwayhash.c).  hash.Hash wrapper
 * semantics follow cmd/bitrot-streaming.go:57-59 (Reset / Write / Sum(nil)).
 * Pinned by cmd/bitrot.go:37 (pi key), cmd/bitrot.go:228 (chain digest) and fixture frames.
 */
#include <string.h>
#include "oracle.h"

const uint8_t orc_magic_hh_key[32] = { /* cmd/bitrot.go:37 */
    0x4b, 0xe7, 0x34, 0xfa, 0x8e, 0x23, 0x8a, 0xcd, 0x26, 0x3e, 0x83, 0xe6, 0xbb, 0x96, 0x85, 0x52,
    0x04, 0x0f, 0x93, 0x5d, 0xa3, 0x9f, 0x44, 0x14, 0x97, 0xe0, 0x9d, 0x13, 0x22, 0xde, 0x36, 0xa0};

static uint64_t le64(const uint8_t *p) {
  uint64_t v = 0;
  for (int i = 7; i >= 0; i--) v = (v << 8) | p[i];
  return v;
}

static void hh_reset(orc_hh256_ctx *s) {
  static const uint64_t init0[4] = {0xdbe6d5d5fe4cce2full, 0xa4093822299f31d0ull,
                                    0x13198a2e03707344ull, 0x243f6a8885a308d3ull};
  static const uint64_t init1[4] = {0x3bd39e10cb0ef593ull, 0xc0acf169b5f18a8cull,
                                    0xbe5466cf34e90c6cull, 0x452821e638d01377ull};
  for (int i = 0; i < 4; i++) {
    s->mul0[i] = init0[i];
    s->mul1[i] = init1[i];
    s->v0[i] = init0[i] ^ s->key[i];
    s->v1[i] = init1[i] ^ ((s->key[i] >> 32) | (s->key[i] << 32));
  }
  s->nbuf = 0;
}

static void zipper_merge_add(uint64_t v1, uint64_t v0, uint64_t *add1, uint64_t *add0) {
  *add0 += (((v0 & 0x00000000ff000000ull) | (v1 & 0x000000ff00000000ull)) >> 24) |
           (((v0 & 0x0000ff0000000000ull) | (v1 & 0x00ff000000000000ull)) >> 16) |
           (v0 & 0x0000000000ff0000ull) | ((v0 & 0x000000000000ff00ull) << 32) |
           ((v1 & 0xff00000000000000ull) >> 8) | (v0 << 56);
  *add1 += (((v1 & 0x00000000ff000000ull) | (v0 & 0x000000ff00000000ull)) >> 24) |
           (v1 & 0x0000000000ff0000ull) | ((v1 & 0x0000ff0000000000ull) >> 16) |
           ((v1 & 0x000000000000ff00ull) << 24) | ((v0 & 0x00ff000000000000ull) >> 8) |
           ((v1 & 0x00000000000000ffull) << 48) | (v0 & 0xff00000000000000ull);
}

static void hh_update(orc_hh256_ctx *s, const uint64_t lanes[4]) {
  for (int i = 0; i < 4; i++) {
    s->v1[i] += s->mul0[i] + lanes[i];
    s->mul0[i] ^= (s->v1[i] & 0xffffffffull) * (s->v0[i] >> 32);
    s->v0[i] += s->mul1[i];
    s->mul1[i] ^= (s->v0[i] & 0xffffffffull) * (s->v1[i] >> 32);
  }
  zipper_merge_add(s->v1[1], s->v1[0], &s->v0[1], &s->v0[0]);
  zipper_merge_add(s->v1[3], s->v1[2], &s->v0[3], &s->v0[2]);
  zipper_merge_add(s->v0[1], s->v0[0], &s->v1[1], &s->v1[0]);
  zipper_merge_add(s->v0[3], s->v0[2], &s->v1[3], &s->v1[2]);
}

static void hh_update_packet(orc_hh256_ctx *s, const uint8_t *p) {
  uint64_t lanes[4] = {le64(p), le64(p + 8), le64(p + 16), le64(p + 24)};
  hh_update(s, lanes);
}

static void hh_update_remainder(orc_hh256_ctx *s, const uint8_t *bytes, size_t size_mod32) {
  const size_t size_mod4 = size_mod32 & 3;
  const uint8_t *remainder = bytes + (size_mod32 & ~(size_t)3);
  uint8_t packet[32] = {0};
  for (int i = 0; i < 4; i++) s->v0[i] += ((uint64_t)size_mod32 << 32) + size_mod32;
  for (int i = 0; i < 4; i++) { /* rotate each 32-bit half left by size_mod32 */
    uint32_t h0 = (uint32_t)s->v1[i], h1 = (uint32_t)(s->v1[i] >> 32);
    unsigned c = (unsigned)size_mod32;
    h0 = (h0 << c) | (h0 >> (32 - c));
    h1 = (h1 << c) | (h1 >> (32 - c));
    s->v1[i] = (uint64_t)h0 | ((uint64_t)h1 << 32);
  }
  memcpy(packet, bytes, (size_t)(remainder - bytes));
  if (size_mod32 & 16) {
    for (int i = 0; i < 4; i++) packet[28 + i] = remainder[i + (int)size_mod4 - 4];
  } else if (size_mod4) {
    packet[16] = remainder[0];
    packet[17] = remainder[size_mod4 >> 1];
    packet[18] = remainder[size_mod4 - 1];
  }
  hh_update_packet(s, packet);
}

void orc_hh256_init(orc_hh256_ctx *s, const uint8_t key[32]) {
  for (int i = 0; i < 4; i++) s->key[i] = le64(key + 8 * i);
  hh_reset(s);
}

void orc_hh256_write(orc_hh256_ctx *s, const uint8_t *p, size_t n) {
  if (s->nbuf) {
    size_t take = 32 - s->nbuf;
    if (take > n) take = n;
    memcpy(s->buf + s->nbuf, p, take);
    s->nbuf += (uint32_t)take;
    p += take;
    n -= take;
    if (s->nbuf == 32) { hh_update_packet(s, s->buf); s->nbuf = 0; }
  }
  while (n >= 32) { hh_update_packet(s, p); p += 32; n -= 32; }
  if (n) { memcpy(s->buf, p, n); s->nbuf = (uint32_t)n; }
}

static void modred(uint64_t a3u, uint64_t a2, uint64_t a1, uint64_t a0, uint64_t *m1, uint64_t *m0) {
  uint64_t a3 = a3u & 0x3FFFFFFFFFFFFFFFull;
  *m1 = a1 ^ ((a3 << 1) | (a2 >> 63)) ^ ((a3 << 2) | (a2 >> 62));
  *m0 = a0 ^ (a2 << 1) ^ (a2 << 2);
}

void orc_hh256_sum(const orc_hh256_ctx *in, uint8_t out[32]) {
  orc_hh256_ctx s = *in;
  if (s.nbuf) hh_update_remainder(&s, s.buf, s.nbuf);
  for (int r = 0; r < 10; r++) {
    uint64_t p[4];
    p[0] = (s.v0[2] >> 32) | (s.v0[2] << 32);
    p[1] = (s.v0[3] >> 32) | (s.v0[3] << 32);
    p[2] = (s.v0[0] >> 32) | (s.v0[0] << 32);
    p[3] = (s.v0[1] >> 32) | (s.v0[1] << 32);
    hh_update(&s, p);
  }
  uint64_t h[4];
  modred(s.v1[1] + s.mul1[1], s.v1[0] + s.mul1[0], s.v0[1] + s.mul0[1], s.v0[0] + s.mul0[0], &h[1], &h[0]);
  modred(s.v1[3] + s.mul1[3], s.v1[2] + s.mul1[2], s.v0[3] + s.mul0[3], s.v0[2] + s.mul0[2], &h[3], &h[2]);
  for (int i = 0; i < 4; i++)
    for (int b = 0; b < 8; b++) out[8 * i + b] = (uint8_t)(h[i] >> (8 * b));
}

void orc_hh256(const uint8_t key[32], const uint8_t *p, size_t n, uint8_t out[32]) {
  orc_hh256_ctx s;
  orc_hh256_init(&s, key);
  orc_hh256_write(&s, p, n);
  orc_hh256_sum(&s, out);
}
