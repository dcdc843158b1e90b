// gf256.h — GF(2^8) arithmetic (poly 0x11D, generator 2) and the Reed-Solomon coding matrix that
// reedsolomon.New(k, m) builds with default options (Vandermonde rows r^c made systematic by
// multiplying with the inverse of the top k×k block) — the construction MinIO selects at
// cmd/erasure-coding.go:63.  Everything is constexpr so the CUDA kernels can be specialised on the
// matrix at compile time; the same functions run on the host for runtime (decode) matrices.
#pragma once
#include "rtc_compat.h"

namespace mec {

struct GfTables {
  uint8_t exp[512];
  uint8_t log[256];
};

constexpr GfTables make_gf_tables() {
  GfTables t{};
  unsigned x = 1;
  for (int i = 0; i < 255; i++) {
    t.exp[i] = static_cast<uint8_t>(x);
    t.log[x] = static_cast<uint8_t>(i);
    x <<= 1;
    if (x & 0x100) x ^= 0x11D;
  }
  for (int i = 255; i < 512; i++) t.exp[i] = t.exp[i - 255];
  return t;
}
inline constexpr GfTables kGf = make_gf_tables();

constexpr uint8_t gf_mul(uint8_t a, uint8_t b) {
  if (a == 0 || b == 0) return 0;
  return kGf.exp[kGf.log[a] + kGf.log[b]];
}
constexpr uint8_t gf_inv(uint8_t a) { return a ? kGf.exp[255 - kGf.log[a]] : 0; }
constexpr uint8_t gf_pow(uint8_t a, int n) {
  if (n == 0) return 1;
  if (a == 0) return 0;
  return kGf.exp[(static_cast<int>(kGf.log[a]) * n) % 255];
}

// In-place inverse of an n×n matrix (row-major, leading dimension n) by Gauss-Jordan elimination.
// Returns false when singular.  `scratch` must hold 2*n*n bytes.
constexpr bool gf_invert(uint8_t* mat, int n, uint8_t* scratch) {
  const int w = 2 * n;
  uint8_t* a = scratch;
  for (int r = 0; r < n; r++)
    for (int c = 0; c < w; c++) a[r * w + c] = c < n ? mat[r * n + c] : (c - n == r ? 1 : 0);
  for (int r = 0; r < n; r++) {
    if (a[r * w + r] == 0) {
      int below = -1;
      for (int rb = r + 1; rb < n; rb++)
        if (a[rb * w + r]) { below = rb; break; }
      if (below < 0) return false;
      for (int c = 0; c < w; c++) {
        uint8_t tmp = a[r * w + c];
        a[r * w + c] = a[below * w + c];
        a[below * w + c] = tmp;
      }
    }
    const uint8_t p = a[r * w + r];
    if (p != 1) {
      const uint8_t s = gf_inv(p);
      for (int c = 0; c < w; c++) a[r * w + c] = gf_mul(a[r * w + c], s);
    }
    for (int r2 = 0; r2 < n; r2++) {
      if (r2 == r) continue;
      const uint8_t f = a[r2 * w + r];
      if (!f) continue;
      for (int c = 0; c < w; c++) a[r2 * w + c] ^= gf_mul(f, a[r * w + c]);
    }
  }
  for (int r = 0; r < n; r++)
    for (int c = 0; c < n; c++) mat[r * n + c] = a[r * w + n + c];
  return true;
}

constexpr int kMaxShards = 256;

// Compile-time coding matrix for the kernels specialised on (K, M).
template <int K, int M>
struct CodingMatrix {
  uint8_t v[K + M][K];
};

template <int K, int M>
constexpr CodingMatrix<K, M> build_coding_matrix() {
  CodingMatrix<K, M> out{};
  uint8_t vm[(K + M) * K] = {};
  uint8_t top[K * K] = {};
  uint8_t scratch[2 * K * K] = {};
  for (int r = 0; r < K + M; r++)
    for (int c = 0; c < K; c++) vm[r * K + c] = gf_pow(static_cast<uint8_t>(r), c);
  for (int i = 0; i < K * K; i++) top[i] = vm[i];
  gf_invert(top, K, scratch);
  for (int r = 0; r < K + M; r++)
    for (int c = 0; c < K; c++) {
      uint8_t acc = 0;
      for (int t = 0; t < K; t++) acc ^= gf_mul(vm[r * K + t], top[t * K + c]);
      out.v[r][c] = acc;
    }
  return out;
}

// Runtime (host) versions — rs_matrix.cc
// out: (k+m) x k row-major.  Returns false for invalid (k, m).
bool rs_coding_matrix(int k, int m, uint8_t* out);
// Rows that rebuild the shards listed in `missing` from the first k present shards (ascending
// index, returned in `valid`).  rows: nmiss x k.  Data shards come from the inverted sub-matrix;
// parity shards from (parity row) x (inverted sub-matrix), which equals klauspost's two-step
// "rebuild data, then re-encode parity" because the code is MDS.  Returns false if < k present.
bool rs_decode_rows(int k, int m, const uint8_t* present, const int* missing, int nmiss, uint8_t* rows,
                    int* valid);

}  // namespace mec
