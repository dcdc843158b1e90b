// rs_matrix.cc — host-side runtime coding / decode matrices (see gf256.h).
// Mirrors what reedsolomon.New / Reconstruct compute for cmd/erasure-coding.go:63,106,112.
#include <vector>
#include "gf256.h"

namespace mec {

bool rs_coding_matrix(int k, int m, uint8_t* out) {
  if (k <= 0 || m < 0 || k + m > kMaxShards) return false;
  const int n = k + m;
  std::vector<uint8_t> vm(static_cast<size_t>(n) * k), top(static_cast<size_t>(k) * k),
      scratch(static_cast<size_t>(2) * k * k);
  for (int r = 0; r < n; r++)
    for (int c = 0; c < k; c++) vm[static_cast<size_t>(r) * k + c] = gf_pow(static_cast<uint8_t>(r), c);
  for (size_t i = 0; i < top.size(); i++) top[i] = vm[i];
  if (!gf_invert(top.data(), k, scratch.data())) return false;
  for (int r = 0; r < n; r++)
    for (int c = 0; c < k; c++) {
      uint8_t acc = 0;
      for (int t = 0; t < k; t++) acc ^= gf_mul(vm[static_cast<size_t>(r) * k + t], top[static_cast<size_t>(t) * k + c]);
      out[static_cast<size_t>(r) * k + c] = acc;
    }
  return true;
}

bool rs_decode_rows(int k, int m, const uint8_t* present, const int* missing, int nmiss, uint8_t* rows,
                    int* valid) {
  const int n = k + m;
  std::vector<uint8_t> mat(static_cast<size_t>(n) * k);
  if (!rs_coding_matrix(k, m, mat.data())) return false;
  std::vector<uint8_t> sub(static_cast<size_t>(k) * k), scratch(static_cast<size_t>(2) * k * k);
  int t = 0;
  for (int i = 0; i < n && t < k; i++)
    if (present[i]) {
      for (int c = 0; c < k; c++) sub[static_cast<size_t>(t) * k + c] = mat[static_cast<size_t>(i) * k + c];
      valid[t++] = i;
    }
  if (t < k) return false;
  if (!gf_invert(sub.data(), k, scratch.data())) return false;
  for (int q = 0; q < nmiss; q++) {
    const int idx = missing[q];
    for (int c = 0; c < k; c++) {
      if (idx < k) {
        rows[static_cast<size_t>(q) * k + c] = sub[static_cast<size_t>(idx) * k + c];
      } else {
        uint8_t acc = 0;
        for (int u = 0; u < k; u++)
          acc ^= gf_mul(mat[static_cast<size_t>(idx) * k + u], sub[static_cast<size_t>(u) * k + c]);
        rows[static_cast<size_t>(q) * k + c] = acc;
      }
    }
  }
  return true;
}

}  // namespace mec
