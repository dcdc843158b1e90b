/*
 * oracle/rs.c — CPU ORACLE (test infrastructure, not product code).
 * GF(2^8) arithmetic and the default Reed-Solomon construction of
 * github.com/klauspost/reedsolomon v1.12.4 (absent from /root/reference; go.mod:49),
 * restated from its published algorithm (galois.go / matrix.go / reedsolomon.go) and
 * pinned by cmd/erasure-coding.go:160 (60 golden xxhash64 values) — see SURVEY.md App. A.
 * Call sites followed: cmd/erasure-coding.go:63 (New), :81 (Split), :85 (Encode),
 * :106 (ReconstructData), :112 (Reconstruct).
 */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

static uint8_t gf_exp_tbl[512];
static uint8_t gf_log_tbl[256];
static int gf_ready;

static void gf_init(void) {
  if (gf_ready) return;
  unsigned x = 1;
  for (int i = 0; i < 255; i++) {
    gf_exp_tbl[i] = (uint8_t)x;
    gf_log_tbl[x] = (uint8_t)i;
    x <<= 1;
    if (x & 0x100) x ^= 0x11D; /* x^8+x^4+x^3+x^2+1 */
  }
  for (int i = 255; i < 512; i++) gf_exp_tbl[i] = gf_exp_tbl[i - 255];
  gf_ready = 1;
}

uint8_t orc_gf_mul(uint8_t a, uint8_t b) {
  gf_init();
  if (!a || !b) return 0;
  return gf_exp_tbl[gf_log_tbl[a] + gf_log_tbl[b]];
}
uint8_t orc_gf_inv(uint8_t a) {
  gf_init();
  return a ? gf_exp_tbl[255 - gf_log_tbl[a]] : 0;
}
/* galExp: a^n */
uint8_t orc_gf_exp(uint8_t a, int n) {
  gf_init();
  if (n == 0) return 1;
  if (a == 0) return 0;
  return gf_exp_tbl[((int)gf_log_tbl[a] * n) % 255];
}

int orc_gf_invert(uint8_t *mat, int n) {
  gf_init();
  int w = 2 * n;
  uint8_t *a = (uint8_t *)calloc((size_t)n * w, 1);
  for (int r = 0; r < n; r++) {
    memcpy(a + (size_t)r * w, mat + (size_t)r * n, n);
    a[(size_t)r * w + n + r] = 1;
  }
  for (int r = 0; r < n; r++) {
    if (a[(size_t)r * w + r] == 0) {
      int below = -1;
      for (int rb = r + 1; rb < n; rb++)
        if (a[(size_t)rb * w + r]) { below = rb; break; }
      if (below < 0) { free(a); return -1; }
      for (int c = 0; c < w; c++) {
        uint8_t t = a[(size_t)r * w + c];
        a[(size_t)r * w + c] = a[(size_t)below * w + c];
        a[(size_t)below * w + c] = t;
      }
    }
    uint8_t p = a[(size_t)r * w + r];
    if (p != 1) {
      uint8_t s = orc_gf_inv(p);
      for (int c = 0; c < w; c++) a[(size_t)r * w + c] = orc_gf_mul(a[(size_t)r * w + c], s);
    }
    for (int r2 = 0; r2 < n; r2++) {
      if (r2 == r) continue;
      uint8_t f = a[(size_t)r2 * w + r];
      if (!f) continue;
      for (int c = 0; c < w; c++) a[(size_t)r2 * w + c] ^= orc_gf_mul(f, a[(size_t)r * w + c]);
    }
  }
  for (int r = 0; r < n; r++) memcpy(mat + (size_t)r * n, a + (size_t)r * w + n, n);
  free(a);
  return 0;
}

/* buildMatrix: vandermonde(total, k) * inverse(top k rows) */
int orc_rs_matrix(int k, int m, uint8_t *out) {
  if (k <= 0 || m < 0) return ORC_ERR_INV_SHARD_NUM;
  if (k + m > 256) return ORC_ERR_MAX_SHARD_NUM;
  int n = k + m;
  uint8_t *vm = (uint8_t *)malloc((size_t)n * k);
  uint8_t *top = (uint8_t *)malloc((size_t)k * k);
  for (int r = 0; r < n; r++)
    for (int c = 0; c < k; c++) vm[(size_t)r * k + c] = orc_gf_exp((uint8_t)r, c);
  memcpy(top, vm, (size_t)k * k);
  if (orc_gf_invert(top, k)) { free(vm); free(top); return ORC_ERR_INV_SHARD_NUM; }
  for (int r = 0; r < n; r++)
    for (int c = 0; c < k; c++) {
      uint8_t acc = 0;
      for (int t = 0; t < k; t++) acc ^= orc_gf_mul(vm[(size_t)r * k + t], top[(size_t)t * k + c]);
      out[(size_t)r * k + c] = acc;
    }
  free(vm);
  free(top);
  return 0;
}

int64_t orc_rs_split(int k, int m, const uint8_t *data, int64_t len, uint8_t *dst) {
  (void)m;
  if (len == 0) return ORC_ERR_SHORT_DATA;
  int64_t per = (len + k - 1) / k;
  memcpy(dst, data, (size_t)len);
  memset(dst + len, 0, (size_t)(per * k - len));
  return per;
}

static void mul_acc_row(const uint8_t *row, int k, uint8_t *const *in, uint8_t *out, int64_t per) {
  memset(out, 0, (size_t)per);
  for (int c = 0; c < k; c++) {
    uint8_t f = row[c];
    if (!f) continue;
    const uint8_t *s = in[c];
    unsigned lf = gf_log_tbl[f];
    for (int64_t x = 0; x < per; x++) {
      uint8_t v = s[x];
      if (v) out[x] ^= gf_exp_tbl[lf + gf_log_tbl[v]];
    }
  }
}

int orc_rs_encode(int k, int m, uint8_t *const *shards, int64_t per) {
  gf_init();
  if (per == 0) return ORC_ERR_SHARD_NO_DATA;
  uint8_t *mat = (uint8_t *)malloc((size_t)(k + m) * k);
  int rc = orc_rs_matrix(k, m, mat);
  if (rc) { free(mat); return rc; }
  for (int j = 0; j < m; j++) mul_acc_row(mat + (size_t)(k + j) * k, k, shards, shards[k + j], per);
  free(mat);
  return 0;
}

int orc_rs_decode_rows(int k, int m, const uint8_t *present, const int *missing, int nmiss,
                       uint8_t *rows, int *valid) {
  gf_init();
  int n = k + m, np = 0;
  for (int i = 0; i < n; i++) np += present[i] != 0;
  if (np < k) return ORC_ERR_TOO_FEW_SHARDS;
  uint8_t *mat = (uint8_t *)malloc((size_t)n * k);
  int rc = orc_rs_matrix(k, m, mat);
  if (rc) { free(mat); return rc; }
  uint8_t *sub = (uint8_t *)malloc((size_t)k * k);
  int t = 0;
  for (int i = 0; i < n && t < k; i++)
    if (present[i]) { memcpy(sub + (size_t)t * k, mat + (size_t)i * k, k); valid[t++] = i; }
  if (orc_gf_invert(sub, k)) { free(mat); free(sub); return ORC_ERR_TOO_FEW_SHARDS; }
  /* data shard d = row d of inverse; parity shard p = M[p] * inverse */
  for (int q = 0; q < nmiss; q++) {
    int idx = missing[q];
    for (int c = 0; c < k; c++) {
      if (idx < k) rows[(size_t)q * k + c] = sub[(size_t)idx * k + c];
      else {
        uint8_t acc = 0;
        for (int u = 0; u < k; u++) acc ^= orc_gf_mul(mat[(size_t)idx * k + u], sub[(size_t)u * k + c]);
        rows[(size_t)q * k + c] = acc;
      }
    }
  }
  free(mat);
  free(sub);
  return 0;
}

/* reedsolomon.reconstruct(shards, dataOnly): follows the published control flow —
 * all present -> nil; none -> ErrShardNoData; fewer than k -> ErrTooFewShards;
 * missing data from inverse of the first-k-present sub-matrix; then (unless dataOnly)
 * missing parity re-encoded from the complete data with the parity rows. */
int orc_rs_reconstruct(int k, int m, uint8_t *const *shards, const uint8_t *present, int64_t per,
                       int data_only) {
  gf_init();
  int n = k + m, np = 0;
  for (int i = 0; i < n; i++) np += present[i] != 0;
  if (np == 0) return ORC_ERR_SHARD_NO_DATA;
  if (np == n) return 0;
  if (np < k) return ORC_ERR_TOO_FEW_SHARDS;
  int *missing = (int *)malloc(sizeof(int) * n), nd = 0;
  for (int i = 0; i < k; i++)
    if (!present[i]) missing[nd++] = i;
  int rc = 0;
  if (nd) {
    uint8_t *rows = (uint8_t *)malloc((size_t)nd * k);
    int *valid = (int *)malloc(sizeof(int) * k);
    rc = orc_rs_decode_rows(k, m, present, missing, nd, rows, valid);
    if (!rc) {
      uint8_t **in = (uint8_t **)malloc(sizeof(uint8_t *) * k);
      for (int t = 0; t < k; t++) in[t] = shards[valid[t]];
      for (int q = 0; q < nd; q++) mul_acc_row(rows + (size_t)q * k, k, in, shards[missing[q]], per);
      free(in);
    }
    free(rows);
    free(valid);
  }
  if (!rc && !data_only) {
    uint8_t *mat = (uint8_t *)malloc((size_t)n * k);
    orc_rs_matrix(k, m, mat);
    for (int p = k; p < n; p++)
      if (!present[p]) mul_acc_row(mat + (size_t)p * k, k, shards, shards[p], per);
    free(mat);
  }
  free(missing);
  return rc;
}
