/*
 * oracle/erasure.c — CPU ORACLE (test infrastructure, not product code).
 * Whole-object drivers restating MinIO's streaming layer:
 *   Erasure.Encode   cmd/erasure-encode.go:69-110   (block loop, EncodeData, per-shard writers)
 *   Erasure.Decode   cmd/erasure-decode.go:239-314  (+ parallelReader.Read :127-235,
 *                                                    writeDataBlocks cmd/erasure-utils.go:42-105)
 *   Erasure.Heal     cmd/erasure-decode.go:317-364
 *   streamingBitrotWriter.Write / streamingBitrotReader.ReadAt  cmd/bitrot-streaming.go:44-75,161-200
 *   wholeBitrotWriter cmd/bitrot-whole.go:35-45, bitrotVerify cmd/bitrot.go:164-216
 *   size helpers     cmd/erasure-coding.go:116-141, cmd/bitrot.go:156-161, cmd/utils.go:689-705
 * Readers are tried in shard-index order (no `prefer`), as parallelReader does by default.
 */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

int64_t orc_ceil_frac(int64_t num, int64_t den) { /* cmd/utils.go:689 */
  if (den == 0) return 0;
  if (den < 0) { num = -num; den = -den; }
  int64_t c = num / den;
  if (num > 0 && num % den != 0) c++;
  return c;
}
int64_t orc_shard_size(int64_t bs, int k) { return orc_ceil_frac(bs, k); }
int64_t orc_shard_file_size(int64_t bs, int k, int64_t total) {
  if (total == 0) return 0;
  if (total == -1) return -1;
  int64_t num = total / bs, last = total % bs;
  return num * orc_shard_size(bs, k) + orc_ceil_frac(last, k);
}
int64_t orc_shard_file_offset(int64_t bs, int k, int64_t start, int64_t len, int64_t total) {
  int64_t ss = orc_shard_size(bs, k), sfs = orc_shard_file_size(bs, k, total);
  int64_t end_shard = (start + len) / bs;
  int64_t till = end_shard * ss + ss;
  return till < sfs ? till : sfs;
}
int64_t orc_bitrot_shard_file_size(int64_t size, int64_t shard_size, int algo) {
  if (algo != ORC_HIGHWAYHASH256S) return size;
  return orc_ceil_frac(size, shard_size) * 32 + size;
}

typedef union { orc_hh256_ctx hh; orc_sha256_ctx sha; orc_blake2b_ctx b2; } whole_ctx;
static void whole_init(int algo, whole_ctx *c) {
  if (algo == ORC_SHA256) orc_sha256_init(&c->sha);
  else if (algo == ORC_BLAKE2B512) orc_blake2b512_init(&c->b2);
  else orc_hh256_init(&c->hh, orc_magic_hh_key);
}
static void whole_write(int algo, whole_ctx *c, const uint8_t *p, size_t n) {
  if (algo == ORC_SHA256) orc_sha256_write(&c->sha, p, n);
  else if (algo == ORC_BLAKE2B512) orc_blake2b512_write(&c->b2, p, n);
  else orc_hh256_write(&c->hh, p, n);
}
static void whole_sum(int algo, const whole_ctx *c, uint8_t *out) {
  if (algo == ORC_SHA256) orc_sha256_sum(&c->sha, out);
  else if (algo == ORC_BLAKE2B512) orc_blake2b512_sum(&c->b2, out);
  else orc_hh256_sum(&c->hh, out);
}

int64_t orc_erasure_encode(int k, int m, int64_t bs, int algo, const uint8_t *src, int64_t len,
                           uint8_t *const *files, uint8_t *sums) {
  if (k <= 0 || m < 0) return ORC_ERR_INV_SHARD_NUM;
  if (k + m > 256) return ORC_ERR_MAX_SHARD_NUM;
  int n = k + m;
  int64_t S = orc_shard_size(bs, k);
  uint8_t *buf = (uint8_t *)malloc((size_t)(n * S) + 1);
  uint8_t **sh = (uint8_t **)malloc(sizeof(uint8_t *) * n);
  int64_t *pos = (int64_t *)calloc(n, sizeof(int64_t));
  whole_ctx *wc = (whole_ctx *)malloc(sizeof(whole_ctx) * n);
  int streaming = algo == ORC_HIGHWAYHASH256S;
  if (!streaming) for (int i = 0; i < n; i++) whole_init(algo, &wc[i]);
  int64_t total = 0;
  int rc = 0;
  while (total < len) { /* len==0: EncodeData returns n nil shards, writers get empty writes */
    int64_t nb = len - total < bs ? len - total : bs;
    int64_t per = orc_rs_split(k, m, src + total, nb, buf);
    for (int i = 0; i < n; i++) sh[i] = buf + (size_t)i * per;
    if (m > 0) { rc = orc_rs_encode(k, m, sh, per); if (rc) break; }
    for (int i = 0; i < n; i++) {
      if (streaming) { /* bitrot-streaming.go:57-65: hash first, then the shard */
        orc_hh256(orc_magic_hh_key, sh[i], (size_t)per, files[i] + pos[i]);
        pos[i] += 32;
      } else {
        whole_write(algo, &wc[i], sh[i], (size_t)per);
      }
      memcpy(files[i] + pos[i], sh[i], (size_t)per);
      pos[i] += per;
    }
    total += nb;
  }
  if (!rc && !streaming && sums)
    for (int i = 0; i < n; i++) whole_sum(algo, &wc[i], sums + (size_t)i * 64);
  free(buf); free(sh); free(pos); free(wc);
  return rc ? rc : total;
}

int orc_bitrot_verify(int algo, const uint8_t *file, int64_t file_len, int64_t part_len,
                      int64_t shard_size, const uint8_t *want) {
  uint8_t got[64];
  if (algo != ORC_HIGHWAYHASH256S) {
    int ds = orc_bitrot_hash(algo, file, (size_t)file_len, got);
    if (ds < 0) return ORC_ERR_INVALID_ARGUMENT;
    return memcmp(got, want, (size_t)ds) ? ORC_ERR_FILE_CORRUPT : 0;
  }
  if (file_len != orc_bitrot_shard_file_size(part_len, shard_size, algo)) return ORC_ERR_FILE_CORRUPT;
  int64_t left = file_len, off = 0;
  while (left > 0) {
    if (left < 32) return ORC_ERR_FILE_CORRUPT;
    const uint8_t *hb = file + off;
    off += 32; left -= 32;
    if (left < shard_size) shard_size = left;
    orc_hh256(orc_magic_hh_key, file + off, (size_t)shard_size, got);
    off += shard_size; left -= shard_size;
    if (memcmp(got, hb, 32)) return ORC_ERR_FILE_CORRUPT;
  }
  return 0;
}

/* parallelReader.Read for one block: fills shards[i] (pointers into `store`) for the first k
 * readable shards in index order; alive[] persists across blocks (a failed reader is dropped). */
static int read_block(int k, int n, int algo, const uint8_t *const *files, uint8_t *alive,
                      int64_t shard_off, int64_t cur, int64_t S, uint8_t **shards, uint8_t *store,
                      uint8_t *present, uint8_t *corrupt_out, int *saw_corrupt) {
  int got = 0;
  memset(present, 0, (size_t)n);
  for (int i = 0; i < n; i++) shards[i] = store + (size_t)i * S;
  for (int i = 0; i < n && got < k; i++) {
    if (!alive[i]) continue;
    if (algo == ORC_HIGHWAYHASH256S) {
      int64_t so = (shard_off / S) * 32 + shard_off; /* bitrot-streaming.go:171 */
      uint8_t d[32];
      orc_hh256(orc_magic_hh_key, files[i] + so + 32, (size_t)cur, d);
      if (memcmp(d, files[i] + so, 32)) { /* errFileCorrupt: drop reader, try the next */
        alive[i] = 0;
        if (corrupt_out) corrupt_out[i] = 1;
        *saw_corrupt = 1;
        continue;
      }
      memcpy(shards[i], files[i] + so + 32, (size_t)cur);
    } else {
      memcpy(shards[i], files[i] + shard_off, (size_t)cur); /* whole-file verify happens in bitrotVerify */
    }
    present[i] = 1;
    got++;
  }
  return got >= k ? 0 : ORC_ERR_READ_QUORUM;
}

int64_t orc_erasure_decode(int k, int m, int64_t bs, int algo, const uint8_t *const *files,
                           const uint8_t *avail, int64_t offset, int64_t length, int64_t total,
                           uint8_t *dst, uint8_t *corrupt_out) {
  if (offset < 0 || length < 0 || offset + length > total) return ORC_ERR_INVALID_ARGUMENT;
  if (length == 0) return 0;
  int n = k + m;
  int64_t S = orc_shard_size(bs, k), sfs = orc_shard_file_size(bs, k, total);
  uint8_t *store = (uint8_t *)malloc((size_t)(n * S) + 1);
  uint8_t **shards = (uint8_t **)malloc(sizeof(uint8_t *) * n);
  uint8_t *present = (uint8_t *)malloc((size_t)n), *alive = (uint8_t *)malloc((size_t)n);
  memcpy(alive, avail, (size_t)n);
  if (corrupt_out) memset(corrupt_out, 0, (size_t)n);
  int64_t start_block = offset / bs, end_block = (offset + length) / bs, written = 0;
  int64_t shard_off = start_block * S;
  int saw_corrupt = 0;
  int64_t rc = 0;
  for (int64_t block = start_block; block <= end_block; block++) {
    int64_t bo, bl;
    if (start_block == end_block) { bo = offset % bs; bl = length; }
    else if (block == start_block) { bo = offset % bs; bl = bs - bo; }
    else if (block == end_block) { bo = 0; bl = (offset + length) % bs; }
    else { bo = 0; bl = bs; }
    if (bl == 0) break;
    int64_t cur = S;
    if (shard_off + cur > sfs) cur = sfs - shard_off;
    if (cur <= 0) break;
    rc = read_block(k, n, algo, files, alive, shard_off, cur, S, shards, store, present, corrupt_out, &saw_corrupt);
    if (rc) break;
    shard_off += cur;
    /* DecodeDataBlocks: only when some shard is empty (erasure-coding.go:94-106) */
    int missing = 0;
    for (int i = 0; i < n; i++) missing += !present[i];
    if (missing) { rc = orc_rs_reconstruct(k, m, shards, present, cur, 1); if (rc) break; }
    /* writeDataBlocks (erasure-utils.go:42) */
    if (k * cur < bl) { rc = ORC_ERR_SHORT_DATA; break; }
    int64_t o = bo, w = bl;
    for (int i = 0; i < k && w > 0; i++) {
      if (o >= cur) { o -= cur; continue; }
      int64_t take = cur - o;
      if (take > w) take = w;
      memcpy(dst + written, shards[i] + o, (size_t)take);
      written += take; w -= take; o = 0;
    }
  }
  free(store); free(shards); free(present); free(alive);
  if (rc) return rc;
  if (written != length) return ORC_ERR_LESS_DATA;
  (void)saw_corrupt;
  return written;
}

int orc_erasure_heal(int k, int m, int64_t bs, int algo, const uint8_t *const *files,
                     const uint8_t *avail, const uint8_t *stale, int64_t total,
                     uint8_t *const *out_files) {
  int n = k + m;
  int64_t S = orc_shard_size(bs, k), sfs = orc_shard_file_size(bs, k, total);
  uint8_t *store = (uint8_t *)malloc((size_t)(n * S) + 1);
  uint8_t **shards = (uint8_t **)malloc(sizeof(uint8_t *) * n);
  uint8_t *present = (uint8_t *)malloc((size_t)n), *alive = (uint8_t *)malloc((size_t)n);
  int64_t *pos = (int64_t *)calloc(n, sizeof(int64_t));
  memcpy(alive, avail, (size_t)n);
  int64_t nblocks = total / bs + (total % bs != 0), shard_off = 0;
  int saw = 0, rc = 0;
  for (int64_t b = 0; b < nblocks && !rc; b++) {
    int64_t cur = S;
    if (shard_off + cur > sfs) cur = sfs - shard_off;
    rc = read_block(k, n, algo, files, alive, shard_off, cur, S, shards, store, present, NULL, &saw);
    if (rc) break;
    shard_off += cur;
    rc = orc_rs_reconstruct(k, m, shards, present, cur, 0);
    if (rc) break;
    for (int i = 0; i < n; i++) {
      if (!stale[i]) continue;
      if (algo == ORC_HIGHWAYHASH256S) {
        orc_hh256(orc_magic_hh_key, shards[i], (size_t)cur, out_files[i] + pos[i]);
        pos[i] += 32;
      }
      memcpy(out_files[i] + pos[i], shards[i], (size_t)cur);
      pos[i] += cur;
    }
  }
  free(store); free(shards); free(present); free(alive); free(pos);
  return rc;
}
