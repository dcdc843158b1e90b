/*
 * oracle/oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the arithmetic on MinIO's erasure-code + bitrot hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may link or call this library.  The product (minio_b200/csrc) never does.
 *
 * The arithmetic lives in two Go modules that are NOT vendored under /root/reference:
 *   github.com/klauspost/reedsolomon v1.12.4 (go.mod:49)  — restated in rs.c / gf256.c
 *   github.com/minio/highwayhash   v1.0.3  (go.mod:58)  — restated in hh256.c
 * Parity is PINNED by golden values that live in the reference tree itself
 * (tests/test_oracle_goldens.py): the 60 erasureSelfTest xxhash64 values
 * (cmd/erasure-coding.go:160), the bitrotSelfTest chain digests (cmd/bitrot.go:225-229),
 * the pi-derived HighwayHash key (cmd/bitrot.go:36-37) and real shard files/inline frames
 * from the cmd/testdata tgz/zip fixtures (committed as small fixtures under tests/golden/).
 */
#ifndef MINIO_B200_ORACLE_H
#define MINIO_B200_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* error codes mirror klauspost/reedsolomon Err* and MinIO storage errors */
enum {
  ORC_OK = 0,
  ORC_ERR_INV_SHARD_NUM = -1,  /* reedsolomon.ErrInvShardNum  (erasure-coding.go:45) */
  ORC_ERR_MAX_SHARD_NUM = -2,  /* reedsolomon.ErrMaxShardNum  (erasure-coding.go:49) */
  ORC_ERR_TOO_FEW_SHARDS = -3, /* reedsolomon.ErrTooFewShards */
  ORC_ERR_SHARD_NO_DATA = -4,  /* reedsolomon.ErrShardNoData */
  ORC_ERR_SHARD_SIZE = -5,     /* reedsolomon.ErrShardSize */
  ORC_ERR_SHORT_DATA = -6,     /* reedsolomon.ErrShortData */
  ORC_ERR_FILE_CORRUPT = -7,   /* errFileCorrupt (storage-errors.go:104) */
  ORC_ERR_LESS_DATA = -8,      /* errLessData   (storage-errors.go:114) */
  ORC_ERR_UNEXPECTED = -9,     /* errUnexpected (storage-errors.go:29) */
  ORC_ERR_READ_QUORUM = -10,   /* errErasureReadQuorum  (erasure-errors.go:23) */
  ORC_ERR_WRITE_QUORUM = -11,  /* errErasureWriteQuorum (erasure-errors.go:26) */
  ORC_ERR_INVALID_ARGUMENT = -12, /* errInvalidArgument */
};

/* bitrot algorithms: cmd/xl-storage-format-v1.go:142-159 */
enum { ORC_SHA256 = 1, ORC_HIGHWAYHASH256 = 2, ORC_HIGHWAYHASH256S = 3, ORC_BLAKE2B512 = 4 };

/* ---- GF(2^8), poly 0x11D, generator 2 (klauspost galois.go) ---- */
uint8_t orc_gf_mul(uint8_t a, uint8_t b);
uint8_t orc_gf_inv(uint8_t a);
uint8_t orc_gf_exp(uint8_t a, int n);
/* square matrix inverse over GF(2^8) (Gauss-Jordan); returns 0 or -1 if singular */
int orc_gf_invert(uint8_t *mat, int n);

/* ---- Reed-Solomon, klauspost default matrix (reedsolomon.New, no options) ---- */
/* writes the (k+m) x k systematic coding matrix, row-major */
int orc_rs_matrix(int k, int m, uint8_t *out);
/* Split semantic (erasure-coding.go:81): per = ceil(len/k).  Copies data into k contiguous
 * shards of `per` bytes at stride `per` in `dst` (dst must hold (k+m)*per), zero padded. */
int64_t orc_rs_split(int k, int m, const uint8_t *data, int64_t len, uint8_t *dst);
/* Encode: shards[0..k) data in, shards[k..k+m) parity out, each `per` bytes */
int orc_rs_encode(int k, int m, uint8_t *const *shards, int64_t per);
/* Reconstruct / ReconstructData: present[i]!=0 marks shard i as available.  Missing
 * shards are written in place (buffers must exist).  data_only skips parity. */
int orc_rs_reconstruct(int k, int m, uint8_t *const *shards, const uint8_t *present, int64_t per,
                       int data_only);
/* rows that rebuild the `nmiss` shards listed in `missing` from the first k present shards
 * (ascending index, listed back in `valid`): rows is nmiss x k row-major. */
int orc_rs_decode_rows(int k, int m, const uint8_t *present, const int *missing, int nmiss,
                       uint8_t *rows, int *valid);

/* ---- HighwayHash-256 (minio/highwayhash; google/highwayhash portable algorithm) ---- */
typedef struct {
  uint64_t v0[4], v1[4], mul0[4], mul1[4];
  uint8_t buf[32];
  uint32_t nbuf;
  uint64_t key[4];
} orc_hh256_ctx;
extern const uint8_t orc_magic_hh_key[32]; /* cmd/bitrot.go:37 */
void orc_hh256_init(orc_hh256_ctx *s, const uint8_t key[32]);
void orc_hh256_write(orc_hh256_ctx *s, const uint8_t *p, size_t n);
void orc_hh256_sum(const orc_hh256_ctx *s, uint8_t out[32]); /* non-destructive, like hash.Hash.Sum */
void orc_hh256(const uint8_t key[32], const uint8_t *p, size_t n, uint8_t out[32]);

/* ---- SHA-256 / BLAKE2b-512 / xxhash64 (legacy bitrot algos, self-test digest) ---- */
typedef struct { uint32_t h[8]; uint8_t buf[64]; uint64_t len; uint32_t nbuf; } orc_sha256_ctx;
void orc_sha256_init(orc_sha256_ctx *c);
void orc_sha256_write(orc_sha256_ctx *c, const uint8_t *p, size_t n);
void orc_sha256_sum(const orc_sha256_ctx *c, uint8_t out[32]);
void orc_sha256(const uint8_t *p, size_t n, uint8_t out[32]);
typedef struct { uint64_t h[8]; uint64_t t[2]; uint8_t buf[128]; uint32_t nbuf; } orc_blake2b_ctx;
void orc_blake2b512_init(orc_blake2b_ctx *c);
void orc_blake2b512_write(orc_blake2b_ctx *c, const uint8_t *p, size_t n);
void orc_blake2b512_sum(const orc_blake2b_ctx *c, uint8_t out[64]);
void orc_blake2b512(const uint8_t *p, size_t n, uint8_t out[64]);
uint64_t orc_xxh64(const uint8_t *p, size_t n, uint64_t seed);

/* generic one-shot bitrot hash; returns digest size (32 or 64) */
int orc_bitrot_hash(int algo, const uint8_t *p, size_t n, uint8_t *out);
int orc_bitrot_digest_size(int algo);

/* ---- size helpers (cmd/erasure-coding.go:116-141, cmd/bitrot.go:156, cmd/utils.go:689) ---- */
int64_t orc_ceil_frac(int64_t num, int64_t den);
int64_t orc_shard_size(int64_t block_size, int k);
int64_t orc_shard_file_size(int64_t block_size, int k, int64_t total);
int64_t orc_shard_file_offset(int64_t block_size, int k, int64_t start, int64_t len, int64_t total);
int64_t orc_bitrot_shard_file_size(int64_t size, int64_t shard_size, int algo);

/* ---- whole-object drivers: Erasure.Encode / Decode / Heal + bitrot framing ----
 * files[i] receives shard file i exactly as part.N would hold it:
 *   HighwayHash256S: ([32B digest][shard bytes])* per erasure block (bitrot-streaming.go:44-75)
 *   other algos    : raw shard bytes; sums[i*64..] receives the whole-file digest
 * Each files[i] must hold orc_bitrot_shard_file_size(orc_shard_file_size(...)) bytes.
 * Returns total bytes consumed or a negative error. */
int64_t orc_erasure_encode(int k, int m, int64_t block_size, int algo, const uint8_t *src,
                           int64_t len, uint8_t *const *files, uint8_t *sums);
/* bitrotVerify (cmd/bitrot.go:164) over one shard file */
int orc_bitrot_verify(int algo, const uint8_t *file, int64_t file_len, int64_t part_len,
                      int64_t shard_size, const uint8_t *want);
/* Erasure.Decode (erasure-decode.go:239): files as produced above; avail[i]==0 means the drive is
 * offline (nil reader).  Corrupt frames are detected via the digests and treated as missing
 * (corrupt_out[i] set).  Writes `length` bytes starting at object offset `offset` into dst. */
int64_t orc_erasure_decode(int k, int m, int64_t block_size, int algo, const uint8_t *const *files,
                           const uint8_t *avail, int64_t offset, int64_t length, int64_t total,
                           uint8_t *dst, uint8_t *corrupt_out);
/* Erasure.Heal (erasure-decode.go:317): rebuilds every shard file with stale[i]!=0 into
 * out_files[i] (full frames) from the available ones. */
int orc_erasure_heal(int k, int m, int64_t block_size, int algo, const uint8_t *const *files,
                     const uint8_t *avail, const uint8_t *stale, int64_t total,
                     uint8_t *const *out_files);

/* ---- SIMD + pthreads variant: the honest CPU baseline (stand-in for the klauspost /
 * highwayhash assembly + WithAutoGoroutines).  Same results as the scalar functions. ---- */
const char *orc_simd_level(void); /* "gfni-avx512" | "avx2" | "scalar" */
void orc_rs_encode_fast(int k, int m, uint8_t *const *shards, int64_t per);
void orc_hh256_fast(const uint8_t key[32], const uint8_t *p, size_t n, uint8_t out[32]);
/* encode + HighwayHash256S-frame nblocks full blocks of block_size with `threads` pthreads over
 * independent blocks (two passes per block, like the reference).  parity: nblocks*m*S bytes,
 * digests: nblocks*(k+m)*32.  Returns seconds spent. */
double orc_encode_hash_blocks_mt(int k, int m, int64_t block_size, const uint8_t *src,
                                 int64_t nblocks, uint8_t *parity, uint8_t *digests, int threads,
                                 int reps);
/* Persistent worker pool for the CPU arm (threads created once, pinned over all allowed CPUs, source first-touched by the
 * worker that encodes it): orc_pool_fill writes synthetic bytes into src (nblocks * block_size) and clears the outputs;
 * orc_pool_encode_hash runs `reps` passes of SIMD RS encode + HighwayHash of every shard and returns the seconds between
 * releasing the workers and the last one finishing. */
/* object checksums (internal/hash/checksum.go): finalized CRC values as Go's hash/crc32, hash/crc64 return them */
uint64_t orc_crc32(const uint8_t *p, size_t n);
uint64_t orc_crc32c(const uint8_t *p, size_t n);
uint64_t orc_crc64nvme(const uint8_t *p, size_t n);

void orc_encode_hash_blocks_st(int k, int m, int64_t block_size, const uint8_t *src, int64_t nblocks, uint8_t *parity, uint8_t *digests);
typedef struct orc_pool orc_pool;
orc_pool *orc_pool_new(int threads);
void orc_pool_fill(orc_pool *p, int k, int m, int64_t block_size, uint8_t *src, int64_t nblocks, uint8_t *parity, uint8_t *digests, uint64_t seed);
double orc_pool_encode_hash(orc_pool *p, int k, int m, int64_t block_size, const uint8_t *src, int64_t nblocks, uint8_t *parity,
                            uint8_t *digests, int reps);
void orc_pool_free(orc_pool *p);
void orc_pool_set_mode(int mode); /* 0 = encode + hash, 1 = encode only, 2 = hash only */

#ifdef __cplusplus
}
#endif
#endif
