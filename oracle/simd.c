/*
 * oracle/simd.c — CPU ORACLE, fast variant (test/bench infrastructure, not product code).
 * The honest CPU baseline: what klauspost/reedsolomon's amd64 assembly (pshufb nibble tables on
 * AVX2, vgf2p8affineqb on GFNI+AVX-512) + WithAutoGoroutines (cmd/erasure-coding.go:63) and
 * minio/highwayhash's AVX2 assembly do, written with intrinsics + pthreads over independent
 * erasure blocks.  Must produce exactly the bytes of rs.c / hh256.c (tests/test_oracle_simd.py).
 */
#define _GNU_SOURCE
#include <immintrin.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "oracle.h"

static int lvl = -1; /* 0 scalar, 1 avx2, 2 gfni-avx512 */
static int level(void) {
  if (lvl >= 0) return lvl;
  __builtin_cpu_init();
  lvl = 0;
  if (__builtin_cpu_supports("avx2")) lvl = 1;
  if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
      __builtin_cpu_supports("gfni"))
    lvl = 2;
  const char *e = getenv("ORC_SIMD");
  if (e) { int want = atoi(e); if (want < lvl) lvl = want; }
  return lvl;
}
const char *orc_simd_level(void) {
  switch (level()) { case 2: return "gfni-avx512"; case 1: return "avx2"; }
  return "scalar";
}

/* ---------------- RS encode ---------------- */
__attribute__((target("avx2"))) static void rs_rows_avx2(const uint8_t *rows, int k, int nrows,
                                                         uint8_t *const *in, uint8_t *const *out,
                                                         int64_t per) {
  /* per (row, col) nibble tables */
  uint8_t *tbl = (uint8_t *)aligned_alloc(32, (size_t)nrows * k * 32);
  for (int j = 0; j < nrows; j++)
    for (int c = 0; c < k; c++) {
      uint8_t f = rows[(size_t)j * k + c], *t = tbl + ((size_t)j * k + c) * 32;
      for (int v = 0; v < 16; v++) { t[v] = orc_gf_mul(f, (uint8_t)v); t[16 + v] = orc_gf_mul(f, (uint8_t)(v << 4)); }
    }
  const __m256i mask = _mm256_set1_epi8(0x0f);
  int64_t x = 0;
  for (; x + 32 <= per; x += 32) {
    for (int j0 = 0; j0 < nrows; j0 += 4) {
      int nj = nrows - j0 < 4 ? nrows - j0 : 4;
      __m256i acc[4] = {_mm256_setzero_si256(), _mm256_setzero_si256(), _mm256_setzero_si256(), _mm256_setzero_si256()};
      for (int c = 0; c < k; c++) {
        __m256i v = _mm256_loadu_si256((const __m256i *)(in[c] + x));
        __m256i lo = _mm256_and_si256(v, mask), hi = _mm256_and_si256(_mm256_srli_epi64(v, 4), mask);
        for (int j = 0; j < nj; j++) {
          const uint8_t *t = tbl + ((size_t)(j0 + j) * k + c) * 32;
          __m256i tl = _mm256_broadcastsi128_si256(_mm_load_si128((const __m128i *)t));
          __m256i th = _mm256_broadcastsi128_si256(_mm_load_si128((const __m128i *)(t + 16)));
          acc[j] = _mm256_xor_si256(acc[j], _mm256_xor_si256(_mm256_shuffle_epi8(tl, lo), _mm256_shuffle_epi8(th, hi)));
        }
      }
      for (int j = 0; j < nj; j++) _mm256_storeu_si256((__m256i *)(out[j0 + j] + x), acc[j]);
    }
  }
  for (; x < per; x++)
    for (int j = 0; j < nrows; j++) {
      uint8_t a = 0;
      for (int c = 0; c < k; c++) a ^= orc_gf_mul(rows[(size_t)j * k + c], in[c][x]);
      out[j][x] = a;
    }
  free(tbl);
}

static uint64_t gfni_matrix(uint8_t f) {
  /* vgf2p8affineqb: out bit i = parity(A.byte[7-i] & x) */
  uint64_t A = 0;
  for (int i = 0; i < 8; i++) {
    uint8_t row = 0;
    for (int b = 0; b < 8; b++)
      if ((orc_gf_mul(f, (uint8_t)(1u << b)) >> i) & 1) row |= (uint8_t)(1u << b);
    A |= (uint64_t)row << (8 * (7 - i));
  }
  return A;
}
__attribute__((target("avx512f,avx512bw,gfni"))) static void rs_rows_gfni(
    const uint8_t *rows, int k, int nrows, uint8_t *const *in, uint8_t *const *out, int64_t per) {
  uint64_t *A = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)nrows * k);
  for (int i = 0; i < nrows * k; i++) A[i] = gfni_matrix(rows[i]);
  int64_t x = 0;
  for (; x + 64 <= per; x += 64) {
    for (int j0 = 0; j0 < nrows; j0 += 8) {
      int nj = nrows - j0 < 8 ? nrows - j0 : 8;
      __m512i acc[8];
      for (int j = 0; j < nj; j++) acc[j] = _mm512_setzero_si512();
      for (int c = 0; c < k; c++) {
        __m512i v = _mm512_loadu_si512((const void *)(in[c] + x));
        for (int j = 0; j < nj; j++)
          acc[j] = _mm512_xor_si512(acc[j], _mm512_gf2p8affine_epi64_epi8(v, _mm512_set1_epi64((long long)A[(size_t)(j0 + j) * k + c]), 0));
      }
      for (int j = 0; j < nj; j++) _mm512_storeu_si512((void *)(out[j0 + j] + x), acc[j]);
    }
  }
  for (; x < per; x++)
    for (int j = 0; j < nrows; j++) {
      uint8_t a = 0;
      for (int c = 0; c < k; c++) a ^= orc_gf_mul(rows[(size_t)j * k + c], in[c][x]);
      out[j][x] = a;
    }
  free(A);
}

static void rs_rows(const uint8_t *rows, int k, int nrows, uint8_t *const *in, uint8_t *const *out, int64_t per) {
  if (level() == 2) rs_rows_gfni(rows, k, nrows, in, out, per);
  else if (level() == 1) rs_rows_avx2(rows, k, nrows, in, out, per);
  else {
    for (int64_t x = 0; x < per; x++)
      for (int j = 0; j < nrows; j++) {
        uint8_t a = 0;
        for (int c = 0; c < k; c++) a ^= orc_gf_mul(rows[(size_t)j * k + c], in[c][x]);
        out[j][x] = a;
      }
  }
}

void orc_rs_encode_fast(int k, int m, uint8_t *const *shards, int64_t per) {
  uint8_t *mat = (uint8_t *)malloc((size_t)(k + m) * k);
  orc_rs_matrix(k, m, mat);
  rs_rows(mat + (size_t)k * k, k, m, shards, shards + k, per);
  free(mat);
}

/* ---------------- HighwayHash-256, AVX2 body + scalar tail/finalisation ---------------- */
__attribute__((target("avx2"))) static void hh_body_avx2(orc_hh256_ctx *s, const uint8_t *p, size_t npackets) {
  __m256i v0 = _mm256_loadu_si256((const __m256i *)s->v0), v1 = _mm256_loadu_si256((const __m256i *)s->v1);
  __m256i m0 = _mm256_loadu_si256((const __m256i *)s->mul0), m1 = _mm256_loadu_si256((const __m256i *)s->mul1);
  const __m256i zip = _mm256_setr_epi8(3, 12, 2, 5, 14, 1, 15, 0, 11, 4, 10, 13, 9, 6, 8, 7,
                                       3, 12, 2, 5, 14, 1, 15, 0, 11, 4, 10, 13, 9, 6, 8, 7);
  for (size_t i = 0; i < npackets; i++, p += 32) {
    __m256i pk = _mm256_loadu_si256((const __m256i *)p);
    v1 = _mm256_add_epi64(v1, _mm256_add_epi64(m0, pk));
    m0 = _mm256_xor_si256(m0, _mm256_mul_epu32(v1, _mm256_srli_epi64(v0, 32)));
    v0 = _mm256_add_epi64(v0, m1);
    m1 = _mm256_xor_si256(m1, _mm256_mul_epu32(v0, _mm256_srli_epi64(v1, 32)));
    v0 = _mm256_add_epi64(v0, _mm256_shuffle_epi8(v1, zip));
    v1 = _mm256_add_epi64(v1, _mm256_shuffle_epi8(v0, zip));
  }
  _mm256_storeu_si256((__m256i *)s->v0, v0); _mm256_storeu_si256((__m256i *)s->v1, v1);
  _mm256_storeu_si256((__m256i *)s->mul0, m0); _mm256_storeu_si256((__m256i *)s->mul1, m1);
}

void orc_hh256_fast(const uint8_t key[32], const uint8_t *p, size_t n, uint8_t out[32]) {
  orc_hh256_ctx s;
  orc_hh256_init(&s, key);
  if (level() >= 1) {
    size_t np = n / 32;
    hh_body_avx2(&s, p, np);
    p += np * 32; n -= np * 32;
  }
  orc_hh256_write(&s, p, n);
  orc_hh256_sum(&s, out);
}

/* ---------------- multi-threaded encode + hash over independent blocks ---------------- */
static int g_mt_mode = 0; /* 0 = encode + hash, 1 = encode only, 2 = hash only (bench legs; set between runs, never during one) */
typedef struct {
  int k, m, tid, threads, reps;
  int64_t bs, nblocks;
  const uint8_t *src;
  uint8_t *parity, *digests;
  const uint8_t *rows;
} mt_arg;

static void *mt_worker(void *vp) {
  mt_arg *a = (mt_arg *)vp;
  int k = a->k, m = a->m, n = k + m;
  int64_t S = orc_shard_size(a->bs, k);
  uint8_t *last = (uint8_t *)aligned_alloc(64, (size_t)((S + 63) / 64 * 64));
  uint8_t **sh = (uint8_t **)malloc(sizeof(uint8_t *) * n);
  for (int r = 0; r < a->reps; r++)
    for (int64_t b = a->tid; b < a->nblocks; b += a->threads) {
      const uint8_t *blk = a->src + b * a->bs;
      for (int c = 0; c < k; c++) sh[c] = (uint8_t *)(blk + (size_t)c * S);
      int64_t tail = a->bs - (int64_t)(k - 1) * S; /* Split: zero-pad the last data shard */
      if (tail < S) {
        memcpy(last, blk + (size_t)(k - 1) * S, (size_t)tail);
        memset(last + tail, 0, (size_t)(S - tail));
        sh[k - 1] = last;
      }
      for (int j = 0; j < m; j++) sh[k + j] = a->parity + ((size_t)b * m + j) * S;
      if (g_mt_mode != 2) rs_rows(a->rows, k, m, sh, sh + k, S);   /* pass 1: Encode */
      if (g_mt_mode != 1)
        for (int i = 0; i < n; i++)                                  /* pass 2: bitrot hash per shard */
          orc_hh256_fast(orc_magic_hh_key, sh[i], (size_t)S, a->digests + ((size_t)b * n + i) * 32);
    }
  free(last); free(sh);
  return NULL;
}

/* single-threaded, on the calling thread (the per-request shape: every MinIO request goroutine encodes its own blocks) */
void orc_encode_hash_blocks_st(int k, int m, int64_t bs, const uint8_t *src, int64_t nblocks, uint8_t *parity, uint8_t *digests) {
  static uint8_t *mat = NULL;
  static int mk = 0, mm = 0;
  static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  pthread_mutex_lock(&mu);
  if (!mat || mk != k || mm != m) {
    uint8_t *nm = (uint8_t *)malloc((size_t)(k + m) * k);
    orc_rs_matrix(k, m, nm);
    orc_gf_mul(1, 1);
    mat = nm; mk = k; mm = m; /* the previous matrix (if any) is leaked on purpose: other threads may still read it */
  }
  const uint8_t *rows = mat + (size_t)k * k;
  pthread_mutex_unlock(&mu);
  mt_arg a = {k, m, 0, 1, 1, bs, nblocks, src, parity, digests, rows};
  mt_worker(&a);
}

double orc_encode_hash_blocks_mt(int k, int m, int64_t bs, const uint8_t *src, int64_t nblocks,
                                 uint8_t *parity, uint8_t *digests, int threads, int reps) {
  if (threads < 1) threads = 1;
  uint8_t *mat = (uint8_t *)malloc((size_t)(k + m) * k);
  orc_rs_matrix(k, m, mat);
  orc_gf_mul(1, 1); /* make sure the tables exist before threads start */
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
  mt_arg *args = (mt_arg *)malloc(sizeof(mt_arg) * threads);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < threads; t++) {
    args[t] = (mt_arg){k, m, t, threads, reps, bs, nblocks, src, parity, digests, mat + (size_t)k * k};
    pthread_create(&th[t], NULL, mt_worker, &args[t]);
  }
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(th); free(args); free(mat);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ---------------- persistent worker pool (the CPU arm of bench.py) ----------------
 * klauspost/reedsolomon keeps its goroutines for the life of the encoder (WithAutoGoroutines, cmd/erasure-coding.go:63) and
 * MinIO's request goroutines run on every core of both sockets.  The stand-in: T pthreads created ONCE (outside every timed
 * region), pinned round-robin over the CPUs the process may use, each first-touching the source blocks it will later
 * encode (so its pages are node-local), then released per run by a generation counter.  A run is timed from the moment
 * the workers are released to the moment the last one reports back. */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <sched.h>

struct orc_pool {
  int threads, stop;
  pthread_t *th;
  pthread_mutex_t mu;
  pthread_cond_t cv_go, cv_done;
  uint64_t gen;
  int pending;
  /* job */
  int job, k, m, reps;
  int64_t bs, nblocks;
  uint8_t *src, *parity, *digests;
  const uint8_t *rows;
  uint64_t seed;
  struct orc_pool_worker *w;
};
struct orc_pool_worker { struct orc_pool *p; int tid; };

static void pool_fill(struct orc_pool *p, int tid) {  /* job 1: first touch + synthetic bytes, same block deal as job 2 */
  for (int64_t b = tid; b < p->nblocks; b += p->threads) {
    uint64_t x = p->seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(b + 1));
    uint64_t *q = (uint64_t *)(p->src + b * p->bs);
    for (int64_t i = 0; i < p->bs / 8; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; q[i] = x; }
    if (p->parity) memset(p->parity + (size_t)b * p->m * orc_shard_size(p->bs, p->k), 0, (size_t)p->m * orc_shard_size(p->bs, p->k));
    if (p->digests) memset(p->digests + (size_t)b * (p->k + p->m) * 32, 0, (size_t)(p->k + p->m) * 32);
  }
}

static void *pool_main(void *vp) {
  struct orc_pool_worker *w = (struct orc_pool_worker *)vp;
  struct orc_pool *p = w->p;
  uint64_t seen = 0;
  for (;;) {
    pthread_mutex_lock(&p->mu);
    while (!p->stop && p->gen == seen) pthread_cond_wait(&p->cv_go, &p->mu);
    if (p->stop) { pthread_mutex_unlock(&p->mu); return NULL; }
    seen = p->gen;
    pthread_mutex_unlock(&p->mu);
    if (p->job == 1) {
      pool_fill(p, w->tid);
    } else {
      mt_arg a = {p->k, p->m, w->tid, p->threads, p->reps, p->bs, p->nblocks, p->src, p->parity, p->digests, p->rows};
      mt_worker(&a);
    }
    pthread_mutex_lock(&p->mu);
    if (--p->pending == 0) pthread_cond_signal(&p->cv_done);
    pthread_mutex_unlock(&p->mu);
  }
}

orc_pool *orc_pool_new(int threads) {
  if (threads < 1) threads = 1;
  orc_pool *p = (orc_pool *)calloc(1, sizeof(*p));
  p->threads = threads;
  p->th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
  p->w = (struct orc_pool_worker *)malloc(sizeof(struct orc_pool_worker) * threads);
  pthread_mutex_init(&p->mu, NULL);
  pthread_cond_init(&p->cv_go, NULL);
  pthread_cond_init(&p->cv_done, NULL);
  orc_gf_mul(1, 1); /* tables exist before any worker runs */
  cpu_set_t allowed;
  int cpus[CPU_SETSIZE], ncpu = 0;
  if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0)
    for (int c = 0; c < CPU_SETSIZE; c++)
      if (CPU_ISSET(c, &allowed)) cpus[ncpu++] = c;
  for (int t = 0; t < threads; t++) {
    p->w[t].p = p; p->w[t].tid = t;
    pthread_create(&p->th[t], NULL, pool_main, &p->w[t]);
    if (ncpu > 0) {  /* spread over everything the process may use — both sockets on a two-socket host */
      cpu_set_t one;
      CPU_ZERO(&one);
      CPU_SET(cpus[(int)(((int64_t)t * ncpu) / threads)], &one);
      pthread_setaffinity_np(p->th[t], sizeof(one), &one);
    }
  }
  return p;
}

static double pool_run(orc_pool *p) {
  struct timespec t0, t1;
  pthread_mutex_lock(&p->mu);
  p->pending = p->threads;
  p->gen++;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  pthread_cond_broadcast(&p->cv_go);
  while (p->pending > 0) pthread_cond_wait(&p->cv_done, &p->mu);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  pthread_mutex_unlock(&p->mu);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

void orc_pool_fill(orc_pool *p, int k, int m, int64_t bs, uint8_t *src, int64_t nblocks, uint8_t *parity, uint8_t *digests, uint64_t seed) {
  p->job = 1; p->k = k; p->m = m; p->bs = bs; p->src = src; p->nblocks = nblocks; p->parity = parity; p->digests = digests; p->seed = seed;
  pool_run(p);
}

double orc_pool_encode_hash(orc_pool *p, int k, int m, int64_t bs, const uint8_t *src, int64_t nblocks, uint8_t *parity,
                            uint8_t *digests, int reps) {
  uint8_t *mat = (uint8_t *)malloc((size_t)(k + m) * k);
  orc_rs_matrix(k, m, mat);
  p->job = 2; p->k = k; p->m = m; p->bs = bs; p->src = (uint8_t *)src; p->nblocks = nblocks; p->parity = parity; p->digests = digests;
  p->reps = reps; p->rows = mat + (size_t)k * k;
  double s = pool_run(p);
  free(mat);
  return s;
}

void orc_pool_free(orc_pool *p) {
  if (!p) return;
  pthread_mutex_lock(&p->mu);
  p->stop = 1;
  pthread_cond_broadcast(&p->cv_go);
  pthread_mutex_unlock(&p->mu);
  for (int t = 0; t < p->threads; t++) pthread_join(p->th[t], NULL);
  free(p->th); free(p->w); free(p);
}

/* which passes the workers run: 0 = RS encode + HighwayHash (the reference's work), 1 = encode only, 2 = hash only */
void orc_pool_set_mode(int mode) { g_mt_mode = mode; }
