// ec_api.cu — the C ABI of include/minio_ec.h on top of the fused kernel engine.
//
// Host-buffer entry points stage through device memory owned by the codec with a 3-slot
// H2D -> kernel -> D2H pipeline; *_device entry points run on caller-provided device memory.
// There is NO CPU fallback anywhere in this file: without a usable GPU every call fails.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>
#include "../../include/minio_ec.h"
#include "ec_engine.h"
#include "whole_hash.cuh"
#include "crc.cuh"
#include "batch_copy.cuh"
#include <nvtx3/nvToolsExt.h>

struct NvtxRange {
  explicit NvtxRange(const char* n) { nvtxRangePushA(n); }
  ~NvtxRange() { nvtxRangePop(); }
};

using namespace mec;

static const uint8_t kMagicKey[32] = {  // cmd/bitrot.go:37
    0x4b, 0xe7, 0x34, 0xfa, 0x8e, 0x23, 0x8a, 0xcd, 0x26, 0x3e, 0x83, 0xe6, 0xbb, 0x96, 0x85, 0x52,
    0x04, 0x0f, 0x93, 0x5d, 0xa3, 0x9f, 0x44, 0x14, 0x97, 0xe0, 0x9d, 0x13, 0x22, 0xde, 0x36, 0xa0};

static inline int64_t ceil_frac(int64_t num, int64_t den) {  // cmd/utils.go:689
  if (den == 0) return 0;
  if (den < 0) { num = -num; den = -den; }
  int64_t c = num / den;
  if (num > 0 && num % den != 0) c++;
  return c;
}
static inline int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

constexpr int kSlots = 6;

// grow-only page-locked host staging (digests and digest verdicts: small, read by the host right after a stream drains)
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t n) {
    if (n <= cap) return MEC_OK;
    release();
    const size_t want = n + (n >> 2) + 4096;
    MEC_CUDA_OK(cudaHostAlloc(&p, want, cudaHostAllocPortable));
    cap = want;
    return MEC_OK;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

struct Slot {
  DevBuf src, out, dig, aux, flags;  // aux: survivor arena of a reconstruct chunk / shard files of the whole-file path
  DevBuf crc_part, crc_out;          // object checksums of an encode chunk (crc.cuh)
  DevBuf dreq;                       // request table of a merged batch (batch_copy.cuh)
  DevBuf blk;                        // per-block geometry of a merged launch (SmallBlock / per-frame lengths)
  PinBuf hdig, hflags, hcrc, hreq, hblk;
  cudaStream_t st = nullptr;
};

struct mec_codec {
  int k = 0, m = 0, n = 0;
  int64_t block_size = 0;
  int algo = 0, device = 0;
  std::vector<uint8_t> matrix;  // n x k
  std::unique_ptr<Engine> eng;
  EngineOptions opt;
  std::mutex mu;
  Slot slots[kSlots];
  DevBuf flags, crc_tables;
  std::unique_ptr<CrcTables> crc_host;  // host copy of the CRC tables (combine across chunks)
  int checksums = 0;                    // option "checksums": MEC_CRC32 | MEC_CRC32C | MEC_CRC64NVME computed by the host-buffer encode calls
  uint64_t last_crc[kCrcTypes] = {0, 0, 0};
  int64_t last_crc_len = 0;
  // boundary counters (mec_get_stat)
  int64_t st_blocks_encoded = 0, st_blocks_read = 0, st_shards_rebuilt = 0, st_corrupt = 0, st_h2d = 0, st_d2h = 0;
  int64_t S() const { return ceil_frac(block_size, k); }
};

// ------------------------------------------------------------------------------------------------
extern "C" int mec_device_count(void) {
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) return 0;
  return c;
}
extern "C" const char* mec_last_error(void) { return get_last_error(); }
extern "C" const char* mec_version(void) { return "minio_b200 0.1 (sm_100a; fused RS+HighwayHash256)"; }

extern "C" int64_t mec_ceil_frac(int64_t a, int64_t b) { return ceil_frac(a, b); }
extern "C" int64_t mec_shard_size(const mec_codec* c) { return c->S(); }
extern "C" int64_t mec_shard_file_size(const mec_codec* c, int64_t total) {
  if (total == 0) return 0;
  if (total == -1) return -1;
  const int64_t num = total / c->block_size, last = total % c->block_size;
  return num * c->S() + ceil_frac(last, c->k);
}
extern "C" int64_t mec_shard_file_offset(const mec_codec* c, int64_t start, int64_t length, int64_t total) {
  const int64_t ss = c->S(), sfs = mec_shard_file_size(c, total);
  const int64_t end_shard = (start + length) / c->block_size;
  return std::min(end_shard * ss + ss, sfs);
}
extern "C" int64_t mec_bitrot_shard_file_size(int64_t size, int64_t shard_size, int algo) {
  if (algo != MEC_HIGHWAYHASH256S) return size;
  return ceil_frac(size, shard_size) * 32 + size;
}

extern "C" int mec_codec_new(int k, int m, int64_t block_size, int algo, int device, mec_codec** out) {
  if (!out) return MEC_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (k <= 0 || m < 0) return MEC_ERR_INV_SHARD_NUM;   // cmd/erasure-coding.go:44-46
  if (k + m > 256) return MEC_ERR_MAX_SHARD_NUM;       // cmd/erasure-coding.go:48-50
  if (block_size <= 0) return MEC_ERR_INVALID_ARGUMENT;
  if (algo < MEC_SHA256 || algo > MEC_BLAKE2B512) return MEC_ERR_INVALID_ARGUMENT;
  std::unique_ptr<mec_codec> c(new mec_codec);
  c->k = k; c->m = m; c->n = k + m; c->block_size = block_size; c->algo = algo; c->device = device;
  c->matrix.resize(static_cast<size_t>(c->n) * k);
  if (!rs_coding_matrix(k, m, c->matrix.data())) return MEC_ERR_INV_SHARD_NUM;
  if (const char* e = getenv("MEC_EB")) c->opt.eb = atoi(e);
  if (const char* e = getenv("MEC_FORCE_BYTEWISE")) c->opt.force_bytewise = atoi(e);
  if (const char* e = getenv("MEC_FORCE_DYNAMIC")) c->opt.force_dynamic = atoi(e);
  if (const char* e = getenv("MEC_GRID_MULT")) c->opt.grid_mult = atoi(e);
  if (const char* e = getenv("MEC_BALANCE_GRID")) c->opt.balance_grid = atoi(e);
  if (const char* e = getenv("MEC_CHUNK_BLOCKS")) c->opt.chunk_blocks = atoll(e);
  if (const char* e = getenv("MEC_USE_AUTO")) c->opt.use_auto = atoi(e);
  if (const char* e = getenv("MEC_JIT")) c->opt.jit = atoi(e);
  if (const char* e = getenv("MEC_NO_ROWS3D")) c->opt.no_rows3d = atoi(e);
  if (const char* e = getenv("MEC_SMALL_BLOCKS")) c->opt.small_blocks = atoll(e);
  if (const char* e = getenv("MEC_STATIC_GROUPS")) c->opt.static_groups = atoi(e);
  *out = c.release();
  return MEC_OK;
}

// The encoder is created on first use (like the sync.Once in cmd/erasure-coding.go:59-71); without a
// usable GPU this fails loudly — there is no CPU path behind this ABI.
static int ensure_engine(mec_codec* c) {
  if (c->eng) return MEC_OK;
  std::unique_ptr<Engine> e(new Engine(c->device));
  int rc = e->init();
  if (rc) return rc;
  for (auto& s : c->slots)
    if (cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking) != cudaSuccess) {
      set_last_error(std::string("cudaStreamCreate: ") + cudaGetErrorString(cudaGetLastError()));
      for (auto& q : c->slots)
        if (q.st) { cudaStreamDestroy(q.st); q.st = nullptr; }
      return MEC_ERR_CUDA;
    }

  c->eng = std::move(e);
  return MEC_OK;
}

extern "C" void mec_codec_free(mec_codec* c) {
  if (!c) return;
  if (!c->eng) { delete c; return; }
  cudaSetDevice(c->device);
  for (auto& s : c->slots) {
    if (s.st) { cudaStreamSynchronize(s.st); cudaStreamDestroy(s.st); }
    s.src.release(); s.out.release(); s.dig.release(); s.aux.release(); s.flags.release();
    s.hdig.release(); s.hflags.release(); s.hcrc.release(); s.crc_part.release(); s.crc_out.release(); s.dreq.release(); s.hreq.release(); s.blk.release(); s.hblk.release();
  }
  c->flags.release();
  c->crc_tables.release();
  delete c;
}

extern "C" int mec_set_option(mec_codec* c, const char* name, int64_t v) {
  if (!c || !name) return MEC_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!strcmp(name, "eb")) c->opt.eb = static_cast<int>(v);
  else if (!strcmp(name, "force_bytewise")) c->opt.force_bytewise = static_cast<int>(v);
  else if (!strcmp(name, "force_dynamic")) c->opt.force_dynamic = static_cast<int>(v);
  else if (!strcmp(name, "grid_mult")) c->opt.grid_mult = static_cast<int>(v);
  else if (!strcmp(name, "balance_grid")) c->opt.balance_grid = static_cast<int>(v);
  else if (!strcmp(name, "use_auto")) c->opt.use_auto = static_cast<int>(v);
  else if (!strcmp(name, "jit")) c->opt.jit = static_cast<int>(v);
  else if (!strcmp(name, "no_rows3d")) c->opt.no_rows3d = static_cast<int>(v);
  else if (!strcmp(name, "chunk_blocks")) c->opt.chunk_blocks = v;
  else if (!strcmp(name, "static_groups")) c->opt.static_groups = static_cast<int>(v);
  else if (!strcmp(name, "small_blocks")) c->opt.small_blocks = v;
  else if (!strcmp(name, "checksums")) c->checksums = static_cast<int>(v) & 7;
  else return MEC_ERR_INVALID_ARGUMENT;
  return MEC_OK;
}
extern "C" int64_t mec_get_stat(const mec_codec* c, const char* name) {
  if (!c || !name) return -1;
  if (!strcmp(name, "launches")) return c->eng ? c->eng->launches() : 0;
  if (!strcmp(name, "blocks_encoded")) return c->st_blocks_encoded;
  if (!strcmp(name, "blocks_read")) return c->st_blocks_read;
  if (!strcmp(name, "shards_rebuilt")) return c->st_shards_rebuilt;
  if (!strcmp(name, "corrupt_shards")) return c->st_corrupt;
  if (!strcmp(name, "bytes_h2d")) return c->st_h2d;
  if (!strcmp(name, "bytes_d2h")) return c->st_d2h;
  if (!strcmp(name, "jit_compiles")) return c->eng ? c->eng->jit_compiles() : 0;
  if (!strcmp(name, "jit_launches")) return c->eng ? c->eng->jit_launches() : 0;
  if (!strcmp(name, "small_launches")) return c->eng ? c->eng->small_launches() : 0;
  if (!strcmp(name, "jit_ms")) return c->eng ? static_cast<int64_t>(c->eng->jit_seconds() * 1e3) : 0;
  if (!strcmp(name, "jit_disk_hits")) return c->eng ? c->eng->jit_disk_hits() : 0;
  return -1;
}
extern "C" void mec_shutdown(void) { mec::jit_shutdown(); }

extern "C" int64_t mec_jit_compile_check(int k, int r, const uint8_t* coef, int align, int eb, int rows3d, int hash_outputs) {
  if (k < 1 || k > kMaxK || r < 1 || r > kMaxR || !coef || align < 0 || align > 15 || eb < 0) return MEC_ERR_INVALID_ARGUMENT;
  return mec::jit_compile_check(k, r, coef, align, eb, rows3d != 0, hash_outputs != 0);
}

// Pre-warm the kernel cache for the erasure patterns a drive failure will produce: every single missing data shard (the degraded
// GET of cmd/erasure-decode.go:239 with one drive gone, and the heal of that drive), both CTA shapes.  Compiles run on the
// background thread (and land in the on-disk cache), nothing waits for them.
extern "C" int mec_jit_prewarm(mec_codec* c) {
  if (!c) return MEC_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc = ensure_engine(c);
  if (rc) return rc;
  const int k = c->k, n = c->n;
  if (k > kMaxK || n > kMaxShards || c->m < 1) return 0;
  int queued = 0;
  for (int miss = 0; miss < k; miss++) {
    std::vector<uint8_t> present(n, 0), rows(k);
    int cnt = 0;
    for (int i = 0; i < n && cnt < k; i++)
      if (i != miss) { present[i] = 1; cnt++; }
    int valid[kMaxShards];
    if (!rs_decode_rows(k, c->m, present.data(), &miss, 1, rows.data(), valid)) continue;
    for (int eb_t : {4, 0})
      for (bool hash_out : {false, true}) {
        c->eng->jit_prewarm(k, 1, rows.data(), eb_t, hash_out);
        queued++;
      }
  }
  return queued;
}

extern "C" int64_t mec_launch_count(const mec_codec* c) { return (c && c->eng) ? c->eng->launches() : 0; }

static int require_streaming(mec_codec* c) {
  if (c->algo != MEC_HIGHWAYHASH256S) {
    set_last_error("only HighwayHash256S (streaming bitrot) is implemented on the GPU path");
    return MEC_ERR_UNSUPPORTED;
  }
  std::lock_guard<std::mutex> lk(c->mu);
  return ensure_engine(c);
}

// ------------------------------------------------------------------------------------------------
// object checksums (internal/hash/checksum.go, crc.go): CRC32 / CRC32C / CRC64NVME of a byte stream that is already in HBM
static int ensure_crc(mec_codec* c) {
  if (c->crc_host) return MEC_OK;
  std::unique_ptr<CrcTables> t(new CrcTables);
  crc_build_tables(t.get());
  int rc = c->crc_tables.ensure(sizeof(CrcTables));
  if (rc) return rc;
  MEC_CUDA_OK(cudaMemcpy(c->crc_tables.p, t.get(), sizeof(CrcTables), cudaMemcpyHostToDevice));
  c->crc_host = std::move(t);
  return MEC_OK;
}
// enqueue: CRCs of d_src[0, len) -> d_out3 (device, kCrcTypes words); `part` is scratch for the per-region partials
static int checksums_enqueue(mec_codec* c, const uint8_t* d_src, int64_t len, int which, DevBuf& part, uint64_t* d_out3, cudaStream_t st) {
  int rc = ensure_crc(c);
  if (rc) return rc;
  CrcParams p;
  p.src = d_src; p.len = len; p.which = which & 7;
  p.tables = static_cast<const CrcTables*>(c->crc_tables.p);
  p.nregions = (len + kCrcRegion - 1) / kCrcRegion;
  if ((rc = part.ensure(static_cast<size_t>(std::max<int64_t>(p.nregions, 1)) * kCrcTypes * sizeof(uint64_t)))) return rc;
  p.partial = static_cast<uint64_t*>(part.p);
  p.out = d_out3;
  if (p.nregions > 0) {
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>(p.nregions, static_cast<int64_t>(c->eng->num_sms()) * 8));
    crc_regions_kernel<<<grid, kCrcThreads, 0, st>>>(p);
    MEC_CUDA_OK(cudaGetLastError());
    c->eng->count_launch();
  }
  crc_fold_kernel<<<1, kCrcThreads, 0, st>>>(p);
  MEC_CUDA_OK(cudaGetLastError());
  c->eng->count_launch();
  return MEC_OK;
}

extern "C" uint64_t mec_checksum_combine(int type, uint64_t crc1, uint64_t crc2, int64_t len2) {
  const int ty = type == MEC_CRC32 ? 0 : (type == MEC_CRC32C ? 1 : (type == MEC_CRC64NVME ? 2 : -1));
  if (ty < 0) return 0;
  static const CrcTables* tabs = [] { CrcTables* t = new CrcTables; crc_build_tables(t); return t; }();
  const CrcSpec sp = crc_spec(ty);
  return crc_combine(tabs->xp8[ty], crc1, crc2, len2, sp.poly, sp.bits);  // Checksum.AddPart (internal/hash/crc.go:32-73)
}

extern "C" int mec_checksums_device(mec_codec* c, const uint8_t* d_src, int64_t len, int which, uint64_t* out3, void* stream) {
  if (!c || len < 0 || !out3 || (len > 0 && !d_src)) return MEC_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc = ensure_engine(c);
  if (rc) return rc;
  MEC_CUDA_OK(cudaSetDevice(c->device));
  Slot& s = c->slots[0];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((rc = s.crc_out.ensure(kCrcTypes * sizeof(uint64_t)))) return rc;
  if ((rc = checksums_enqueue(c, d_src, len, which, s.crc_part, static_cast<uint64_t*>(s.crc_out.p), st))) return rc;
  MEC_CUDA_OK(cudaMemcpyAsync(out3, s.crc_out.p, kCrcTypes * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  MEC_CUDA_OK(cudaStreamSynchronize(st));
  return MEC_OK;
}

extern "C" int mec_checksums(mec_codec* c, const uint8_t* src, int64_t len, int which, uint64_t* out3) {
  if (!c || len < 0 || !out3 || (len > 0 && !src)) return MEC_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc = ensure_engine(c);
  if (rc) return rc;
  if ((rc = ensure_crc(c))) return rc;
  MEC_CUDA_OK(cudaSetDevice(c->device));
  struct Drain { mec_codec* c; ~Drain() { for (auto& s : c->slots) if (s.st) cudaStreamSynchronize(s.st); } } drain{c};
  const int64_t chunk = 64ll << 20;
  uint64_t acc[kCrcTypes] = {0, 0, 0};
  int64_t done = 0;
  int64_t clen[kSlots] = {};
  bool busy[kSlots] = {};
  auto retire = [&](int si) {
    const uint64_t* h = static_cast<const uint64_t*>(c->slots[si].hcrc.p);
    for (int ty = 0; ty < kCrcTypes; ty++) {
      const CrcSpec sp = crc_spec(ty);
      acc[ty] = done == 0 ? h[ty] : crc_combine(c->crc_host->xp8[ty], acc[ty], h[ty], clen[si], sp.poly, sp.bits);
    }
    done += clen[si];
    busy[si] = false;
  };
  int si = 0;
  for (int64_t off = 0; off < len; off += chunk, si = (si + 1) % kSlots) {
    Slot& s = c->slots[si];
    MEC_CUDA_OK(cudaStreamSynchronize(s.st));
    if (busy[si]) retire(si);
    const int64_t n = std::min(chunk, len - off);
    if ((rc = s.src.ensure(static_cast<size_t>(round_up(n, 16) + 256)))) return rc;
    if ((rc = s.crc_out.ensure(kCrcTypes * sizeof(uint64_t)))) return rc;
    if ((rc = s.hcrc.ensure(kCrcTypes * sizeof(uint64_t)))) return rc;
    MEC_CUDA_OK(cudaMemcpyAsync(s.src.p, src + off, static_cast<size_t>(n), cudaMemcpyHostToDevice, s.st));
    if ((rc = checksums_enqueue(c, static_cast<const uint8_t*>(s.src.p), n, which, s.crc_part, static_cast<uint64_t*>(s.crc_out.p), s.st))) return rc;
    MEC_CUDA_OK(cudaMemcpyAsync(s.hcrc.p, s.crc_out.p, kCrcTypes * sizeof(uint64_t), cudaMemcpyDeviceToHost, s.st));
    clen[si] = n;
    busy[si] = true;
    c->st_h2d += n;
  }
  for (int q = 0; q < kSlots; q++, si = (si + 1) % kSlots) {
    MEC_CUDA_OK(cudaStreamSynchronize(c->slots[si].st));
    if (busy[si]) retire(si);
  }
  for (int ty = 0; ty < kCrcTypes; ty++) out3[ty] = (which & (1 << ty)) ? acc[ty] : 0;
  return MEC_OK;
}

extern "C" int64_t mec_last_checksums(const mec_codec* c, uint64_t* out3) {
  if (!c || !out3) return MEC_ERR_INVALID_ARGUMENT;
  for (int ty = 0; ty < kCrcTypes; ty++) out3[ty] = c->last_crc[ty];
  return c->last_crc_len;
}

// ------------------------------------------------------------------------------------------------
// encode
static int encode_device_locked(mec_codec* c, const uint8_t* d_src, int64_t len, uint8_t* d_parity, int64_t pitch,
                                uint8_t* d_digests, cudaStream_t st) {
  const int64_t bs = c->block_size, S = c->S();
  const int64_t nfull = len / bs, tail = len % bs;
  FusedDesc d;
  d.k = c->k; d.r = c->m;
  d.coef = c->matrix.data() + static_cast<size_t>(c->k) * c->k;
  d.static_encode = true;
  d.contiguous = true;
  d.key = kMagicKey;
  d.out_pitch = pitch;
  if (nfull > 0 && tail > 0 && c->eng->small_ok(c->opt, nfull + 1)) {
    // a small object with a short last block: ONE launch of the latency kernel, the last block carries its own geometry
    d.nblocks = nfull + 1; d.S = static_cast<int32_t>(S);
    d.in_base = d_src; d.in_block_stride = bs; d.in_block_len = bs;
    d.out = d_parity; d.digests = d_digests;
    d.tail_block = nfull; d.tail_in_off = nfull * bs; d.tail_S = static_cast<int32_t>(ceil_frac(tail, c->k)); d.tail_bytes = static_cast<int32_t>(tail);
    return c->eng->launch_fused(d, c->opt, st);
  }
  if (nfull > 0) {
    d.nblocks = nfull; d.S = static_cast<int32_t>(S);
    d.in_base = d_src; d.in_block_stride = bs; d.in_block_len = bs;
    d.out = d_parity; d.digests = d_digests;
    int rc = c->eng->launch_fused(d, c->opt, st);
    if (rc) return rc;
  }
  if (tail > 0) {
    const int64_t St = ceil_frac(tail, c->k);
    d.nblocks = 1; d.S = static_cast<int32_t>(St);
    d.in_base = d_src + nfull * bs; d.in_block_stride = round_up(tail, 16); d.in_block_len = tail;
    d.out = d_parity + nfull * c->m * pitch;
    d.digests = d_digests ? d_digests + nfull * c->n * 32 : nullptr;  // NULL = parity only, for the tail block too
    int rc = c->eng->launch_fused(d, c->opt, st);
    if (rc) return rc;
  }
  return MEC_OK;
}

extern "C" int mec_encode_blocks_device(mec_codec* c, const uint8_t* d_src, int64_t len, uint8_t* d_parity,
                                        int64_t parity_pitch, uint8_t* d_digests, void* stream) {
  NvtxRange nvtx("mec_encode_blocks_device");
  if (!c || len < 0) return MEC_ERR_INVALID_ARGUMENT;
  int rc;
  if (d_digests == nullptr) {  // parity only (whole-file bitrot algorithms hash separately, mec_whole_hash_device)
    std::lock_guard<std::mutex> lk0(c->mu);
    if ((rc = ensure_engine(c))) return rc;
  } else if ((rc = require_streaming(c))) {
    return rc;
  }
  if (len == 0) return MEC_OK;
  if (c->S() >= (1ll << 31)) return MEC_ERR_UNSUPPORTED;
  std::lock_guard<std::mutex> lk(c->mu);
  return encode_device_locked(c, d_src, len, d_parity, parity_pitch, d_digests, static_cast<cudaStream_t>(stream));
}

static int64_t pick_chunk_blocks(const mec_codec* c) {
  if (c->opt.chunk_blocks > 0) return c->opt.chunk_blocks;
  int64_t cb = (64ll << 20) / c->block_size;
  return cb < 1 ? 1 : cb;
}

// One chunk of the host-buffer encode pipeline, as the sinks below see it.
struct EncChunk {
  int64_t b0, nb;      // erasure blocks [b0, b0 + nb) of the call
  int64_t nfull;       // of which full (block_size) blocks; the call's short tail block, if in this chunk, follows them
  int64_t tail;        // bytes of that tail block (0 = none)
  Slot* s;             // slot whose stream carries the chunk
  int64_t pitch;       // parity row pitch
  // where the chunk lives: full blocks contiguous from d_src / d_out ((b*m + j)*pitch) / digests ([b][n][32]); the tail block
  // has its own addresses (right behind the full blocks in the single-caller pipeline, elsewhere in a merged batch)
  const uint8_t* d_src = nullptr; const uint8_t* d_src_tail = nullptr;
  uint8_t* d_out = nullptr;       uint8_t* d_out_tail = nullptr;
  const uint8_t* h_dig = nullptr; const uint8_t* h_dig_tail = nullptr;  // pinned host copies of the digests (frame sinks)
  cudaStream_t st = nullptr;
};

// The 3-slot H2D -> fused kernel -> D2H pipeline behind every host-buffer encode entry point.  `enqueue(ch)` adds the
// chunk's device->host copies to ch.s->st right behind the kernel; `retire(ch)` runs on the host once that stream has
// drained (digest scatter into frames).  On any error every slot stream is drained before returning, so no copy is
// still writing into caller memory (or reading freed staging) when the caller sees the error code.
template <class Enqueue, class Retire>
static int encode_pipeline(mec_codec* c, const uint8_t* src, int64_t len, Enqueue&& enqueue, Retire&& retire) {
  const int64_t bs = c->block_size, S = c->S(), pitch = round_up(S, 16);
  const int64_t nall = ceil_frac(len, bs);
  int64_t chunk = pick_chunk_blocks(c);
  // an object of a few chunks' worth or less is cut into ~four pieces so that its staging, kernel and copy-back overlap (a launch of
  // a handful of blocks costs the same 0.17 ms as one of a hundred: for 8-64 MiB objects the copies are what is worth overlapping)
  if (c->opt.chunk_blocks <= 0 && nall < 2 * chunk) chunk = std::max<int64_t>(4, ceil_frac(nall, 4));
  struct Drain {
    mec_codec* c;
    ~Drain() { for (auto& s : c->slots) if (s.st) cudaStreamSynchronize(s.st); }
  } drain{c};
  EncChunk inflight[kSlots];
  bool busy[kSlots] = {};
  int si = 0, rc;
  // option "checksums": CRCs of the object bytes ride along — computed on the chunk that was just staged for the encode (the
  // bytes cross PCIe once), merged across chunks on the host the way Checksum.AddPart merges parts
  const int ck = c->checksums;
  int64_t ck_len[kSlots] = {}, ck_done = 0;
  if (ck) {
    if ((rc = ensure_crc(c))) return rc;
    for (auto& v : c->last_crc) v = 0;
    c->last_crc_len = 0;
  }
  auto ck_retire = [&](int q) {
    if (!ck) return;
    const uint64_t* h = static_cast<const uint64_t*>(c->slots[q].hcrc.p);
    for (int ty = 0; ty < kCrcTypes; ty++) {
      if (!(ck & (1 << ty))) continue;
      const CrcSpec sp = crc_spec(ty);
      c->last_crc[ty] = ck_done == 0 ? h[ty] : crc_combine(c->crc_host->xp8[ty], c->last_crc[ty], h[ty], ck_len[q], sp.poly, sp.bits);
    }
    ck_done += ck_len[q];
    c->last_crc_len = ck_done;
  };
  for (int64_t b0 = 0; b0 < nall; b0 += chunk, si = (si + 1) % kSlots) {
    Slot& s = c->slots[si];
    MEC_CUDA_OK(cudaStreamSynchronize(s.st));
    if (busy[si]) { ck_retire(si); retire(inflight[si]); busy[si] = false; }
    EncChunk ch;
    ch.b0 = b0; ch.nb = std::min(chunk, nall - b0); ch.s = &s; ch.pitch = pitch;
    const int64_t off = b0 * bs, bytes = std::min(len - off, ch.nb * bs);
    ch.nfull = bytes / bs; ch.tail = bytes % bs;
    if ((rc = s.src.ensure(static_cast<size_t>(round_up(bytes, 16) + 256)))) return rc;
    if ((rc = s.out.ensure(static_cast<size_t>(ch.nb * std::max(c->m, 1) * pitch)))) return rc;
    if ((rc = s.dig.ensure(static_cast<size_t>(ch.nb * c->n * 32)))) return rc;
    MEC_CUDA_OK(cudaMemcpyAsync(s.src.p, src + off, static_cast<size_t>(bytes), cudaMemcpyHostToDevice, s.st));
    if ((rc = s.hdig.ensure(static_cast<size_t>(ch.nb * c->n * 32)))) return rc;
    ch.st = s.st;
    ch.d_src = static_cast<const uint8_t*>(s.src.p); ch.d_src_tail = ch.d_src + ch.nfull * bs;
    ch.d_out = static_cast<uint8_t*>(s.out.p); ch.d_out_tail = ch.d_out + ch.nfull * c->m * pitch;
    ch.h_dig = static_cast<const uint8_t*>(s.hdig.p); ch.h_dig_tail = ch.h_dig + ch.nfull * c->n * 32;
    if ((rc = encode_device_locked(c, static_cast<const uint8_t*>(s.src.p), bytes, static_cast<uint8_t*>(s.out.p), pitch,
                                   static_cast<uint8_t*>(s.dig.p), s.st)))
      return rc;
    if (ck) {
      if ((rc = s.crc_out.ensure(kCrcTypes * sizeof(uint64_t)))) return rc;
      if ((rc = s.hcrc.ensure(kCrcTypes * sizeof(uint64_t)))) return rc;
      if ((rc = checksums_enqueue(c, static_cast<const uint8_t*>(s.src.p), bytes, ck, s.crc_part, static_cast<uint64_t*>(s.crc_out.p), s.st))) return rc;
      MEC_CUDA_OK(cudaMemcpyAsync(s.hcrc.p, s.crc_out.p, kCrcTypes * sizeof(uint64_t), cudaMemcpyDeviceToHost, s.st));
      ck_len[si] = bytes;
    }
    if ((rc = enqueue(ch))) return rc;
    inflight[si] = ch;
    busy[si] = true;
  }
  for (int q = 0; q < kSlots; q++, si = (si + 1) % kSlots) {  // oldest first
    MEC_CUDA_OK(cudaStreamSynchronize(c->slots[si].st));
    if (busy[si]) { ck_retire(si); retire(inflight[si]); busy[si] = false; }
  }
  c->st_blocks_encoded += nall;
  c->st_h2d += len;
  return MEC_OK;
}

extern "C" int mec_encode_blocks(mec_codec* c, const uint8_t* src, int64_t len, uint8_t* parity, uint8_t* digests) {
  NvtxRange nvtx("mec_encode_blocks");
  if (!c || len < 0) return MEC_ERR_INVALID_ARGUMENT;
  int rc = require_streaming(c);
  if (rc) return rc;
  if (len == 0) return MEC_OK;  // cmd/erasure-coding.go:78-80
  if (c->S() >= (1ll << 31)) return MEC_ERR_UNSUPPORTED;
  std::lock_guard<std::mutex> lk(c->mu);
  MEC_CUDA_OK(cudaSetDevice(c->device));
  const int64_t S = c->S();
  rc = encode_pipeline(
      c, src, len,
      [&](const EncChunk& ch) -> int {
        Slot& s = *ch.s;
        if (c->m > 0 && ch.nfull > 0)
          MEC_CUDA_OK(cudaMemcpy2DAsync(parity + ch.b0 * c->m * S, static_cast<size_t>(S), s.out.p, static_cast<size_t>(ch.pitch),
                                        static_cast<size_t>(S), static_cast<size_t>(ch.nfull * c->m), cudaMemcpyDeviceToHost, s.st));
        if (c->m > 0 && ch.tail > 0) {
          const int64_t St = ceil_frac(ch.tail, c->k);
          MEC_CUDA_OK(cudaMemcpy2DAsync(parity + (ch.b0 + ch.nfull) * c->m * S, static_cast<size_t>(S),
                                        static_cast<uint8_t*>(s.out.p) + ch.nfull * c->m * ch.pitch, static_cast<size_t>(ch.pitch),
                                        static_cast<size_t>(St), static_cast<size_t>(c->m), cudaMemcpyDeviceToHost, s.st));
        }
        MEC_CUDA_OK(cudaMemcpyAsync(digests + ch.b0 * c->n * 32, s.dig.p, static_cast<size_t>(ch.nb * c->n * 32),
                                    cudaMemcpyDeviceToHost, s.st));
        return MEC_OK;
      },
      [](const EncChunk&) {});
  if (rc) return rc;
  c->st_d2h += ceil_frac(len, c->block_size) * (c->m * S + c->n * 32);
  return MEC_OK;
}

// a train of independent copies as one driver call (CUDA 12.8 cudaMemcpyBatchAsync); MEC_NO_BATCH_COPY=1: one call per copy
static int copy_batch(std::vector<void*>& dsts, std::vector<void*>& srcs, std::vector<size_t>& sizes, cudaStream_t st) {
  if (dsts.empty()) return MEC_OK;
  static const bool no_batch = getenv("MEC_NO_BATCH_COPY") != nullptr;
  if (!no_batch && dsts.size() > 1) {
    cudaMemcpyAttributes at = {};
    at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
    size_t idx0 = 0, fail = 0;
    const cudaError_t e = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), dsts.size(), &at, &idx0, 1, &fail, st);
    if (e == cudaSuccess) return MEC_OK;
    cudaGetLastError();  // older driver: fall through to separate calls
  }
  for (size_t i = 0; i < dsts.size(); i++) MEC_CUDA_OK(cudaMemcpyAsync(dsts[i], srcs[i], sizes[i], cudaMemcpyDefault, st));
  return MEC_OK;
}

struct CopyList {
  std::vector<void*> dsts, srcs;
  std::vector<size_t> sizes;
  void add(void* d, const void* s, int64_t n) {
    if (n <= 0) return;
    dsts.push_back(d); srcs.push_back(const_cast<void*>(s)); sizes.push_back(static_cast<size_t>(n));
  }
  int flush(cudaStream_t st) {
    const int rc = copy_batch(dsts, srcs, sizes, st);
    dsts.clear(); srcs.clear(); sizes.clear();
    return rc;
  }
};
// chunks of at most this many erasure blocks move as ONE batched copy call per direction instead of one 2-D copy per shard:
// a small GetObject is two driver calls instead of two dozen (the driver lock is what concurrent small requests queue on)
constexpr int64_t kBatchedCopyBlocks = 8;

// Frame sinks of Erasure.Encode.  files[i] is the part.N image of drive i: ([32 B digest][shard bytes])* with full frames
// of 32 + S bytes and one short last frame.  Everything that is bulk moves by DMA straight from device memory into the
// caller's frames (2-D copies: one row per erasure block); only the 32-byte digests go through a pinned staging buffer and
// are scattered by the host while the next chunks are in flight.  No pageable temporaries, no assembly memcpy of shard bytes.
//   parity drives:  frames of drive k+j come from the kernel's parity rows
//   data drives:    with_data = true copies the data shards back from the staged object bytes (the device already holds
//                   them); false leaves the data part to the caller (mec_encode_sg: the writer emits the digest followed
//                   by the slice of its own source buffer, as streamingBitrotWriter.Write does — no copy anywhere)
static int frames_enqueue(mec_codec* c, const EncChunk& ch, uint8_t* const* files, bool with_data, uint8_t* data_digests, CopyList* list = nullptr) {
  const int k = c->k, m = c->m;
  const int64_t bs = c->block_size, S = c->S(), fstride = 32 + S;
  cudaStream_t st = ch.st;
  const int64_t per_t = ch.tail > 0 ? ceil_frac(ch.tail, k) : 0;
  for (int j = 0; j < m; j++) {
    uint8_t* f = files[k + j];
    if (!f) continue;
    f += ch.b0 * fstride + 32;
    if (list) {
      for (int64_t b = 0; b < ch.nfull; b++) list->add(f + b * fstride, ch.d_out + (b * m + j) * ch.pitch, S);
      if (ch.tail > 0) list->add(f + ch.nfull * fstride, ch.d_out_tail + j * ch.pitch, per_t);
    } else {
      if (ch.nfull > 0)
        MEC_CUDA_OK(cudaMemcpy2DAsync(f, static_cast<size_t>(fstride), ch.d_out + j * ch.pitch,
                                      static_cast<size_t>(m * ch.pitch), static_cast<size_t>(S), static_cast<size_t>(ch.nfull),
                                      cudaMemcpyDeviceToHost, st));
      if (ch.tail > 0)
        MEC_CUDA_OK(cudaMemcpyAsync(f + ch.nfull * fstride, ch.d_out_tail + j * ch.pitch,
                                    static_cast<size_t>(per_t), cudaMemcpyDeviceToHost, st));
    }
    c->st_d2h += ch.nfull * S + per_t;
  }
  if (with_data) {
    for (int i = 0; i < k; i++) {
      uint8_t* f = files[i];
      if (!f) continue;
      f += ch.b0 * fstride + 32;
      const int64_t w = std::max<int64_t>(0, std::min(S, bs - static_cast<int64_t>(i) * S));  // Split: the last shard is short, the rest is zero padding
      if (list) {
        for (int64_t b = 0; b < ch.nfull; b++) list->add(f + b * fstride, ch.d_src + b * bs + static_cast<int64_t>(i) * S, w);
      } else if (ch.nfull > 0 && w > 0) {
        MEC_CUDA_OK(cudaMemcpy2DAsync(f, static_cast<size_t>(fstride), ch.d_src + static_cast<int64_t>(i) * S,
                                      static_cast<size_t>(bs), static_cast<size_t>(w), static_cast<size_t>(ch.nfull), cudaMemcpyDeviceToHost, st));
      }
      if (ch.tail > 0) {
        const int64_t start = static_cast<int64_t>(i) * per_t, have = std::max<int64_t>(0, std::min(per_t, ch.tail - start));
        if (list) list->add(f + ch.nfull * fstride, ch.d_src_tail + start, have);
        else if (have > 0)
          MEC_CUDA_OK(cudaMemcpyAsync(f + ch.nfull * fstride, ch.d_src_tail + start,
                                      static_cast<size_t>(have), cudaMemcpyDeviceToHost, st));
      }
      c->st_d2h += ch.nfull * w;
    }
  }
  (void)data_digests;
  return MEC_OK;
}

static void frames_retire(mec_codec* c, const EncChunk& ch, uint8_t* const* files, bool with_data, uint8_t* data_digests) {
  const int k = c->k, n = c->n;
  const int64_t bs = c->block_size, S = c->S(), fstride = 32 + S;
  const int64_t per_t = ch.tail > 0 ? ceil_frac(ch.tail, k) : 0;
  for (int64_t b = 0; b < ch.nb; b++) {
    const bool is_tail = b >= ch.nfull;
    for (int i = 0; i < n; i++) {
      const uint8_t* dg = is_tail ? ch.h_dig_tail + i * 32 : ch.h_dig + (b * n + i) * 32;
      if (data_digests && i < k) memcpy(data_digests + ((ch.b0 + b) * k + i) * 32, dg, 32);
      uint8_t* f = files ? files[i] : nullptr;
      if (!f || (i < k && !with_data)) continue;
      f += (ch.b0 + b) * fstride;
      memcpy(f, dg, 32);  // hash first, then the shard (cmd/bitrot-streaming.go:60,65)
      if (i < k) {        // Split's zero padding behind the object bytes (cmd/erasure-coding.go:81)
        const int64_t per = is_tail ? per_t : S, blen = is_tail ? ch.tail : bs;
        const int64_t have = std::max<int64_t>(0, std::min(per, blen - static_cast<int64_t>(i) * per));
        if (have < per) memset(f + 32 + have, 0, static_cast<size_t>(per - have));
      }
    }
  }
}

static int64_t encode_frames(mec_codec* c, const uint8_t* src, int64_t len, uint8_t* const* files, bool with_data,
                             uint8_t* data_digests, int write_quorum, const char* what) {
  NvtxRange nvtx(what);
  if (!c || len < 0 || !files) return MEC_ERR_INVALID_ARGUMENT;
  int online = 0;
  for (int i = 0; i < c->n; i++) online += files[i] != nullptr;
  if (online < write_quorum) return MEC_ERR_WRITE_QUORUM;  // cmd/erasure-encode.go:59-65
  if (len == 0) return 0;
  int rc = require_streaming(c);
  if (rc) return rc;
  if (c->S() >= (1ll << 31)) return MEC_ERR_UNSUPPORTED;
  std::lock_guard<std::mutex> lk(c->mu);
  MEC_CUDA_OK(cudaSetDevice(c->device));
  rc = encode_pipeline(
      c, src, len,
      [&](const EncChunk& ch) -> int {
        if (ch.nb <= kBatchedCopyBlocks) {  // a small object: digests and every frame piece in ONE batched copy call
          CopyList list;
          list.add(ch.s->hdig.p, ch.s->dig.p, ch.nb * c->n * 32);
          const int e = frames_enqueue(c, ch, files, with_data, data_digests, &list);
          return e ? e : list.flush(ch.st);
        }
        MEC_CUDA_OK(cudaMemcpyAsync(ch.s->hdig.p, ch.s->dig.p, static_cast<size_t>(ch.nb * c->n * 32), cudaMemcpyDeviceToHost, ch.st));
        return frames_enqueue(c, ch, files, with_data, data_digests);
      },
      [&](const EncChunk& ch) { frames_retire(c, ch, files, with_data, data_digests); });
  return rc ? rc : len;
}

// Erasure.Encode + streaming bitrot writers (cmd/erasure-encode.go:69, cmd/bitrot-streaming.go:44-75): complete part.N images.
extern "C" int64_t mec_encode(mec_codec* c, const uint8_t* src, int64_t len, uint8_t* const* files, int write_quorum) {
  return encode_frames(c, src, len, files, true, nullptr, write_quorum, "mec_encode");
}

// Scatter-gather form of the same call: the k data shards stay where Split left them — inside `src` — and only what the GPU
// produced comes back: complete frames for the parity drives (files[k..n), NULL = offline) and the digests of the data shards
// ([block][k][32]).  files[0..k) are only tested for NULL (offline writers count against the quorum).
extern "C" int64_t mec_encode_sg(mec_codec* c, const uint8_t* src, int64_t len, uint8_t* const* files, uint8_t* data_digests,
                                 int write_quorum) {
  if (!data_digests) return MEC_ERR_INVALID_ARGUMENT;
  return encode_frames(c, src, len, files, false, data_digests, write_quorum, "mec_encode_sg");
}

// ------------------------------------------------------------------------------------------------
// reconstruct over frames (host)
struct FrameGeom {
  int64_t nblocks, S, last_len;
  int64_t file_bytes() const { return nblocks <= 0 ? 0 : (nblocks - 1) * (32 + S) + 32 + last_len; }
  // frames are staged on the device at a 16-byte aligned pitch so that every TMA row is aligned
  int64_t dpitch() const { return round_up(32 + S, 16); }
  int64_t dev_bytes() const { return nblocks * dpitch() + 512; }
};

// host frames ([32+S] stride, last frame possibly short) -> device frames at pitch g.dpitch()
static int stage_frames(const uint8_t* host, void* dev, const FrameGeom& g, cudaStream_t st) {
  const int64_t nfull = (g.last_len == g.S) ? g.nblocks : g.nblocks - 1;
  if (nfull > 0)
    MEC_CUDA_OK(cudaMemcpy2DAsync(dev, static_cast<size_t>(g.dpitch()), host, static_cast<size_t>(32 + g.S),
                                  static_cast<size_t>(32 + g.S), static_cast<size_t>(nfull), cudaMemcpyHostToDevice, st));
  if (nfull < g.nblocks)
    MEC_CUDA_OK(cudaMemcpyAsync(static_cast<uint8_t*>(dev) + nfull * g.dpitch(), host + nfull * (32 + g.S),
                                static_cast<size_t>(32 + g.last_len), cudaMemcpyHostToDevice, st));
  return MEC_OK;
}

// Where the rebuilt / read shards of a range go.
//   frame sink  (Heal, mec_reconstruct_frames): out[i] = frame-layout image of shard i, same geometry as the inputs
//   object sink (Decode): the data shards of every block, cut to the object bytes [lo, hi) and laid end to end in dst —
//                writeDataBlocks (cmd/erasure-utils.go:42) done by the copy engines: one 2-D copy per data shard and chunk
struct RangeSink {
  uint8_t* const* out = nullptr;
  uint8_t* dst = nullptr;
  int64_t first_block = 0;   // index (within the part) of the range's first block
  int64_t lo = 0, hi = 0;    // wanted part bytes
  int64_t bs = 0, total = 0; // block size, part size
};

// One chunk of the reconstruct pipeline: blocks [b0, b0 + nb) of the range, read through reader set `chosen`.
struct RChunk {
  int64_t b0 = 0, nb = 0;
  int chosen[kMaxK];
  int targets[kMaxR];
  int r = 0;
  uint64_t epoch = 0;  // reader-set generation the chunk was submitted under
};

// survivors of a chunk sit in ONE arena per slot at a uniform stride (position t of the reader set at t * stride)
static inline uint8_t* arena_ptr(Slot& s, int64_t stride, int t) { return static_cast<uint8_t*>(s.aux.p) + t * stride; }

// Rebuild `ch.targets` for the chunk from the staged survivor frames.  d_out rows: (b*r + q); digests [b][k + r]; flags [b][k].
static int launch_reconstruct(mec_codec* c, const FrameGeom& g, const RChunk& ch, Slot& s, int64_t stride, const uint8_t* rows,
                              int64_t pitch, bool hash_outputs) {
  const int k = c->k, r = ch.r;
  FusedDesc d;
  d.k = k; d.r = r; d.coef = rows; d.static_encode = false; d.contiguous = false; d.hash_outputs = hash_outputs;
  d.key = kMagicKey; d.out_pitch = pitch; d.expect_block_stride = g.dpitch(); d.in_block_stride = g.dpitch();
  // the range's short last block (if this chunk holds it) is a second launch with its own shard length
  const bool has_short = (ch.b0 + ch.nb == g.nblocks) && g.last_len != g.S;
  const int64_t nfull = has_short ? ch.nb - 1 : ch.nb;
  if (has_short && nfull > 0 && c->eng->small_ok(c->opt, ch.nb)) {  // one launch: the last block carries its own shard length
    d.nblocks = ch.nb;
    d.S = static_cast<int32_t>(g.S);
    d.tail_block = nfull; d.tail_S = static_cast<int32_t>(g.last_len);
    for (int t = 0; t < k; t++) {
      const uint8_t* base = arena_ptr(s, stride, t);
      d.map_base[t] = base; d.map_len[t] = stride; d.expect_ptr[t] = base; d.in_ptr[t] = base + 32;
    }
    d.out = static_cast<uint8_t*>(s.out.p);
    d.digests = static_cast<uint8_t*>(s.dig.p);
    d.corrupt = static_cast<uint8_t*>(s.flags.p);
    return c->eng->launch_fused(d, c->opt, s.st);
  }
  for (int pass = 0; pass < 2; pass++) {
    const int64_t first = pass == 0 ? 0 : nfull;
    const int64_t nb = pass == 0 ? nfull : ch.nb - nfull;
    if (nb <= 0) continue;
    d.nblocks = nb;
    d.S = static_cast<int32_t>(pass == 0 ? g.S : g.last_len);
    for (int t = 0; t < k; t++) {
      const uint8_t* base = arena_ptr(s, stride, t) + first * g.dpitch();
      d.map_base[t] = base;
      d.map_len[t] = stride - first * g.dpitch();
      d.expect_ptr[t] = base;
      d.in_ptr[t] = base + 32;
    }
    d.out = static_cast<uint8_t*>(s.out.p) + first * r * pitch;
    d.digests = static_cast<uint8_t*>(s.dig.p) + first * (k + r) * 32;
    d.corrupt = static_cast<uint8_t*>(s.flags.p) + first * k;
    int rc = c->eng->launch_fused(d, c->opt, s.st);
    if (rc) return rc;
  }
  return MEC_OK;
}

// parallelReader.Read's choice of readers (cmd/erasure-decode.go:145-221): the first k alive readers in index order,
// preferred ones first (preferReaders, :92-123)
static int choose_readers(const mec_codec* c, const uint8_t* const* frames, const uint8_t* alive, const uint8_t* prefer, int* chosen) {
  const int k = c->k, n = c->n;
  int nch = 0;
  if (prefer) {
    int order[kMaxShards], next = 0;
    for (int i = 0; i < n; i++) order[i] = i;
    for (int i = 0; i < n; i++) {
      if (!prefer[i] || frames[i] == nullptr) continue;
      if (i == next) { next++; continue; }
      std::swap(order[next], order[i]);
      next++;
    }
    for (int q = 0; q < n && nch < k; q++)
      if (alive[order[q]]) chosen[nch++] = order[q];
    std::sort(chosen, chosen + nch);  // decode rows are defined on ascending shard indices
  } else {
    for (int i = 0; i < n && nch < k; i++)
      if (alive[i]) chosen[nch++] = i;
  }
  return nch;
}

// device->host copies of one verified-or-not chunk into the object sink (optimistic: a chunk that turns out to hold a corrupt
// frame is redone from the bad block with other readers and its bytes are overwritten before the call returns)
static int emit_object(mec_codec* c, const FrameGeom& g, const RangeSink& sk, const RChunk& ch, Slot& s, int64_t stride, int64_t pitch,
                       CopyList* list = nullptr) {
  const int k = c->k, r = ch.r;
  const int64_t bs = sk.bs, S = g.S, P = g.dpitch();
  for (int i = 0; i < k; i++) {
    // device source of data shard i: a staged survivor frame or a rebuilt row
    const uint8_t* base = nullptr;
    int64_t spitch = 0;
    for (int t = 0; t < k && !base; t++)
      if (ch.chosen[t] == i) { base = arena_ptr(s, stride, t) + 32; spitch = P; }
    for (int q = 0; q < r && !base; q++)
      if (ch.targets[q] == i) { base = static_cast<uint8_t*>(s.out.p) + q * pitch; spitch = r * pitch; }
    if (!base) return MEC_ERR_UNEXPECTED;
    // blocks of the chunk: a run of blocks that are wanted in full goes out as one 2-D copy, partial ones one by one
    int64_t run0 = -1;
    auto flush = [&](int64_t run1) -> int {  // blocks [run0, run1) of the chunk, all full and fully wanted
      if (run0 < 0 || run1 <= run0) { run0 = -1; return MEC_OK; }
      const int64_t w = std::max<int64_t>(0, std::min(S, bs - static_cast<int64_t>(i) * S));
      if (w > 0) {
        const int64_t B = sk.first_block + ch.b0 + run0;
        if (list) {
          for (int64_t q = 0; q < run1 - run0; q++)
            list->add(sk.dst + ((B + q) * bs + static_cast<int64_t>(i) * S - sk.lo), base + (run0 + q) * spitch, w);
        } else {
          MEC_CUDA_OK(cudaMemcpy2DAsync(sk.dst + (B * bs + static_cast<int64_t>(i) * S - sk.lo), static_cast<size_t>(bs), base + run0 * spitch,
                                        static_cast<size_t>(spitch), static_cast<size_t>(w), static_cast<size_t>(run1 - run0), cudaMemcpyDeviceToHost, s.st));
        }
        c->st_d2h += w * (run1 - run0);
      }
      run0 = -1;
      return MEC_OK;
    };
    int rc;
    for (int64_t b = 0; b < ch.nb; b++) {
      const int64_t B = sk.first_block + ch.b0 + b;
      const int64_t blo = B * bs, bhi = std::min(blo + bs, sk.total);          // part bytes of this block
      const int64_t cur = (ch.b0 + b == g.nblocks - 1) ? g.last_len : S;       // its shard length
      const bool whole = blo >= sk.lo && bhi <= sk.hi && bhi - blo == bs && cur == S;
      if (whole) { if (run0 < 0) run0 = b; continue; }
      if ((rc = flush(b))) return rc;
      const int64_t slo = blo + static_cast<int64_t>(i) * cur, shi = std::min(slo + cur, bhi);
      const int64_t a = std::max(slo, sk.lo), e = std::min(shi, sk.hi);
      if (e > a) {
        if (list) list->add(sk.dst + (a - sk.lo), base + b * spitch + (a - slo), e - a);
        else MEC_CUDA_OK(cudaMemcpyAsync(sk.dst + (a - sk.lo), base + b * spitch + (a - slo), static_cast<size_t>(e - a), cudaMemcpyDeviceToHost, s.st));
        c->st_d2h += e - a;
      }
    }
    if ((rc = flush(ch.nb))) return rc;
  }
  return list ? list->flush(s.st) : MEC_OK;
}

// rebuilt shards of a chunk -> frame-layout outputs (digest + shard per block)
static int emit_frames(mec_codec* c, const FrameGeom& g, const RangeSink& sk, const RChunk& ch, Slot& s, int64_t pitch, CopyList* list = nullptr) {
  const int k = c->k, r = ch.r;
  const int64_t fstride = 32 + g.S;
  const bool has_short = (ch.b0 + ch.nb == g.nblocks) && g.last_len != g.S;
  const int64_t nfull = has_short ? ch.nb - 1 : ch.nb;
  for (int q = 0; q < r; q++) {
    uint8_t* dst = sk.out[ch.targets[q]];
    if (!dst) continue;
    dst += ch.b0 * fstride;
    if (list) {
      for (int64_t b = 0; b < ch.nb; b++) {
        list->add(dst + b * fstride, static_cast<uint8_t*>(s.dig.p) + (b * (k + r) + k + q) * 32, 32);
        list->add(dst + b * fstride + 32, static_cast<uint8_t*>(s.out.p) + (b * r + q) * pitch, (has_short && b == nfull) ? g.last_len : g.S);
      }
      c->st_d2h += nfull * (32 + g.S) + (has_short ? 32 + g.last_len : 0);
      continue;
    }
    if (nfull > 0) {
      MEC_CUDA_OK(cudaMemcpy2DAsync(dst, static_cast<size_t>(fstride), static_cast<uint8_t*>(s.dig.p) + (k + q) * 32,
                                    static_cast<size_t>((k + r) * 32), 32, static_cast<size_t>(nfull), cudaMemcpyDeviceToHost, s.st));
      MEC_CUDA_OK(cudaMemcpy2DAsync(dst + 32, static_cast<size_t>(fstride), static_cast<uint8_t*>(s.out.p) + q * pitch,
                                    static_cast<size_t>(r * pitch), static_cast<size_t>(g.S), static_cast<size_t>(nfull), cudaMemcpyDeviceToHost, s.st));
    }
    if (has_short) {
      const int64_t b = nfull;
      MEC_CUDA_OK(cudaMemcpyAsync(dst + b * fstride, static_cast<uint8_t*>(s.dig.p) + (b * (k + r) + k + q) * 32, 32, cudaMemcpyDeviceToHost, s.st));
      MEC_CUDA_OK(cudaMemcpyAsync(dst + b * fstride + 32, static_cast<uint8_t*>(s.out.p) + (b * r + q) * pitch, static_cast<size_t>(g.last_len),
                                  cudaMemcpyDeviceToHost, s.st));
    }
    c->st_d2h += nfull * (32 + g.S) + (has_short ? 32 + g.last_len : 0);
  }
  return list ? list->flush(s.st) : MEC_OK;
}

// Core of Decode / Heal: frames[i] point at the first frame of the range (host).  The range is cut into chunks that run through
// the codec's three slots — while chunk c is in the kernel, chunk c+1 is being staged and chunk c-1 is on its way back — each
// slot with its own stream, survivor arena, output and flag buffers.  Chunks retire in order: the per-frame digest verdicts
// ([block][reader] flags written by the hash threads) are read, and on the first mismatch the failing readers are dropped
// for the rest of the call (parallelReader.Read: p.readers[i] = nil, cmd/erasure-decode.go:196-199), everything submitted
// behind the bad block is discarded and the range resumes AT the bad block with the next readers in order — exactly the
// block-sequential fail-over of the reference, without giving up the pipelining in the common clean case.
static int reconstruct_range(mec_codec* c, const uint8_t* const* frames, const FrameGeom& g, const uint8_t* want,
                             int data_only, const RangeSink& sk, uint8_t* corrupt, uint8_t* alive /*n, in/out*/,
                             const uint8_t* prefer = nullptr, bool hash_outputs = true) {
  NvtxRange nvtx("mec_reconstruct_range");
  const int k = c->k, n = c->n;
  if (k > kMaxK) return MEC_ERR_UNSUPPORTED;
  MEC_CUDA_OK(cudaSetDevice(c->device));
  struct Drain {
    mec_codec* c;
    ~Drain() { for (auto& s : c->slots) if (s.st) cudaStreamSynchronize(s.st); }
  } drain{c};
  const int64_t fstride = 32 + g.S, P = g.dpitch(), pitch = round_up(g.S, 16);
  // chunk: ~32 MiB of object per chunk keeps short ranges pipelined over several slots; long ranges take larger chunks (up to four
  // times that) because every chunk costs a host round trip when it retires (measured: 34.3 -> 38.4 GiB/s on a 2 GiB GetObject)
  int64_t chunk = std::max<int64_t>(1, (32ll << 20) / std::max<int64_t>(1, g.S * k));
  chunk = std::max(chunk, std::min(4 * chunk, g.nblocks / 12));
  if (g.nblocks < 2 * chunk) chunk = std::max<int64_t>(4, ceil_frac(g.nblocks, 4));  // short ranges: ~four pieces, copies overlap (see encode_pipeline)
  if (c->opt.chunk_blocks > 0) chunk = c->opt.chunk_blocks;
  chunk = std::min(chunk, g.nblocks);
  const int64_t stride = round_up(chunk * P + 512, 256);

  RChunk inflight[kSlots];
  std::vector<uint8_t> rows_of[kSlots];
  bool busy[kSlots] = {};
  int head = 0, tail = 0, nbusy = 0;  // ring: tail = oldest in flight, head = next free slot
  int64_t next = 0, done = 0;         // next block to submit, blocks accepted so far
  uint64_t epoch = 0;
  int rc;

  auto submit = [&](int si, int64_t b0) -> int {
    Slot& s = c->slots[si];
    RChunk& ch = inflight[si];
    ch.b0 = b0; ch.nb = std::min(chunk, g.nblocks - b0); ch.epoch = epoch;
    if (choose_readers(c, frames, alive, prefer, ch.chosen) < k) return MEC_ERR_READ_QUORUM;
    std::vector<uint8_t> present(n, 0);
    for (int t = 0; t < k; t++) present[ch.chosen[t]] = 1;
    ch.r = 0;
    for (int i = 0; i < n; i++)
      if (want[i] && !present[i] && !(data_only && i >= k)) {
        if (ch.r >= kMaxR) return MEC_ERR_UNSUPPORTED;
        ch.targets[ch.r++] = i;
      }
    rows_of[si].assign(static_cast<size_t>(std::max(ch.r, 1)) * k, 0);
    int valid[kMaxShards];
    if (ch.r > 0 && !rs_decode_rows(k, c->m, present.data(), ch.targets, ch.r, rows_of[si].data(), valid)) return MEC_ERR_TOO_FEW_SHARDS;
    int e;
    if ((e = s.aux.ensure(static_cast<size_t>(k * stride)))) return e;
    if ((e = s.out.ensure(static_cast<size_t>(ch.nb * std::max(ch.r, 1) * pitch)))) return e;
    if ((e = s.dig.ensure(static_cast<size_t>(ch.nb * (k + ch.r) * 32)))) return e;
    if ((e = s.flags.ensure(static_cast<size_t>(ch.nb * k)))) return e;
    if ((e = s.hflags.ensure(static_cast<size_t>(ch.nb * k)))) return e;
    FrameGeom sub = g;  // geometry of the chunk: its last block is short only if it is the range's last block
    sub.nblocks = ch.nb;
    sub.last_len = (b0 + ch.nb == g.nblocks) ? g.last_len : g.S;
    // (measured: one contiguous copy per survivor file + a device-side move to the aligned pitch is slower than these 2-D copies,
    //  36.4 vs 38.3 GiB/s on the 2 GiB GetObject, although a bare copy-engine probe favours contiguous transfers — tools/copy_probe.cu)
    const bool batched = ch.nb <= kBatchedCopyBlocks;
    CopyList list;
    if (batched) {
      for (int t = 0; t < k; t++)
        for (int64_t j = 0; j < ch.nb; j++)
          list.add(arena_ptr(s, stride, t) + j * P, frames[ch.chosen[t]] + (b0 + j) * fstride, 32 + (j == ch.nb - 1 ? sub.last_len : g.S));
      if ((e = list.flush(s.st))) return e;
    } else {
      for (int t = 0; t < k; t++)
        if ((e = stage_frames(frames[ch.chosen[t]] + b0 * fstride, arena_ptr(s, stride, t), sub, s.st))) return e;
    }
    c->st_h2d += sub.file_bytes() * k;
    MEC_CUDA_OK(cudaMemsetAsync(s.flags.p, 0, static_cast<size_t>(ch.nb * k), s.st));
    if ((e = launch_reconstruct(c, g, ch, s, stride, rows_of[si].data(), pitch, hash_outputs))) return e;
    MEC_CUDA_OK(cudaMemcpyAsync(s.hflags.p, s.flags.p, static_cast<size_t>(ch.nb * k), cudaMemcpyDeviceToHost, s.st));
    if (sk.dst) { if ((e = emit_object(c, g, sk, ch, s, stride, pitch, batched ? &list : nullptr))) return e; }
    else if (sk.out) { if ((e = emit_frames(c, g, sk, ch, s, pitch, batched ? &list : nullptr))) return e; }
    return MEC_OK;
  };

  while (done < g.nblocks) {
    while (nbusy < kSlots && next < g.nblocks) {
      if ((rc = submit(head, next))) return rc;
      next += inflight[head].nb;
      busy[head] = true;
      head = (head + 1) % kSlots;
      nbusy++;
    }
    // retire the oldest chunk
    const int si = tail;
    Slot& s = c->slots[si];
    RChunk& ch = inflight[si];
    MEC_CUDA_OK(cudaStreamSynchronize(s.st));
    busy[si] = false;
    tail = (tail + 1) % kSlots;
    nbusy--;
    const uint8_t* fl = static_cast<const uint8_t*>(s.hflags.p);
    int64_t bad = ch.nb;
    for (int64_t b = 0; b < ch.nb && bad == ch.nb; b++)
      for (int t = 0; t < k; t++)
        if (fl[b * k + t]) { bad = b; break; }
    const int64_t good = bad;  // blocks [b0, b0 + good) are accepted
    if (good > 0 && sk.out) {  // wanted shards that were read are passed through unchanged
      std::vector<uint8_t> present(n, 0);
      for (int t = 0; t < k; t++) present[ch.chosen[t]] = 1;
      for (int i = 0; i < n; i++) {
        if (!want[i] || !present[i] || !sk.out[i] || sk.out[i] == frames[i]) continue;
        const int64_t bytes = (ch.b0 + good == g.nblocks) ? (g.file_bytes() - ch.b0 * fstride) : good * fstride;
        memcpy(sk.out[i] + ch.b0 * fstride, frames[i] + ch.b0 * fstride, static_cast<size_t>(bytes));
      }
    }
    c->st_blocks_read += good;
    c->st_shards_rebuilt += good * ch.r;
    done = ch.b0 + good;
    if (bad < ch.nb) {  // drop every chosen reader that failed at block `bad`, resume there with the next readers
      for (int t = 0; t < k; t++)
        if (fl[bad * k + t]) {
          if (alive[ch.chosen[t]]) c->st_corrupt++;
          alive[ch.chosen[t]] = 0;
          if (corrupt) corrupt[ch.chosen[t]] = 1;
        }
      for (int q = 0; q < kSlots; q++)  // whatever was submitted behind it ran with the old reader set: discard
        if (busy[q]) { MEC_CUDA_OK(cudaStreamSynchronize(c->slots[q].st)); busy[q] = false; }
      head = tail = 0; nbusy = 0;
      next = done;
      epoch++;
    }
  }
  return MEC_OK;
}

extern "C" int mec_reconstruct_device(mec_codec* c, const uint8_t* const* d_frames, int64_t frame_pitch, int64_t nblocks,
                                      const uint8_t* want, int flags, uint8_t* d_out, int64_t out_pitch,
                                      uint8_t* d_digests, uint8_t* d_corrupt, void* stream) {
  if (!c || !d_frames || !want || nblocks < 0) return MEC_ERR_INVALID_ARGUMENT;
  const int data_only = flags & MEC_RECONSTRUCT_DATA_ONLY;
  int rc = require_streaming(c);
  if (rc) return rc;
  if (nblocks == 0) return MEC_OK;
  const int k = c->k, n = c->n;
  if (n > kMaxShards || k > kMaxK) return MEC_ERR_UNSUPPORTED;
  const int64_t S = c->S();
  if ((frame_pitch & 15) || frame_pitch < 32 + S) return MEC_ERR_INVALID_ARGUMENT;
  if (d_corrupt && !d_digests) return MEC_ERR_INVALID_ARGUMENT;  // the digest check runs in the hash threads: no digests, no bitrot verdict
  int chosen[kMaxShards], nch = 0;
  std::vector<uint8_t> present(n, 0);
  for (int i = 0; i < n && nch < k; i++)
    if (d_frames[i]) { chosen[nch++] = i; present[i] = 1; }
  if (nch < k) return MEC_ERR_READ_QUORUM;
  int targets[kMaxShards], r = 0;
  for (int i = 0; i < n; i++)
    if (want[i] && !present[i] && !(data_only && i >= k)) targets[r++] = i;
  if (r > kMaxR) return MEC_ERR_UNSUPPORTED;
  std::vector<uint8_t> rows(static_cast<size_t>(std::max(r, 1)) * k);
  int valid[kMaxShards];
  if (r > 0 && !rs_decode_rows(k, c->m, present.data(), targets, r, rows.data(), valid)) return MEC_ERR_TOO_FEW_SHARDS;
  std::lock_guard<std::mutex> lk(c->mu);
  MEC_CUDA_OK(cudaSetDevice(c->device));
  FusedDesc d;
  d.k = k; d.r = r; d.coef = rows.data(); d.static_encode = false; d.contiguous = false;
  d.key = kMagicKey; d.nblocks = nblocks; d.S = static_cast<int32_t>(S);
  d.in_block_stride = frame_pitch; d.expect_block_stride = frame_pitch;
  for (int t = 0; t < k; t++) {
    const uint8_t* base = d_frames[chosen[t]];
    d.map_base[t] = base;
    d.map_len[t] = nblocks * frame_pitch;
    d.expect_ptr[t] = d_corrupt ? base : nullptr;
    d.in_ptr[t] = base + 32;
  }
  d.out = d_out; d.out_pitch = out_pitch; d.digests = d_digests; d.corrupt = d_corrupt;
  d.hash_outputs = !(flags & MEC_RECONSTRUCT_NO_OUTPUT_DIGESTS);
  return c->eng->launch_fused(d, c->opt, static_cast<cudaStream_t>(stream));
}

extern "C" int mec_reconstruct_frames(mec_codec* c, const uint8_t* const* frames, int64_t nblocks, int64_t last_shard_len,
                                      const uint8_t* want, int data_only, uint8_t* const* out, uint8_t* corrupt) {
  if (!c || !frames || !want || !out || nblocks < 0) return MEC_ERR_INVALID_ARGUMENT;
  int rc = require_streaming(c);
  if (rc) return rc;
  if (c->n > kMaxShards) return MEC_ERR_UNSUPPORTED;
  std::lock_guard<std::mutex> lk(c->mu);
  if (corrupt) memset(corrupt, 0, static_cast<size_t>(c->n));
  if (nblocks == 0) return MEC_OK;
  FrameGeom g{nblocks, c->S(), last_shard_len > 0 ? last_shard_len : c->S()};
  if (g.last_len > g.S) return MEC_ERR_INVALID_ARGUMENT;
  std::vector<uint8_t> alive(c->n);
  for (int i = 0; i < c->n; i++) alive[i] = frames[i] != nullptr;
  RangeSink sk;
  sk.out = out;
  return reconstruct_range(c, frames, g, want, data_only, sk, corrupt, alive.data());
}

extern "C" int64_t mec_decode(mec_codec* c, const uint8_t* const* files, int64_t offset, int64_t length,
                              int64_t total, uint8_t* dst, int* heal_hint) {
  return mec_decode_prefer(c, files, nullptr, offset, length, total, dst, heal_hint);
}

extern "C" int64_t mec_decode_prefer(mec_codec* c, const uint8_t* const* files, const uint8_t* prefer, int64_t offset,
                                     int64_t length, int64_t total, uint8_t* dst, int* heal_hint) {
  NvtxRange nvtx("mec_decode");
  if (heal_hint) *heal_hint = 0;
  if (!c || !files) return MEC_ERR_INVALID_ARGUMENT;
  if (offset < 0 || length < 0) return MEC_ERR_INVALID_ARGUMENT;      // cmd/erasure-decode.go:240-242
  if (offset + length > total) return MEC_ERR_INVALID_ARGUMENT;       // :243-245
  if (length == 0) return 0;                                          // :247-249
  if (!dst) return MEC_ERR_INVALID_ARGUMENT;
  int rc = require_streaming(c);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  const int k = c->k, n = c->n;
  const int64_t bs = c->block_size, S = c->S(), sfs = mec_shard_file_size(c, total);
  const int64_t start_block = offset / bs, end_block = (offset + length) / bs;
  // blocks that are actually read: [start_block, last_block]
  int64_t last_block = end_block;
  if ((offset + length) % bs == 0) last_block = end_block - 1;  // blockLength == 0 => loop breaks (:279-281)
  const int64_t nblocks_total = ceil_frac(sfs, S);
  if (last_block >= nblocks_total) last_block = nblocks_total - 1;
  FrameGeom g;
  g.nblocks = last_block - start_block + 1;
  g.S = S;
  g.last_len = std::min(S, sfs - last_block * S);
  // writeDataBlocks (cmd/erasure-utils.go:42): a block whose k shards hold fewer bytes than asked of it is ErrShortData
  // (every block but the part's last holds k*S >= block_size bytes, so only that one can fall short)
  if (static_cast<int64_t>(k) * g.last_len < std::min(offset + length, total) - last_block * bs) return MEC_ERR_SHORT_DATA;
  const int64_t fstride = 32 + S, foff = start_block * fstride;
  std::vector<const uint8_t*> in(n);
  std::vector<uint8_t> alive(n), want(n, 0), corrupt(n, 0);
  for (int i = 0; i < n; i++) { in[i] = files[i] ? files[i] + foff : nullptr; alive[i] = files[i] != nullptr; }
  for (int i = 0; i < k; i++) want[i] = 1;
  RangeSink sk;
  sk.dst = dst; sk.first_block = start_block; sk.lo = offset; sk.hi = offset + length; sk.bs = bs; sk.total = total;
  // GetObject only consumes the data bytes of the rebuilt shards: their digests are not computed
  rc = reconstruct_range(c, in.data(), g, want.data(), 1, sk, corrupt.data(), alive.data(), prefer, false);
  if (rc) return rc;
  if (heal_hint)
    for (int i = 0; i < n; i++)
      if (corrupt[i]) *heal_hint = MEC_ERR_FILE_CORRUPT;
  return length;
}

// Erasure.Heal (cmd/erasure-decode.go:317).  Like the reference it reports bitrot met on the way: when a source reader failed
// its digest the stale shards are still rebuilt from the others and written, and the call returns MEC_ERR_FILE_CORRUPT
// (Heal's `derr`, :338-341,366 — healObject aborts the part on it, cmd/erasure-healing.go:603-608); corrupt[i] names the readers.
extern "C" int mec_heal_prefer(mec_codec* c, const uint8_t* const* files, const uint8_t* prefer, int64_t total,
                               uint8_t* const* out_files, uint8_t* corrupt) {
  NvtxRange nvtx("mec_heal");
  if (!c || !files || !out_files) return MEC_ERR_INVALID_ARGUMENT;
  if (corrupt) memset(corrupt, 0, static_cast<size_t>(c->n));
  int rc = require_streaming(c);
  if (rc) return rc;
  if (total <= 0) return MEC_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  const int n = c->n;
  const int64_t S = c->S(), sfs = mec_shard_file_size(c, total);
  FrameGeom g;
  g.nblocks = ceil_frac(total, c->block_size);
  g.S = S;
  g.last_len = sfs - (g.nblocks - 1) * S;
  std::vector<uint8_t> alive(n), want(n, 0), bad(n, 0);
  for (int i = 0; i < n; i++) { alive[i] = files[i] != nullptr; want[i] = out_files[i] != nullptr; }
  RangeSink sk;
  sk.out = out_files;
  rc = reconstruct_range(c, files, g, want.data(), 0, sk, bad.data(), alive.data(), prefer);
  if (rc) return rc;
  bool any = false;
  for (int i = 0; i < n; i++) { any |= bad[i] != 0; if (corrupt) corrupt[i] = bad[i]; }
  return any ? MEC_ERR_FILE_CORRUPT : MEC_OK;
}
extern "C" int mec_heal(mec_codec* c, const uint8_t* const* files, int64_t total, uint8_t* const* out_files) {
  return mec_heal_prefer(c, files, nullptr, total, out_files, nullptr);
}

// Batched heal (BASELINE config 4): objects are independent, so a pool of codec handles — each with
// its own streams and staging buffers — is driven by one host thread per handle; the H2D staging of one object then
// overlaps the kernel and the D2H of the others (healObject's callers, cmd/global-heal.go:152, do the same with goroutines).
// rcs[o] = MEC_ERR_FILE_CORRUPT means "healed, but a source reader of object o failed its digest" (see mec_heal_prefer);
// the return value is the first result that is neither MEC_OK nor that.
extern "C" int mec_heal_batch(mec_codec* const* pool, int npool, int64_t nobjects, const uint8_t* const* const* files,
                              const int64_t* totals, uint8_t* const* const* out_files, int* rcs) {
  if (!pool || npool <= 0 || nobjects < 0 || (nobjects > 0 && (!files || !totals || !out_files))) return MEC_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < npool; i++)
    if (!pool[i]) return MEC_ERR_INVALID_ARGUMENT;
  std::atomic<int64_t> next{0};
  std::atomic<int> first_err{MEC_OK};
  auto work = [&](int w) {
    for (;;) {
      const int64_t o = next.fetch_add(1);
      if (o >= nobjects) return;
      const int rc = mec_heal(pool[w], files[o], totals[o], out_files[o]);
      if (rcs) rcs[o] = rc;
      if (rc != MEC_OK && rc != MEC_ERR_FILE_CORRUPT) {
        int expected = MEC_OK;
        first_err.compare_exchange_strong(expected, rc);
      }
    }
  };
  const int nthreads = static_cast<int>(std::min<int64_t>(npool, nobjects));
  std::vector<std::thread> th;
  for (int w = 1; w < nthreads; w++) th.emplace_back(work, w);
  if (nthreads > 0) work(0);
  for (auto& t : th) t.join();
  return first_err.load();
}

// bitrotVerify (cmd/bitrot.go:164-215), streaming algorithm: every frame's digest is recomputed and compared.  The scan of one
// or many shard files (the scanner's deep-scan visits every part of every drive of an object: cmd/xl-storage.go:2550-2600) is cut
// into chunks of frames that run through the codec's slots: staging of chunk c+1 overlaps the hash-only launch (k = 1, r = 0) of
// chunk c; a chunk's verdicts come back when its slot is reused.  result[f] = MEC_OK / MEC_ERR_FILE_CORRUPT per file.
static int verify_files_locked(mec_codec* c, int64_t nfiles, const uint8_t* const* files, const int64_t* file_lens, const int64_t* part_lens,
                               int* results) {
  const int64_t S = c->S(), fstride = 32 + S, P = round_up(32 + S, 16);
  MEC_CUDA_OK(cudaSetDevice(c->device));
  struct Drain { mec_codec* c; ~Drain() { for (auto& s : c->slots) if (s.st) cudaStreamSynchronize(s.st); } } drain{c};
  struct Piece { int64_t file, nb; };
  std::vector<Piece> inflight[kSlots];
  const int64_t chunk = std::max<int64_t>(1, (32ll << 20) / fstride);  // frames per chunk (~32 MiB)
  int rc;
  auto retire = [&](int si) {
    const uint8_t* fl = static_cast<const uint8_t*>(c->slots[si].hflags.p);
    int64_t o = 0;
    for (const Piece& pc : inflight[si]) {
      for (int64_t b = 0; b < pc.nb; b++)
        if (fl[o + b]) results[pc.file] = MEC_ERR_FILE_CORRUPT;
      o += pc.nb;
    }
    inflight[si].clear();
  };
  int si = 0;
  int64_t f = 0, b0 = 0;  // next file / next frame of it
  while (f < nfiles) {
    Slot& s = c->slots[si];
    MEC_CUDA_OK(cudaStreamSynchronize(s.st));
    retire(si);
    if ((rc = s.src.ensure(static_cast<size_t>(chunk * P + 512)))) return rc;
    if ((rc = s.dig.ensure(static_cast<size_t>(chunk * 32)))) return rc;
    if ((rc = s.flags.ensure(static_cast<size_t>(chunk)))) return rc;
    if ((rc = s.hflags.ensure(static_cast<size_t>(chunk)))) return rc;
    MEC_CUDA_OK(cudaMemsetAsync(s.flags.p, 0, static_cast<size_t>(chunk), s.st));
    int64_t used = 0;
    // fill the chunk with frames of consecutive files.  When the chunk is a launch the latency kernel takes (it is, unless options
    // say otherwise: ~380 frames of a 1 MiB-block part file), every frame — full or short — hashes in ONE launch with a per-frame
    // length table, and every frame is staged by ONE batched copy call: a scan over thousands of small part files costs two driver
    // calls per 32 MiB, not two per file.  Otherwise: full frames in one launch, short last frames one launch each.
    const bool one_launch = c->eng->small_ok(c->opt, chunk);
    CopyList list;
    int32_t* lens = nullptr;
    if (one_launch) {
      if ((rc = s.hreq.ensure(static_cast<size_t>(chunk) * sizeof(int32_t)))) return rc;
      if ((rc = s.dreq.ensure(static_cast<size_t>(chunk) * sizeof(int32_t)))) return rc;
      lens = static_cast<int32_t*>(s.hreq.p);
    }
    struct Tail { int64_t slot, len; };
    std::vector<Tail> tails;
    while (f < nfiles && used < chunk) {
      if (results[f] != MEC_OK || part_lens[f] == 0) { f++; b0 = 0; continue; }
      const int64_t nblocks = ceil_frac(part_lens[f], S), last_len = part_lens[f] - (nblocks - 1) * S;
      const int64_t take = std::min(chunk - used, nblocks - b0);
      FrameGeom sub{take, S, (b0 + take == nblocks) ? last_len : S};
      if (one_launch) {  // every frame's length goes into the per-frame table; short runs of frames join the chunk's batched copy call
        for (int64_t j = 0; j < take; j++) lens[used + j] = static_cast<int32_t>((j == take - 1) ? sub.last_len : S);
      }
      if (one_launch && take < kBatchedCopyBlocks) {
        for (int64_t j = 0; j < take; j++)
          list.add(static_cast<uint8_t*>(s.src.p) + (used + j) * P, files[f] + (b0 + j) * fstride, 32 + lens[used + j]);
      } else if ((rc = stage_frames(files[f] + b0 * fstride, static_cast<uint8_t*>(s.src.p) + used * P, sub, s.st))) return rc;
      c->st_h2d += sub.file_bytes();
      if (sub.last_len != S) tails.push_back(Tail{used + take - 1, sub.last_len});
      inflight[si].push_back(Piece{f, take});
      used += take;
      b0 += take;
      if (b0 == nblocks) { f++; b0 = 0; }
    }
    if (used == 0) break;
    const int32_t* block_len = nullptr;
    auto launch = [&](int64_t slot0, int64_t cnt, int64_t len) -> int {
      FusedDesc d;
      d.k = 1; d.r = 0; d.contiguous = false; d.key = kMagicKey;
      d.block_len = block_len;
      d.in_block_stride = P; d.expect_block_stride = P;
      const uint8_t* base = static_cast<const uint8_t*>(s.src.p) + slot0 * P;
      d.nblocks = cnt; d.S = static_cast<int32_t>(len);
      d.map_base[0] = base; d.map_len[0] = (chunk - slot0) * P + 512;
      d.expect_ptr[0] = base; d.in_ptr[0] = base + 32;
      d.digests = static_cast<uint8_t*>(s.dig.p) + slot0 * 32;
      d.corrupt = static_cast<uint8_t*>(s.flags.p) + slot0;
      return c->eng->launch_fused(d, c->opt, s.st);
    };
    if (one_launch) {
      if ((rc = list.flush(s.st))) return rc;
      MEC_CUDA_OK(cudaMemcpyAsync(s.dreq.p, lens, static_cast<size_t>(used) * sizeof(int32_t), cudaMemcpyHostToDevice, s.st));
      block_len = static_cast<const int32_t*>(s.dreq.p);
      if ((rc = launch(0, used, S))) return rc;
      tails.clear();
    }
    // runs of full frames between the short ones (a short frame's digest covers fewer bytes: its own launch)
    int64_t at = one_launch ? used : 0;
    for (size_t q = 0; q <= tails.size(); q++) {
      const int64_t end = q < tails.size() ? tails[q].slot : used;
      if (end > at && (rc = launch(at, end - at, S))) return rc;
      if (q < tails.size()) {
        if ((rc = launch(tails[q].slot, 1, tails[q].len))) return rc;
        at = tails[q].slot + 1;
      }
    }
    MEC_CUDA_OK(cudaMemcpyAsync(s.hflags.p, s.flags.p, static_cast<size_t>(used), cudaMemcpyDeviceToHost, s.st));
    c->st_blocks_read += used;
    si = (si + 1) % kSlots;
  }
  for (int q = 0; q < kSlots; q++, si = (si + 1) % kSlots) {
    MEC_CUDA_OK(cudaStreamSynchronize(c->slots[si].st));
    retire(si);
  }
  return MEC_OK;
}

extern "C" int mec_bitrot_verify_batch(mec_codec* c, int64_t nfiles, const uint8_t* const* files, const int64_t* file_lens,
                                       const int64_t* part_lens, int* results) {
  if (!c || nfiles < 0 || (nfiles > 0 && (!files || !file_lens || !part_lens || !results))) return MEC_ERR_INVALID_ARGUMENT;
  const int64_t S = c->S();
  bool any = false;
  for (int64_t f = 0; f < nfiles; f++) {
    results[f] = MEC_OK;
    if (!files[f] && part_lens[f] > 0) { results[f] = MEC_ERR_INVALID_ARGUMENT; continue; }
    if (file_lens[f] != mec_bitrot_shard_file_size(part_lens[f], S, c->algo)) results[f] = MEC_ERR_FILE_CORRUPT;  // cmd/bitrot.go:183
    else if (part_lens[f] > 0) any = true;
  }
  if (!any) return MEC_OK;
  int rc = require_streaming(c);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  return verify_files_locked(c, nfiles, files, file_lens, part_lens, results);
}

extern "C" int mec_bitrot_verify(mec_codec* c, const uint8_t* file, int64_t file_len, int64_t part_len) {
  if (!c || !file) return MEC_ERR_INVALID_ARGUMENT;
  int result = MEC_OK;
  const int rc = mec_bitrot_verify_batch(c, 1, &file, &file_len, &part_len, &result);
  return rc ? rc : result;
}

// ------------------------------------------------------------------------------------------------
// legacy whole-file bitrot algorithms
extern "C" int mec_digest_size(int algo) {
  switch (algo) {
    case MEC_SHA256: case MEC_HIGHWAYHASH256: case MEC_HIGHWAYHASH256S: return 32;
    case MEC_BLAKE2B512: return 64;
  }
  return -1;
}

static int whole_hash_device(mec_codec* c, int algo, const uint8_t* d_data, int64_t pitch, int64_t len, int64_t count,
                             uint8_t* d_out, cudaStream_t st) {
  WholeHashParams hp;
  hp.data = d_data; hp.pitch = pitch; hp.len = len; hp.nstreams = static_cast<int>(count);
  hp.algo = algo == MEC_HIGHWAYHASH256S ? MEC_HIGHWAYHASH256 : algo;
  hp.out = d_out;
  memcpy(hp.key, kMagicKey, 32);
  const int threads = 64;
  const int64_t nthreads = hp.algo == MEC_HIGHWAYHASH256 ? 2 * count : count;
  whole_hash_kernel<<<static_cast<unsigned>((nthreads + threads - 1) / threads), threads, 0, st>>>(hp);
  MEC_CUDA_OK(cudaGetLastError());
  (void)c;
  return MEC_OK;
}

extern "C" int mec_whole_hash(mec_codec* c, int algo, const uint8_t* msgs, int64_t msg_len, int64_t count, uint8_t* digests) {
  if (!c || msg_len < 0 || count < 0 || mec_digest_size(algo) < 0) return MEC_ERR_INVALID_ARGUMENT;
  if (count == 0) return MEC_OK;
  if (count >= (1ll << 30)) return MEC_ERR_UNSUPPORTED;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc;
  if ((rc = ensure_engine(c))) return rc;
  MEC_CUDA_OK(cudaSetDevice(c->device));
  Slot& s = c->slots[0];
  const int64_t pitch = round_up(msg_len, 16) + 128;  // the kernel reads whole 16-byte vectors
  if ((rc = s.src.ensure(static_cast<size_t>(count * pitch + 256)))) return rc;
  if ((rc = s.dig.ensure(static_cast<size_t>(count * 64)))) return rc;
  if (msg_len > 0)
    MEC_CUDA_OK(cudaMemcpy2DAsync(s.src.p, static_cast<size_t>(pitch), msgs, static_cast<size_t>(msg_len), static_cast<size_t>(msg_len),
                                  static_cast<size_t>(count), cudaMemcpyHostToDevice, s.st));
  if ((rc = whole_hash_device(c, algo, static_cast<const uint8_t*>(s.src.p), pitch, msg_len, count, static_cast<uint8_t*>(s.dig.p), s.st))) return rc;
  const int ds = mec_digest_size(algo);
  MEC_CUDA_OK(cudaMemcpy2DAsync(digests, static_cast<size_t>(ds), s.dig.p, 64, static_cast<size_t>(ds), static_cast<size_t>(count),
                                cudaMemcpyDeviceToHost, s.st));
  MEC_CUDA_OK(cudaStreamSynchronize(s.st));
  c->eng->count_launch();
  return MEC_OK;
}

extern "C" int mec_whole_hash_device(mec_codec* c, int algo, const uint8_t* d_msgs, int64_t pitch, int64_t msg_len, int64_t count,
                                     uint8_t* d_digests, void* stream) {
  if (!c || msg_len < 0 || count < 0 || mec_digest_size(algo) < 0 || (pitch & 15) || pitch < msg_len) return MEC_ERR_INVALID_ARGUMENT;
  if (count == 0) return MEC_OK;
  if (count >= (1ll << 30)) return MEC_ERR_UNSUPPORTED;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc;
  if ((rc = ensure_engine(c))) return rc;
  MEC_CUDA_OK(cudaSetDevice(c->device));
  if ((rc = whole_hash_device(c, algo, d_msgs, pitch, msg_len, count, d_digests, static_cast<cudaStream_t>(stream)))) return rc;
  c->eng->count_launch();
  return MEC_OK;
}

extern "C" int mec_bitrot_verify_whole(mec_codec* c, int algo, const uint8_t* file, int64_t file_len, const uint8_t* want) {
  if (!c || !want || file_len < 0) return MEC_ERR_INVALID_ARGUMENT;
  uint8_t got[64];
  int rc = mec_whole_hash(c, algo, file, file_len, 1, got);
  if (rc) return rc;
  return memcmp(got, want, static_cast<size_t>(mec_digest_size(algo))) ? MEC_ERR_FILE_CORRUPT : MEC_OK;  // cmd/bitrot.go:171-173
}

extern "C" int64_t mec_encode_whole(mec_codec* c, const uint8_t* src, int64_t len, uint8_t* const* files, uint8_t* sums,
                                    int write_quorum) {
  if (!c || len < 0 || !files) return MEC_ERR_INVALID_ARGUMENT;
  if (c->algo == MEC_HIGHWAYHASH256S) return MEC_ERR_INVALID_ARGUMENT;
  int online = 0;
  for (int i = 0; i < c->n; i++) online += files[i] != nullptr;
  if (online < write_quorum) return MEC_ERR_WRITE_QUORUM;
  if (c->S() >= (1ll << 31)) return MEC_ERR_UNSUPPORTED;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc;
  if ((rc = ensure_engine(c))) return rc;
  MEC_CUDA_OK(cudaSetDevice(c->device));
  const int k = c->k, m = c->m, n = c->n;
  const int64_t bs = c->block_size, S = c->S(), pitch = round_up(S, 16);
  const int64_t nall = ceil_frac(len, bs), flen = mec_shard_file_size(c, len), fpitch = round_up(flen, 16) + 128;
  Slot& s = c->slots[0];
  if ((rc = s.src.ensure(static_cast<size_t>(round_up(len, 16) + 256)))) return rc;
  if ((rc = s.out.ensure(static_cast<size_t>(std::max<int64_t>(nall, 1) * std::max(m, 1) * pitch)))) return rc;
  if ((rc = s.aux.ensure(static_cast<size_t>(n * fpitch + 256)))) return rc;
  if ((rc = s.dig.ensure(static_cast<size_t>(n * 64)))) return rc;
  if (len > 0) {
    MEC_CUDA_OK(cudaMemcpyAsync(s.src.p, src, static_cast<size_t>(len), cudaMemcpyHostToDevice, s.st));
    if (m > 0) {  // parity only: the hash is not per block for these algorithms
      const int64_t nfull = len / bs, tail = len % bs;
      FusedDesc d;
      d.k = k; d.r = m; d.coef = c->matrix.data() + static_cast<size_t>(k) * k; d.static_encode = true; d.contiguous = true;
      d.key = kMagicKey; d.out_pitch = pitch; d.digests = nullptr;
      if (nfull > 0) {
        d.nblocks = nfull; d.S = static_cast<int32_t>(S); d.in_base = static_cast<const uint8_t*>(s.src.p);
        d.in_block_stride = bs; d.in_block_len = bs; d.out = static_cast<uint8_t*>(s.out.p);
        if ((rc = c->eng->launch_fused(d, c->opt, s.st))) return rc;
      }
      if (tail > 0) {
        d.nblocks = 1; d.S = static_cast<int32_t>(ceil_frac(tail, k)); d.in_base = static_cast<const uint8_t*>(s.src.p) + nfull * bs;
        d.in_block_stride = round_up(tail, 16); d.in_block_len = tail; d.out = static_cast<uint8_t*>(s.out.p) + nfull * m * pitch;
        if ((rc = c->eng->launch_fused(d, c->opt, s.st))) return rc;
      }
    }
    GatherParams gp;
    gp.src = static_cast<const uint8_t*>(s.src.p); gp.parity = static_cast<const uint8_t*>(s.out.p); gp.parity_pitch = pitch;
    gp.block_size = bs; gp.len = len; gp.k = k; gp.m = m; gp.S = S; gp.files = static_cast<uint8_t*>(s.aux.p);
    gp.file_pitch = fpitch; gp.file_len = flen;
    const unsigned gx = static_cast<unsigned>(std::min<int64_t>((flen + 255) / 256, 4096));
    gather_shard_files_kernel<<<dim3(gx, static_cast<unsigned>(n)), 256, 0, s.st>>>(gp);
    MEC_CUDA_OK(cudaGetLastError());
    c->eng->count_launch();
  }
  if ((rc = whole_hash_device(c, c->algo, static_cast<const uint8_t*>(s.aux.p), fpitch, flen, n, static_cast<uint8_t*>(s.dig.p), s.st))) return rc;
  c->eng->count_launch();
  for (int i = 0; i < n; i++)
    if (files[i] && flen > 0)
      MEC_CUDA_OK(cudaMemcpyAsync(files[i], static_cast<uint8_t*>(s.aux.p) + i * fpitch, static_cast<size_t>(flen), cudaMemcpyDeviceToHost, s.st));
  if (sums) MEC_CUDA_OK(cudaMemcpyAsync(sums, s.dig.p, static_cast<size_t>(n * 64), cudaMemcpyDeviceToHost, s.st));
  MEC_CUDA_OK(cudaStreamSynchronize(s.st));
  return len;
}

// ------------------------------------------------------------------------------------------------
// Decode / Heal over whole-file bitrot readers (cmd/bitrot-whole.go:66-81): a wholeBitrotReader reads its shard file on the first
// ReadAt and checks the digest of the WHOLE file (xlStorage.ReadFile with a verifier, cmd/xl-storage.go:1931-1950) — a mismatch
// is errFileCorrupt, the reader is dropped and parallelReader moves on to the next drive.  Here: the first k readers in index
// order are staged and hashed on the GPU (whole_hash_kernel), failing ones are replaced by the next alive reader until k
// verified shard files sit on the device; the rebuild then runs on those raw shards (no per-block digests exist in this format).
static int require_whole(mec_codec* c) {
  if (c->algo == MEC_HIGHWAYHASH256S) {
    set_last_error("whole-file entry points need a codec created with SHA256, BLAKE2b512 or HighwayHash256");
    return MEC_ERR_INVALID_ARGUMENT;
  }
  return ensure_engine(c);
}

// stage + verify: on return chosen[0..k) (ascending) name k verified files, file chosen[t] at s.aux + t*fpitch
static int stage_verified_whole(mec_codec* c, const uint8_t* const* files, const uint8_t* sums, int64_t sfs, uint8_t* alive,
                                uint8_t* corrupt, int* chosen, int64_t fpitch, Slot& s) {
  const int k = c->k, n = c->n, ds = mec_digest_size(c->algo);
  int rc;
  if ((rc = s.aux.ensure(static_cast<size_t>(k * fpitch + 256)))) return rc;
  if ((rc = s.dig.ensure(static_cast<size_t>(k * 64)))) return rc;
  if ((rc = s.hdig.ensure(static_cast<size_t>(k * 64)))) return rc;
  int slot_of[kMaxShards];  // which arena position holds file i (-1 = not staged)
  for (int i = 0; i < n; i++) slot_of[i] = -1;
  bool verified[kMaxShards] = {};
  for (;;) {
    int want[kMaxShards], nw = 0;
    for (int i = 0; i < n && nw < k; i++)
      if (alive[i]) want[nw++] = i;
    if (nw < k) return MEC_ERR_READ_QUORUM;
    // arena positions: keep verified files where they are, give the newcomers the free positions
    bool used[kMaxK] = {};
    for (int q = 0; q < k; q++)
      if (slot_of[want[q]] >= 0) used[slot_of[want[q]]] = true;
    int fresh[kMaxK], nf = 0;
    for (int q = 0; q < k; q++) {
      const int i = want[q];
      if (slot_of[i] >= 0) continue;
      int pos = 0;
      while (used[pos]) pos++;
      used[pos] = true;
      slot_of[i] = pos;
      fresh[nf++] = i;
      if (sfs > 0)
        MEC_CUDA_OK(cudaMemcpyAsync(static_cast<uint8_t*>(s.aux.p) + pos * fpitch, files[i], static_cast<size_t>(sfs), cudaMemcpyHostToDevice, s.st));
      c->st_h2d += sfs;
    }
    if (nf == 0) break;
    for (int q = 0; q < nf; q++) {  // one launch per newcomer keeps the arena positions arbitrary; these files are few and the hash is serial anyway
      const int pos = slot_of[fresh[q]];
      if ((rc = whole_hash_device(c, c->algo, static_cast<const uint8_t*>(s.aux.p) + pos * fpitch, fpitch, sfs, 1,
                                  static_cast<uint8_t*>(s.dig.p) + pos * 64, s.st)))
        return rc;
      c->eng->count_launch();
    }
    MEC_CUDA_OK(cudaMemcpyAsync(s.hdig.p, s.dig.p, static_cast<size_t>(k * 64), cudaMemcpyDeviceToHost, s.st));
    MEC_CUDA_OK(cudaStreamSynchronize(s.st));
    bool dropped = false;
    for (int q = 0; q < nf; q++) {
      const int i = fresh[q], pos = slot_of[i];
      if (memcmp(static_cast<const uint8_t*>(s.hdig.p) + pos * 64, sums + static_cast<size_t>(i) * 64, static_cast<size_t>(ds)) != 0) {
        alive[i] = 0;                 // errFileCorrupt: the reader is gone for the rest of the call
        if (corrupt) corrupt[i] = 1;
        c->st_corrupt++;
        slot_of[i] = -1;
        dropped = true;
      } else {
        verified[i] = true;
      }
    }
    if (!dropped) break;
  }
  // compact into ascending order: position t must hold the t-th chosen file (decode rows are defined on ascending indices)
  int nch = 0;
  for (int i = 0; i < n && nch < k; i++)
    if (alive[i] && verified[i]) chosen[nch++] = i;
  if (nch < k) return MEC_ERR_READ_QUORUM;
  for (int t = 0; t < k; t++) chosen[t] = chosen[t] | (slot_of[chosen[t]] << 16);  // low 16 bits: shard index, high: arena position
  return MEC_OK;
}

// launch the rebuild of `targets` over blocks [first, first + nb) of raw shard files (stride S between blocks, no frames)
static int launch_whole_rebuild(mec_codec* c, Slot& s, const int* chosen, int64_t fpitch, int64_t S, int64_t last_len, int64_t first,
                                int64_t nb, int64_t nblocks_file, const uint8_t* rows, int r, int64_t pitch) {
  const int k = c->k;
  FusedDesc d;
  d.k = k; d.r = r; d.coef = rows; d.static_encode = false; d.contiguous = false; d.hash_outputs = false;
  d.key = kMagicKey; d.out_pitch = pitch; d.in_block_stride = S; d.digests = nullptr;
  const bool has_short = (first + nb == nblocks_file) && last_len != S;
  const int64_t nfull = has_short ? nb - 1 : nb;
  for (int pass = 0; pass < 2; pass++) {
    const int64_t f0 = pass == 0 ? 0 : nfull, cnt = pass == 0 ? nfull : nb - nfull;
    if (cnt <= 0) continue;
    d.nblocks = cnt;
    d.S = static_cast<int32_t>(pass == 0 ? S : last_len);
    for (int t = 0; t < k; t++) {
      const uint8_t* base = static_cast<const uint8_t*>(s.aux.p) + (chosen[t] >> 16) * fpitch;
      d.in_ptr[t] = base + (first + f0) * S;
      d.map_base[t] = base;
      d.map_len[t] = fpitch;
    }
    d.out = static_cast<uint8_t*>(s.out.p) + f0 * r * pitch;
    int rc = c->eng->launch_fused(d, c->opt, s.st);
    if (rc) return rc;
  }
  return MEC_OK;
}

extern "C" int64_t mec_decode_whole(mec_codec* c, const uint8_t* const* files, const uint8_t* sums, int64_t offset, int64_t length,
                                    int64_t total, uint8_t* dst, int* heal_hint) {
  NvtxRange nvtx("mec_decode_whole");
  if (heal_hint) *heal_hint = 0;
  if (!c || !files || !sums) return MEC_ERR_INVALID_ARGUMENT;
  if (offset < 0 || length < 0) return MEC_ERR_INVALID_ARGUMENT;      // cmd/erasure-decode.go:240-242
  if (offset + length > total) return MEC_ERR_INVALID_ARGUMENT;       // :243-245
  if (length == 0) return 0;                                          // :247-249
  if (!dst) return MEC_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc = require_whole(c);
  if (rc) return rc;
  const int k = c->k, n = c->n;
  if (k > kMaxK || n > kMaxShards || c->S() >= (1ll << 31)) return MEC_ERR_UNSUPPORTED;
  MEC_CUDA_OK(cudaSetDevice(c->device));
  const int64_t bs = c->block_size, S = c->S(), sfs = mec_shard_file_size(c, total);
  const int64_t nblocks_file = ceil_frac(sfs, S), last_len_file = sfs - (nblocks_file - 1) * S;
  const int64_t start_block = offset / bs;
  int64_t last_block = (offset + length) / bs;
  if ((offset + length) % bs == 0) last_block--;
  if (last_block >= nblocks_file) last_block = nblocks_file - 1;
  const int64_t nb = last_block - start_block + 1;
  Slot& s = c->slots[0];
  struct Drain { Slot& s; ~Drain() { if (s.st) cudaStreamSynchronize(s.st); } } drain{s};
  const int64_t fpitch = round_up(sfs, 16) + 256, pitch = round_up(S, 16);
  std::vector<uint8_t> alive(n), corrupt(n, 0);
  for (int i = 0; i < n; i++) alive[i] = files[i] != nullptr;
  int chosen[kMaxShards];
  if ((rc = stage_verified_whole(c, files, sums, sfs, alive.data(), corrupt.data(), chosen, fpitch, s))) return rc;
  std::vector<uint8_t> present(n, 0);
  for (int t = 0; t < k; t++) present[chosen[t] & 0xffff] = 1;
  int targets[kMaxShards], r = 0;
  for (int i = 0; i < k; i++)
    if (!present[i]) targets[r++] = i;
  if (r > kMaxR) return MEC_ERR_UNSUPPORTED;
  std::vector<uint8_t> rows(static_cast<size_t>(std::max(r, 1)) * k);
  int valid[kMaxShards];
  if (r > 0 && !rs_decode_rows(k, c->m, present.data(), targets, r, rows.data(), valid)) return MEC_ERR_TOO_FEW_SHARDS;
  if (r > 0) {
    if ((rc = s.out.ensure(static_cast<size_t>(nb * r * pitch)))) return rc;
    if ((rc = launch_whole_rebuild(c, s, chosen, fpitch, S, last_len_file, start_block, nb, nblocks_file, rows.data(), r, pitch))) return rc;
  }
  // writeDataBlocks by DMA: data shard i of block B sits in a verified shard file (stride S) or in a rebuilt row
  for (int i = 0; i < k; i++) {
    const uint8_t* base = nullptr;
    int64_t spitch = 0;
    for (int t = 0; t < k && !base; t++)
      if ((chosen[t] & 0xffff) == i) { base = static_cast<const uint8_t*>(s.aux.p) + (chosen[t] >> 16) * fpitch + start_block * S; spitch = S; }
    for (int q = 0; q < r && !base; q++)
      if (targets[q] == i) { base = static_cast<const uint8_t*>(s.out.p) + q * pitch; spitch = r * pitch; }
    if (!base) return MEC_ERR_UNEXPECTED;
    for (int64_t b = 0; b < nb; b++) {
      const int64_t B = start_block + b, blo = B * bs, bhi = std::min(blo + bs, total);
      const int64_t cur = (B == nblocks_file - 1) ? last_len_file : S;
      const int64_t slo = blo + static_cast<int64_t>(i) * cur, shi = std::min(slo + cur, bhi);
      const int64_t a = std::max(slo, offset), e = std::min(shi, offset + length);
      if (e > a) {
        MEC_CUDA_OK(cudaMemcpyAsync(dst + (a - offset), base + b * spitch + (a - slo), static_cast<size_t>(e - a), cudaMemcpyDeviceToHost, s.st));
        c->st_d2h += e - a;
      }
    }
  }
  MEC_CUDA_OK(cudaStreamSynchronize(s.st));
  c->st_blocks_read += nb;
  c->st_shards_rebuilt += nb * r;
  if (heal_hint)
    for (int i = 0; i < n; i++)
      if (corrupt[i]) *heal_hint = MEC_ERR_FILE_CORRUPT;
  return length;
}

extern "C" int mec_heal_whole(mec_codec* c, const uint8_t* const* files, const uint8_t* sums, int64_t total,
                              uint8_t* const* out_files, uint8_t* out_sums, uint8_t* corrupt_out) {
  NvtxRange nvtx("mec_heal_whole");
  if (!c || !files || !sums || !out_files) return MEC_ERR_INVALID_ARGUMENT;
  if (corrupt_out) memset(corrupt_out, 0, static_cast<size_t>(c->n));
  if (total <= 0) return MEC_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc = require_whole(c);
  if (rc) return rc;
  const int k = c->k, n = c->n;
  if (k > kMaxK || n > kMaxShards || c->S() >= (1ll << 31)) return MEC_ERR_UNSUPPORTED;
  MEC_CUDA_OK(cudaSetDevice(c->device));
  const int64_t S = c->S(), sfs = mec_shard_file_size(c, total);
  const int64_t nblocks = ceil_frac(sfs, S), last_len = sfs - (nblocks - 1) * S;
  Slot& s = c->slots[0];
  struct Drain { Slot& s; ~Drain() { if (s.st) cudaStreamSynchronize(s.st); } } drain{s};
  const int64_t fpitch = round_up(sfs, 16) + 256, pitch = round_up(S, 16);
  std::vector<uint8_t> alive(n), corrupt(n, 0);
  for (int i = 0; i < n; i++) alive[i] = files[i] != nullptr;
  int chosen[kMaxShards];
  if ((rc = stage_verified_whole(c, files, sums, sfs, alive.data(), corrupt.data(), chosen, fpitch, s))) return rc;
  std::vector<uint8_t> present(n, 0);
  for (int t = 0; t < k; t++) present[chosen[t] & 0xffff] = 1;
  int targets[kMaxShards], r = 0;
  for (int i = 0; i < n; i++)
    if (out_files[i] && !present[i]) targets[r++] = i;
  if (r > kMaxR) return MEC_ERR_UNSUPPORTED;
  if (r > 0) {
    std::vector<uint8_t> rows(static_cast<size_t>(r) * k);
    int valid[kMaxShards];
    if (!rs_decode_rows(k, c->m, present.data(), targets, r, rows.data(), valid)) return MEC_ERR_TOO_FEW_SHARDS;
    if ((rc = s.out.ensure(static_cast<size_t>(nblocks * r * pitch)))) return rc;
    if ((rc = launch_whole_rebuild(c, s, chosen, fpitch, S, last_len, 0, nblocks, nblocks, rows.data(), r, pitch))) return rc;
    // rebuilt rows -> contiguous shard files on the device (for the whole-file hash), then files and sums go home
    if ((rc = s.src.ensure(static_cast<size_t>(r * fpitch + 256)))) return rc;
    if ((rc = s.dig.ensure(static_cast<size_t>((k + r) * 64)))) return rc;
    if ((rc = s.hdig.ensure(static_cast<size_t>((k + r) * 64)))) return rc;
    for (int q = 0; q < r; q++) {
      uint8_t* f = static_cast<uint8_t*>(s.src.p) + q * fpitch;
      const int64_t nfull = last_len == S ? nblocks : nblocks - 1;
      if (nfull > 0)
        MEC_CUDA_OK(cudaMemcpy2DAsync(f, static_cast<size_t>(S), static_cast<uint8_t*>(s.out.p) + q * pitch, static_cast<size_t>(r * pitch),
                                      static_cast<size_t>(S), static_cast<size_t>(nfull), cudaMemcpyDeviceToDevice, s.st));
      if (nfull < nblocks)
        MEC_CUDA_OK(cudaMemcpyAsync(f + nfull * S, static_cast<uint8_t*>(s.out.p) + (nfull * r + q) * pitch, static_cast<size_t>(last_len),
                                    cudaMemcpyDeviceToDevice, s.st));
      MEC_CUDA_OK(cudaMemsetAsync(f + sfs, 0, 256, s.st));
    }
    if ((rc = whole_hash_device(c, c->algo, static_cast<const uint8_t*>(s.src.p), fpitch, sfs, r, static_cast<uint8_t*>(s.dig.p) + k * 64, s.st))) return rc;
    c->eng->count_launch();
    for (int q = 0; q < r; q++) {
      MEC_CUDA_OK(cudaMemcpyAsync(out_files[targets[q]], static_cast<uint8_t*>(s.src.p) + q * fpitch, static_cast<size_t>(sfs), cudaMemcpyDeviceToHost, s.st));
      if (out_sums)
        MEC_CUDA_OK(cudaMemcpyAsync(out_sums + static_cast<size_t>(targets[q]) * 64, static_cast<uint8_t*>(s.dig.p) + (k + q) * 64, 64, cudaMemcpyDeviceToHost, s.st));
      c->st_d2h += sfs;
    }
  }
  // wanted shards that were read and verified are passed through
  for (int i = 0; i < n; i++)
    if (out_files[i] && present[i] && out_files[i] != files[i]) {
      memcpy(out_files[i], files[i], static_cast<size_t>(sfs));
      if (out_sums) memcpy(out_sums + static_cast<size_t>(i) * 64, sums + static_cast<size_t>(i) * 64, 64);
    }
  MEC_CUDA_OK(cudaStreamSynchronize(s.st));
  c->st_blocks_read += nblocks;
  c->st_shards_rebuilt += nblocks * r;
  bool any = false;
  for (int i = 0; i < n; i++) { any |= corrupt[i] != 0; if (corrupt_out) corrupt_out[i] = corrupt[i]; }
  return any ? MEC_ERR_FILE_CORRUPT : MEC_OK;  // Heal's derr (cmd/erasure-decode.go:338-341,366)
}

// ------------------------------------------------------------------------------------------------
// shard-shaped low level calls
static int apply_rows_host(mec_codec* c, const uint8_t* rows, int r, const uint8_t* const* in, uint8_t* const* outp,
                           int64_t len, bool is_encode) {
  MEC_CUDA_OK(cudaSetDevice(c->device));
  Slot& s = c->slots[0];
  const int k = c->k;
  const int64_t ipitch = round_up(len, 256), opitch = round_up(len, 16);
  int rc;
  if ((rc = s.src.ensure(static_cast<size_t>(k * ipitch + 512)))) return rc;
  if ((rc = s.out.ensure(static_cast<size_t>(std::max(r, 1) * opitch)))) return rc;
  for (int t = 0; t < k; t++)
    MEC_CUDA_OK(cudaMemcpyAsync(static_cast<uint8_t*>(s.src.p) + t * ipitch, in[t], static_cast<size_t>(len),
                                cudaMemcpyHostToDevice, s.st));
  FusedDesc d;
  d.k = k; d.r = r; d.coef = rows; d.static_encode = is_encode; d.contiguous = false;
  d.nblocks = 1; d.S = static_cast<int32_t>(len); d.in_block_stride = 0;
  for (int t = 0; t < k; t++) {
    d.in_ptr[t] = static_cast<const uint8_t*>(s.src.p) + t * ipitch;
    d.map_base[t] = d.in_ptr[t];
    d.map_len[t] = ipitch;
  }
  d.out = static_cast<uint8_t*>(s.out.p); d.out_pitch = opitch; d.digests = nullptr; d.key = kMagicKey;
  if ((rc = c->eng->launch_fused(d, c->opt, s.st))) return rc;
  for (int q = 0; q < r; q++)
    MEC_CUDA_OK(cudaMemcpyAsync(outp[q], static_cast<uint8_t*>(s.out.p) + q * opitch, static_cast<size_t>(len),
                                cudaMemcpyDeviceToHost, s.st));
  MEC_CUDA_OK(cudaStreamSynchronize(s.st));
  return MEC_OK;
}

extern "C" int mec_rs_encode_shards(mec_codec* c, uint8_t* const* shards, int64_t len) {
  if (!c || !shards) return MEC_ERR_INVALID_ARGUMENT;
  if (len == 0) return MEC_ERR_SHARD_NO_DATA;
  if (len >= (1ll << 31) || c->k > kMaxK || c->m > kMaxR) return MEC_ERR_UNSUPPORTED;
  if (c->m == 0) return MEC_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  if (int rc = ensure_engine(c)) return rc;
  return apply_rows_host(c, c->matrix.data() + static_cast<size_t>(c->k) * c->k, c->m, shards, shards + c->k, len, true);
}

extern "C" int mec_rs_reconstruct_shards(mec_codec* c, uint8_t* const* shards, const uint8_t* present, int64_t len,
                                         int data_only) {
  if (!c || !shards || !present) return MEC_ERR_INVALID_ARGUMENT;
  const int k = c->k, n = c->n;
  int np = 0;
  for (int i = 0; i < n; i++) np += present[i] != 0;
  if (np == 0 || len == 0) return MEC_ERR_SHARD_NO_DATA;
  if (np == n) return MEC_OK;
  if (np < k) return MEC_ERR_TOO_FEW_SHARDS;
  if (len >= (1ll << 31) || k > kMaxK) return MEC_ERR_UNSUPPORTED;
  int targets[kMaxShards], r = 0;
  for (int i = 0; i < n; i++)
    if (!present[i] && !(data_only && i >= k)) targets[r++] = i;
  if (r == 0) return MEC_OK;
  if (r > kMaxR) return MEC_ERR_UNSUPPORTED;
  std::vector<uint8_t> rows(static_cast<size_t>(r) * k);
  int valid[kMaxShards];
  if (!rs_decode_rows(k, c->m, present, targets, r, rows.data(), valid)) return MEC_ERR_TOO_FEW_SHARDS;
  std::vector<const uint8_t*> in(k);
  std::vector<uint8_t*> outp(r);
  for (int t = 0; t < k; t++) in[t] = shards[valid[t]];
  for (int q = 0; q < r; q++) outp[q] = shards[targets[q]];
  std::lock_guard<std::mutex> lk(c->mu);
  if (int rc = ensure_engine(c)) return rc;
  return apply_rows_host(c, rows.data(), r, in.data(), outp.data(), len, false);
}

extern "C" int mec_hh256_batch(mec_codec* c, const uint8_t* msgs, int64_t msg_len, int64_t count, uint8_t* digests) {
  if (!c || msg_len < 0 || count < 0) return MEC_ERR_INVALID_ARGUMENT;
  if (count == 0) return MEC_OK;
  if (msg_len >= (1ll << 31)) return MEC_ERR_UNSUPPORTED;
  std::lock_guard<std::mutex> lk(c->mu);
  int rc;
  if ((rc = ensure_engine(c))) return rc;
  MEC_CUDA_OK(cudaSetDevice(c->device));
  Slot& s = c->slots[0];
  if ((rc = s.dig.ensure(static_cast<size_t>(count * 32)))) return rc;
  const int64_t mp = std::max<int64_t>(round_up(msg_len, 16), 16);
  if ((rc = s.src.ensure(static_cast<size_t>(count * mp + 512)))) return rc;
  if (msg_len > 0)
    MEC_CUDA_OK(cudaMemcpy2DAsync(s.src.p, static_cast<size_t>(mp), msgs, static_cast<size_t>(msg_len),
                                  static_cast<size_t>(msg_len), static_cast<size_t>(count), cudaMemcpyHostToDevice, s.st));
  FusedDesc d;
  d.k = 1; d.r = 0; d.contiguous = false; d.key = kMagicKey;
  d.nblocks = count; d.S = static_cast<int32_t>(msg_len);
  d.in_block_stride = mp;
  d.in_ptr[0] = static_cast<const uint8_t*>(s.src.p);
  d.map_base[0] = d.in_ptr[0];
  d.map_len[0] = count * mp + 256;
  d.digests = static_cast<uint8_t*>(s.dig.p);
  if ((rc = c->eng->launch_fused(d, c->opt, s.st))) return rc;
  MEC_CUDA_OK(cudaMemcpyAsync(digests, s.dig.p, static_cast<size_t>(count * 32), cudaMemcpyDeviceToHost, s.st));
  MEC_CUDA_OK(cudaStreamSynchronize(s.st));
  return MEC_OK;
}

// ------------------------------------------------------------------------------------------------
// self tests (cmd/erasure-coding.go:149-205, cmd/bitrot.go:224-254)
static uint64_t xxh64(const uint8_t* p, size_t n) {
  const uint64_t P1 = 11400714785074694791ull, P2 = 14029467366897019727ull, P3 = 1609587929392839161ull,
                 P4 = 9650029242287828579ull, P5 = 2870177450012600261ull;
  auto rol = [](uint64_t x, int r) { return (x << r) | (x >> (64 - r)); };
  auto rd64 = [](const uint8_t* q) { uint64_t v; memcpy(&v, q, 8); return v; };
  auto rd32 = [](const uint8_t* q) { uint32_t v; memcpy(&v, q, 4); return static_cast<uint64_t>(v); };
  auto round = [&](uint64_t acc, uint64_t in) { acc += in * P2; acc = rol(acc, 31); return acc * P1; };
  auto merge = [&](uint64_t acc, uint64_t v) { acc ^= round(0, v); return acc * P1 + P4; };
  const uint8_t* end = p + n;
  uint64_t h;
  if (n >= 32) {
    uint64_t v1 = P1 + P2, v2 = P2, v3 = 0, v4 = 0 - P1;
    do {
      v1 = round(v1, rd64(p)); v2 = round(v2, rd64(p + 8)); v3 = round(v3, rd64(p + 16)); v4 = round(v4, rd64(p + 24));
      p += 32;
    } while (p + 32 <= end);
    h = rol(v1, 1) + rol(v2, 7) + rol(v3, 12) + rol(v4, 18);
    h = merge(h, v1); h = merge(h, v2); h = merge(h, v3); h = merge(h, v4);
  } else {
    h = P5;
  }
  h += n;
  while (p + 8 <= end) { h ^= round(0, rd64(p)); h = rol(h, 27) * P1 + P4; p += 8; }
  if (p + 4 <= end) { h ^= rd32(p) * P1; h = rol(h, 23) * P2 + P3; p += 4; }
  while (p < end) { h ^= (*p) * P5; h = rol(h, 11) * P1; p++; }
  h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
  return h;
}

extern "C" int mec_selftest(int device) {
  static const struct { uint8_t k, m; uint64_t want; } kGold[] = {  // cmd/erasure-coding.go:160
      {2, 2, 0x23fb21be2496f5d3ull}, {2, 3, 0xa5cd5600ba0d8e7cull}, {3, 1, 0x60ab052148b010b4ull}, {3, 2, 0xe64927daef76435aull},
      {3, 3, 0x672f6f242b227b21ull}, {3, 4, 0x571e41ba23a6dc6ull}, {4, 1, 0x524eaa814d5d86e2ull}, {4, 2, 0x62b9552945504fefull},
      {4, 3, 0xcbf9065ee053e518ull}, {4, 4, 0x9a07581dcd03da8ull}, {4, 5, 0xbf2d27b55370113full}, {5, 1, 0xf71031a01d70dafull},
      {5, 2, 0x8e5845859939d0f4ull}, {5, 3, 0x7ad9161acbb4c325ull}, {5, 4, 0xc446b88830b4f800ull}, {5, 5, 0xabf1573cc6f76165ull},
      {5, 6, 0x7b5598a85045bfb8ull}, {6, 1, 0xe2fc1e677cc7d872ull}, {6, 2, 0x7ed133de5ca6a58eull}, {6, 3, 0x39ef92d0a74cc3c0ull},
      {6, 4, 0xcfc90052bc25d20ull}, {6, 5, 0x71c96f6baeef9c58ull}, {6, 6, 0x4b79056484883e4cull}, {6, 7, 0xb1a0e2427ac2dc1aull},
      {7, 1, 0x937ba2b7af467a22ull}, {7, 2, 0x5fd13a734d27d37aull}, {7, 3, 0x3be2722d9b66912full}, {7, 4, 0x14c628e59011be3dull},
      {7, 5, 0xcc3b39ad4c083b9full}, {7, 6, 0x45af361b7de7a4ffull}, {7, 7, 0x456cc320cec8a6e6ull}, {7, 8, 0x1867a9f4db315b5cull},
      {8, 1, 0xbc5756b9a9ade030ull}, {8, 2, 0xdfd7d9d0b3e36503ull}, {8, 3, 0x72bb72c2cdbcf99dull}, {8, 4, 0x3ba5e9b41bf07f0ull},
      {8, 5, 0xd7dabc15800f9d41ull}, {8, 6, 0xb482a6169fd270full}, {8, 7, 0x50748e0099d657e8ull}, {9, 1, 0xc77ae0144fcaeb6eull},
      {9, 2, 0x8a86c7dbebf27b68ull}, {9, 3, 0xa64e3be6d6fe7e92ull}, {9, 4, 0x239b71c41745d207ull}, {9, 5, 0x2d0803094c5a86ceull},
      {9, 6, 0xa3c2539b3af84874ull}, {10, 1, 0x7d30d91b89fcec21ull}, {10, 2, 0xfa5af9aa9f1857a3ull}, {10, 3, 0x84bc4bda8af81f90ull},
      {10, 4, 0x6c1cba8631de994aull}, {10, 5, 0x4383e58a086cc1acull}, {11, 1, 0x4ed2929a2df690bull}, {11, 2, 0xecd6f1b1399775c0ull},
      {11, 3, 0xc78cfbfc0dc64d01ull}, {11, 4, 0xb2643390973702d6ull}, {12, 1, 0x3b2a88686122d082ull}, {12, 2, 0xfd2f30a48a8e2e9ull},
      {12, 3, 0xd5ce58368ae90b13ull}, {13, 1, 0x9c88e2a9d1b8fff8ull}, {13, 2, 0xcb8460aa4cf6613ull}, {14, 1, 0x78a28bbaec57996eull}};
  uint8_t test[256];
  for (int i = 0; i < 256; i++) test[i] = static_cast<uint8_t>(i);
  for (const auto& g : kGold) {
    mec_codec* c = nullptr;
    int rc = mec_codec_new(g.k, g.m, 1 << 20, MEC_HIGHWAYHASH256S, device, &c);
    if (rc) return rc;
    const int n = g.k + g.m;
    const int64_t per = ceil_frac(256, g.k);
    std::vector<uint8_t> store(static_cast<size_t>(n * per), 0);
    memcpy(store.data(), test, 256);  // Split: zero padded
    std::vector<uint8_t*> sh(n);
    for (int i = 0; i < n; i++) sh[i] = store.data() + i * per;
    rc = mec_rs_encode_shards(c, sh.data(), per);
    if (rc) { mec_codec_free(c); return rc; }
    std::vector<uint8_t> buf;
    for (int i = 0; i < n; i++) { buf.push_back(static_cast<uint8_t>(i)); buf.insert(buf.end(), sh[i], sh[i] + per); }
    if (xxh64(buf.data(), buf.size()) != g.want) {
      set_last_error("erasure self-test mismatch for d:" + std::to_string(g.k) + " p:" + std::to_string(g.m));
      mec_codec_free(c);
      return MEC_ERR_UNEXPECTED;
    }
    std::vector<uint8_t> first(sh[0], sh[0] + per), present(n, 1);
    memset(sh[0], 0xee, static_cast<size_t>(per));
    present[0] = 0;
    rc = mec_rs_reconstruct_shards(c, sh.data(), present.data(), per, 1);
    if (rc == 0 && memcmp(first.data(), sh[0], static_cast<size_t>(per)) != 0) rc = MEC_ERR_UNEXPECTED;
    mec_codec_free(c);
    if (rc) { set_last_error("erasure self-test: first-shard reconstruct failed"); return rc; }
  }
  // bitrot self-test: HighwayHash chain (cmd/bitrot.go:228): msg grows by its own digest
  {
    mec_codec* c = nullptr;
    int rc = mec_codec_new(2, 2, 1 << 20, MEC_HIGHWAYHASH256S, device, &c);
    if (rc) return rc;
    static const uint8_t want[32] = {0x39, 0xc0, 0x40, 0x7e, 0xd3, 0xf0, 0x1b, 0x18, 0xd2, 0x2c, 0x85, 0xdb, 0x4a, 0xef, 0xf1, 0x1e,
                                     0x06, 0x0c, 0xa5, 0xf4, 0x31, 0x31, 0xb0, 0x12, 0x67, 0x31, 0xca, 0x19, 0x7c, 0xd4, 0x23, 0x13};
    std::vector<uint8_t> msg;
    uint8_t sum[32];
    for (int i = 0; i < 32 * 32; i += 32) {
      rc = mec_hh256_batch(c, msg.data(), static_cast<int64_t>(msg.size()), 1, sum);
      if (rc) { mec_codec_free(c); return rc; }
      msg.insert(msg.end(), sum, sum + 32);
    }
    mec_codec_free(c);
    if (memcmp(sum, want, 32) != 0) { set_last_error("bitrot self-test mismatch"); return MEC_ERR_UNEXPECTED; }
  }
  // the whole-file algorithms of the same self test (cmd/bitrot.go:225-229): Size()*BlockSize() bytes, Size() at a time
  {
    static const struct { int algo, size, block; const char* want; } kChains[] = {
        {MEC_SHA256, 32, 64, "a7677ff19e0182e4d52e3a3db727804abc82a5818749336369552e54b838b004"},
        {MEC_BLAKE2B512, 64, 128, "e519b7d84b1c3c917985f544773a35cf265dcab10948be3550320d156bab612124a5ae2ae5a8c73c0eea360f68b0e28136f26e858756dbfe7375a7389f26c669"},
        {MEC_HIGHWAYHASH256, 32, 32, "39c0407ed3f01b18d22c85db4aeff11e060ca5f43131b0126731ca197cd42313"}};
    for (const auto& ch : kChains) {
      mec_codec* c = nullptr;
      int rc = mec_codec_new(2, 2, 1 << 20, ch.algo, device, &c);
      if (rc) return rc;
      std::vector<uint8_t> msg;
      uint8_t sum[64];
      for (int i = 0; i < ch.size * ch.block; i += ch.size) {
        rc = mec_whole_hash(c, ch.algo, msg.data(), static_cast<int64_t>(msg.size()), 1, sum);
        if (rc) { mec_codec_free(c); return rc; }
        msg.insert(msg.end(), sum, sum + ch.size);
      }
      mec_codec_free(c);
      for (int i = 0; i < ch.size; i++) {
        unsigned v = 0;
        sscanf(ch.want + 2 * i, "%2x", &v);
        if (sum[i] != v) { set_last_error("bitrot self-test mismatch (whole-file algorithm " + std::to_string(ch.algo) + ")"); return MEC_ERR_UNEXPECTED; }
      }
    }
  }
  return MEC_OK;
}

// ------------------------------------------------------------------------------------------------
// Cross-request coalescer.  MinIO builds one Erasure per request (cmd/erasure-object.go:1371) and encodes ONE block per loop
// iteration (cmd/erasure-encode.go:76-108) on thousands of concurrent request goroutines (cmd/handler-api.go:139-143).  A GPU
// launch that carries one 1 MiB block keeps one CTA busy on one of 148 SMs for as long as a launch of a thousand blocks takes.
// The batcher turns concurrency into batch size: callers block in mec_batcher_encode*, a worker thread merges whatever is
// queued (same k, m, block size by construction) into ONE staged buffer, ONE launch over all full blocks (plus one small
// launch per short tail block), and one pass of DMA back into every caller's own frames; up to three merged batches are in
// flight on the codec's three slots, so staging, kernel and copy-back of consecutive batches overlap.
#include <chrono>
#include <condition_variable>
#include <deque>

struct BatchReq {
  int kind = 0;                      // 0 = PutObject (encode), 1 = GetObject (decode)
  // GET
  const uint8_t* const* gfiles = nullptr;
  int64_t offset = 0, length = 0, total = 0;
  uint8_t* dst = nullptr;
  int* heal_hint = nullptr;
  uint64_t alive_mask = 0;           // bit i = reader i is online: requests with equal masks share a reader set and a launch
  int64_t get_slot0 = 0, get_tail_slot = -1, get_nblocks = 0;  // placement of the request's blocks in the merged batch
  // PUT
  const uint8_t* src = nullptr;
  int64_t len = 0;
  uint8_t* const* files = nullptr;
  uint8_t* data_digests = nullptr;
  bool with_data = false;
  bool mapped = false;   // every buffer of the call is page-locked (device-addressable): eligible for the gather / scatter kernels
  int64_t result = 0;
  bool done = false;
  std::condition_variable cv;  // one per request: a finished batch wakes its own callers, not every thread parked in the batcher
  EncChunk ch;   // placement inside the merged batch
};

struct mec_batcher {
  mec_codec* codec = nullptr;
  mec_codec* serial = nullptr;       // private handle for calls that cannot ride in a batch (pageable buffers, bitrot fail-over)
  std::mutex serial_mu;
  int64_t max_blocks = 256;
  int max_wait_us = 200;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::deque<BatchReq*> queue;
  bool stop = false;
  std::thread worker;
  std::atomic<int64_t> st_batches{0}, st_requests{0}, st_blocks{0}, st_kernel_batches{0};
  // hybrid staging: part of a batch is read by the gather kernel on `aux` while the copy engine moves the rest
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_table[kSlots] = {}, ev_gather[kSlots] = {};
  int gather_pct = 0;
  // where the worker's time and each batch's stream time go (mec_batcher_stat "us_*"); device phases need MEC_BATCHER_TRACE=1
  bool trace = false;
  cudaEvent_t ev_ph[kSlots][4] = {};
  std::atomic<int64_t> us_submit{0}, us_sync{0}, us_finish{0}, us_idle{0}, us_stage{0}, us_kernel{0}, us_scatter{0};
};

namespace {
struct InflightBatch {
  std::vector<BatchReq*> reqs;
  bool busy = false;
  bool needs_retire = true;  // false: the scatter kernel already assembled the frames
  bool is_get = false;
  bool traced = false;
  int get_k = 0;
};

static mec_codec* batcher_serial(mec_batcher* b) {
  std::lock_guard<std::mutex> lk(b->serial_mu);
  if (!b->serial) {
    mec_codec* c = nullptr;
    if (mec_codec_new(b->codec->k, b->codec->m, b->codec->block_size, MEC_HIGHWAYHASH256S, b->codec->device, &c) == MEC_OK) b->serial = c;
  }
  return b->serial;
}

static void batcher_finish(mec_batcher* b, InflightBatch& fb, int rc) {
  mec_codec* c = b->codec;
  if (fb.is_get) {
    // per-frame digest verdicts of the merged launch: a request that met a corrupt frame is redone on its own through the serial
    // handle, which applies parallelReader's block-sequential fail-over exactly; the others are done (their bytes are in place)
    Slot& s = *fb.reqs.front()->ch.s;
    const uint8_t* fl = static_cast<const uint8_t*>(s.hflags.p);
    const int k = fb.get_k;
    for (BatchReq* r : fb.reqs) {
      int64_t res = rc == MEC_OK ? r->length : rc;
      if (rc == MEC_OK) {
        bool bad = false;
        for (int64_t bb = 0; bb < r->get_nblocks && !bad; bb++) {
          const int64_t slot = (r->get_tail_slot >= 0 && bb == r->get_nblocks - 1) ? r->get_tail_slot : r->get_slot0 + bb;
          for (int t = 0; t < k; t++) bad |= fl[slot * k + t] != 0;
        }
        if (bad) {
          mec_codec* sc = batcher_serial(b);
          res = sc ? mec_decode_prefer(sc, r->gfiles, nullptr, r->offset, r->length, r->total, r->dst, r->heal_hint) : MEC_ERR_UNEXPECTED;
        } else if (r->heal_hint) {
          *r->heal_hint = 0;
        }
      }
      r->result = res;
    }
    {
      std::lock_guard<std::mutex> lk(b->mu);
      for (BatchReq* r : fb.reqs) { r->done = true; r->cv.notify_one(); }
    }
    fb.reqs.clear();
    fb.busy = false;
    fb.is_get = false;
    return;
  }
  for (BatchReq* r : fb.reqs) {
    if (rc == MEC_OK && fb.needs_retire) frames_retire(c, r->ch, r->files, r->with_data, r->data_digests);
    r->result = rc == MEC_OK ? r->len : rc;
  }
  {
    std::lock_guard<std::mutex> lk(b->mu);
    for (BatchReq* r : fb.reqs) { r->done = true; r->cv.notify_one(); }  // under the lock: `r` lives on its caller's stack
  }
  fb.reqs.clear();
  fb.busy = false;
}

// stage, launch and enqueue the copy-back of one merged batch on slot `s`
static int batcher_submit(mec_batcher* b, Slot& s, std::vector<BatchReq*>& reqs, bool* needs_retire) {
  mec_codec* c = b->codec;
  const int k = c->k, m = c->m, n = c->n;
  const int64_t bs = c->block_size, S = c->S(), pitch = round_up(S, 16);
  *needs_retire = true;
  bool all_mapped = n <= kBatchMaxFiles && getenv("MEC_BATCHER_MEMCPY") == nullptr;
  for (BatchReq* r : reqs) all_mapped &= r->mapped;
  int64_t total_full = 0, ntails = 0, tail_bytes = 0;
  for (BatchReq* r : reqs) {
    total_full += r->len / bs;
    if (r->len % bs) { ntails++; tail_bytes += round_up(r->len % bs, 256); }
  }
  const int64_t nslots = total_full + ntails;
  int rc;
  if ((rc = s.src.ensure(static_cast<size_t>(total_full * bs + tail_bytes + 512)))) return rc;
  if ((rc = s.out.ensure(static_cast<size_t>(std::max<int64_t>(nslots, 1) * std::max(m, 1) * pitch)))) return rc;
  if ((rc = s.dig.ensure(static_cast<size_t>(std::max<int64_t>(nslots, 1) * n * 32)))) return rc;
  if ((rc = s.hdig.ensure(static_cast<size_t>(std::max<int64_t>(nslots, 1) * n * 32)))) return rc;
  uint8_t* dsrc = static_cast<uint8_t*>(s.src.p);
  uint8_t* dout = static_cast<uint8_t*>(s.out.p);
  uint8_t* ddig = static_cast<uint8_t*>(s.dig.p);
  const uint8_t* hdig = static_cast<const uint8_t*>(s.hdig.p);
  int64_t fb = 0, tb = total_full, toff = total_full * bs;
  if (all_mapped) {
    // gather kernel -> ONE fused launch over all full blocks (+ one per tail) -> scatter kernel: a handful of launches and one
    // small table copy per batch, whatever the number of requests
    if ((rc = s.hreq.ensure(reqs.size() * sizeof(BatchReqDesc)))) return rc;
    if ((rc = s.dreq.ensure(reqs.size() * sizeof(BatchReqDesc)))) return rc;
    BatchReqDesc* tab = static_cast<BatchReqDesc*>(s.hreq.p);
    int64_t max_len = 0, max_nb = 0;
    for (size_t q = 0; q < reqs.size(); q++) {
      BatchReq* r = reqs[q];
      BatchReqDesc& d = tab[q];
      memset(&d, 0, sizeof(d));
      const int64_t nf = r->len / bs, tl = r->len % bs;
      d.src = r->src; d.len = r->len; d.data_digests = r->data_digests; d.with_data = r->with_data ? 1 : 0;
      for (int i = 0; i < n; i++) d.files[i] = r->files[i];
      d.full0 = fb; d.tail_slot = tl ? tb : -1; d.tail_src_off = tl ? toff : 0;
      r->ch = EncChunk();
      r->ch.nfull = nf; r->ch.tail = tl;
      r->ch.d_src_tail = dsrc + toff; r->ch.d_out_tail = dout + tb * m * pitch; r->ch.h_dig_tail = hdig + tb * n * 32;
      fb += nf;
      if (tl) { tb++; toff += round_up(tl, 256); }
      max_len = std::max(max_len, r->len);
      max_nb = std::max(max_nb, nf + (tl ? 1 : 0));
      c->st_h2d += r->len;
      c->st_d2h += (nf * S + (tl ? ceil_frac(tl, k) : 0)) * (r->with_data ? n : m);
    }
    const int phase_slot = static_cast<int>(&s - c->slots);
    if (b->trace) cudaEventRecord(b->ev_ph[phase_slot][0], s.st);
    MEC_CUDA_OK(cudaMemcpyAsync(s.dreq.p, tab, reqs.size() * sizeof(BatchReqDesc), cudaMemcpyHostToDevice, s.st));
    BatchCopyParams cp;
    cp.reqs = static_cast<const BatchReqDesc*>(s.dreq.p);
    cp.nreq = static_cast<int>(reqs.size()); cp.k = k; cp.m = m; cp.bs = bs; cp.S = S; cp.pitch = pitch;
    cp.staged = dsrc; cp.parity = dout; cp.digests = ddig;
    // staging: the copy engines are the faster reader of host memory (one call per request); the gather kernel (SM loads over
    // PCIe, no per-request call at all) is kept for hosts where the worker thread is the bottleneck — measured in profiles/r2_concurrency.md
    static const bool use_gather = getenv("MEC_BATCHER_GATHER") != nullptr;
    // requests [0, ndma) go through the copy engine, [ndma, nreq) through the gather kernel on the aux stream, concurrently
    size_t ndma = use_gather ? 0 : reqs.size();
    if (!use_gather && b->gather_pct > 0 && b->aux && reqs.size() >= 8)
      ndma = reqs.size() - reqs.size() * static_cast<size_t>(b->gather_pct) / 100;
    if (ndma < reqs.size()) {
      const int si = static_cast<int>(&s - c->slots);
      cudaStream_t gs = s.st;
      if (ndma > 0) {
        gs = b->aux;
        MEC_CUDA_OK(cudaEventRecord(b->ev_table[si], s.st));
        MEC_CUDA_OK(cudaStreamWaitEvent(gs, b->ev_table[si], 0));
      }
      int64_t glen = 0;
      for (size_t q = ndma; q < reqs.size(); q++) glen = std::max(glen, tab[q].len);
      BatchCopyParams gp = cp;
      gp.reqs = cp.reqs + ndma;
      gp.nreq = static_cast<int>(reqs.size() - ndma);
      const unsigned gx = static_cast<unsigned>(std::min<int64_t>(64, std::max<int64_t>(1, ceil_frac(glen, 32 << 10))));
      batch_gather_kernel<<<dim3(gx, static_cast<unsigned>(gp.nreq)), 256, 0, gs>>>(gp);
      MEC_CUDA_OK(cudaGetLastError());
      c->eng->count_launch();
      if (ndma > 0) MEC_CUDA_OK(cudaEventRecord(b->ev_gather[si], gs));
    }
    {
      // ONE cudaMemcpyBatchAsync for the whole batch: the copy engine runs a train of 1 MiB pieces at the link rate (55 GB/s) where
      // the same pieces as separate cudaMemcpyAsync calls reach 40-46 GB/s (tools/copy_probe.cu, profiles/r2_concurrency.md)
      std::vector<void*> dsts, srcs;
      std::vector<size_t> sizes;
      for (size_t q = 0; q < ndma; q++) {
        const BatchReqDesc& d = tab[q];
        const int64_t fbytes = d.len / bs * bs, tl = d.len - fbytes;
        if (fbytes > 0) { dsts.push_back(dsrc + d.full0 * bs); srcs.push_back(const_cast<uint8_t*>(d.src)); sizes.push_back(static_cast<size_t>(fbytes)); }
        if (tl > 0) { dsts.push_back(dsrc + d.tail_src_off); srcs.push_back(const_cast<uint8_t*>(d.src) + fbytes); sizes.push_back(static_cast<size_t>(tl)); }
      }
      if ((rc = copy_batch(dsts, srcs, sizes, s.st))) return rc;
    }
    if (ndma > 0 && ndma < reqs.size()) MEC_CUDA_OK(cudaStreamWaitEvent(s.st, b->ev_gather[static_cast<int>(&s - c->slots)], 0));
    if (b->trace) cudaEventRecord(b->ev_ph[phase_slot][1], s.st);
    if (c->eng->small_ok(c->opt, nslots)) {
      // ONE launch for every block of the batch, full or short (objects smaller than a block are nothing but a short block): the
      // latency kernel takes the geometry of each block from a table
      if ((rc = s.hblk.ensure(static_cast<size_t>(nslots) * sizeof(SmallBlock)))) return rc;
      if ((rc = s.blk.ensure(static_cast<size_t>(nslots) * sizeof(SmallBlock)))) return rc;
      SmallBlock* bt = static_cast<SmallBlock*>(s.hblk.p);
      for (int64_t q = 0; q < total_full; q++) bt[q] = SmallBlock{q * bs, static_cast<int32_t>(S), static_cast<int32_t>(bs)};
      for (size_t q = 0; q < reqs.size(); q++)
        if (tab[q].tail_slot >= 0) {
          const int64_t tl = tab[q].len % bs;
          bt[tab[q].tail_slot] = SmallBlock{tab[q].tail_src_off, static_cast<int32_t>(ceil_frac(tl, k)), static_cast<int32_t>(tl)};
        }
      MEC_CUDA_OK(cudaMemcpyAsync(s.blk.p, bt, static_cast<size_t>(nslots) * sizeof(SmallBlock), cudaMemcpyHostToDevice, s.st));
      FusedDesc d;
      d.k = k; d.r = m;
      d.coef = c->matrix.data() + static_cast<size_t>(k) * k;
      d.static_encode = true; d.contiguous = true; d.key = kMagicKey; d.out_pitch = pitch;
      d.nblocks = nslots; d.S = static_cast<int32_t>(S);
      d.in_base = dsrc; d.in_block_stride = bs; d.in_block_len = bs;
      d.out = dout; d.digests = ddig;
      d.blocks = static_cast<const SmallBlock*>(s.blk.p);
      if ((rc = c->eng->launch_fused(d, c->opt, s.st))) return rc;
    } else {
      if (total_full > 0 && (rc = encode_device_locked(c, dsrc, total_full * bs, dout, pitch, ddig, s.st))) return rc;
      for (BatchReq* r : reqs)
        if (r->ch.tail > 0 && (rc = encode_device_locked(c, r->ch.d_src_tail, r->ch.tail, r->ch.d_out_tail, pitch, ddig + (r->ch.h_dig_tail - hdig), s.st)))
          return rc;
    }
    if (b->trace) cudaEventRecord(b->ev_ph[phase_slot][2], s.st);
    batch_scatter_kernel<<<dim3(static_cast<unsigned>(std::min<int64_t>(max_nb, 16)), static_cast<unsigned>(n), static_cast<unsigned>(reqs.size())), 128, 0, s.st>>>(cp);
    MEC_CUDA_OK(cudaGetLastError());
    c->eng->count_launch();
    if (b->trace) cudaEventRecord(b->ev_ph[phase_slot][3], s.st);
    *needs_retire = false;
    c->st_blocks_encoded += nslots;
    b->st_batches++;
    b->st_kernel_batches++;
    b->st_requests += static_cast<int64_t>(reqs.size());
    b->st_blocks += nslots;
    return MEC_OK;
  }
  for (BatchReq* r : reqs) {
    EncChunk& ch = r->ch;
    ch = EncChunk();
    ch.b0 = 0; ch.nfull = r->len / bs; ch.tail = r->len % bs; ch.nb = ch.nfull + (ch.tail ? 1 : 0);
    ch.s = &s; ch.st = s.st; ch.pitch = pitch;
    ch.d_src = dsrc + fb * bs; ch.d_out = dout + fb * m * pitch; ch.h_dig = hdig + fb * n * 32;
    if (ch.nfull > 0)
      MEC_CUDA_OK(cudaMemcpyAsync(dsrc + fb * bs, r->src, static_cast<size_t>(ch.nfull * bs), cudaMemcpyHostToDevice, s.st));
    if (ch.tail > 0) {
      ch.d_src_tail = dsrc + toff; ch.d_out_tail = dout + tb * m * pitch; ch.h_dig_tail = hdig + tb * n * 32;
      MEC_CUDA_OK(cudaMemcpyAsync(dsrc + toff, r->src + ch.nfull * bs, static_cast<size_t>(ch.tail), cudaMemcpyHostToDevice, s.st));
      toff += round_up(ch.tail, 256);
      tb++;
    }
    fb += ch.nfull;
    c->st_h2d += r->len;
  }
  if (total_full > 0 && (rc = encode_device_locked(c, dsrc, total_full * bs, dout, pitch, ddig, s.st))) return rc;  // ONE launch for every caller's full blocks
  for (BatchReq* r : reqs) {
    const EncChunk& ch = r->ch;
    if (ch.tail > 0 && (rc = encode_device_locked(c, ch.d_src_tail, ch.tail, ch.d_out_tail, pitch,
                                                  ddig + (ch.h_dig_tail - hdig), s.st)))
      return rc;
  }
  MEC_CUDA_OK(cudaMemcpyAsync(s.hdig.p, s.dig.p, static_cast<size_t>(nslots * n * 32), cudaMemcpyDeviceToHost, s.st));
  for (BatchReq* r : reqs)
    if ((rc = frames_enqueue(c, r->ch, r->files, r->with_data, r->data_digests))) return rc;
  c->st_blocks_encoded += nslots;
  b->st_batches++;
  b->st_requests += static_cast<int64_t>(reqs.size());
  b->st_blocks += nslots;
  (void)k;
  return MEC_OK;
}

// one merged launch for the GETs of a batch that share a reader set (alive mask): gather kernel -> fused reconstruct over all full
// blocks (+ one launch per short last block) -> scatter kernel into every caller's destination
static int batcher_submit_get(mec_batcher* b, Slot& s, std::vector<BatchReq*>& reqs) {
  mec_codec* c = b->codec;
  const int k = c->k, n = c->n;
  const int64_t bs = c->block_size, S = c->S(), P = round_up(32 + S, 16), pitch = round_up(S, 16);
  // reader set and decode rows of the shared pattern
  int chosen[kMaxShards], nch = 0;
  const uint64_t mask = reqs.front()->alive_mask;
  for (int i = 0; i < n && nch < k; i++)
    if (mask & (1ull << i)) chosen[nch++] = i;
  if (nch < k) return MEC_ERR_READ_QUORUM;
  std::vector<uint8_t> present(n, 0);
  for (int t = 0; t < k; t++) present[chosen[t]] = 1;
  int targets[kMaxShards], r = 0;
  for (int i = 0; i < k; i++)
    if (!present[i]) targets[r++] = i;
  if (r > kMaxR) return MEC_ERR_UNSUPPORTED;
  std::vector<uint8_t> rows(static_cast<size_t>(std::max(r, 1)) * k);
  int valid[kMaxShards];
  if (r > 0 && !rs_decode_rows(k, c->m, present.data(), targets, r, rows.data(), valid)) return MEC_ERR_TOO_FEW_SHARDS;
  // layout: every request's full blocks first, the short last blocks behind them
  int rc;
  if ((rc = s.hreq.ensure(reqs.size() * sizeof(BatchGetDesc)))) return rc;
  if ((rc = s.dreq.ensure(reqs.size() * sizeof(BatchGetDesc)))) return rc;
  BatchGetDesc* tab = static_cast<BatchGetDesc*>(s.hreq.p);
  int64_t nfull_slots = 0, ntails = 0, max_nb = 0;
  for (size_t q = 0; q < reqs.size(); q++) {
    BatchReq* rq = reqs[q];
    BatchGetDesc& d = tab[q];
    memset(&d, 0, sizeof(d));
    const int64_t sfs = mec_shard_file_size(c, rq->total), nblocks_total = ceil_frac(sfs, S);
    const int64_t start_block = rq->offset / bs;
    int64_t last_block = (rq->offset + rq->length) / bs;
    if ((rq->offset + rq->length) % bs == 0) last_block--;
    if (last_block >= nblocks_total) last_block = nblocks_total - 1;
    const int64_t last_len = std::min(S, sfs - last_block * S);
    d.offset = rq->offset; d.length = rq->length; d.total = rq->total; d.dst = rq->dst;
    for (int i = 0; i < n; i++) d.files[i] = rq->gfiles[i];
    d.start_block = start_block; d.nblocks = last_block - start_block + 1;
    d.last_len = last_len;
    const bool tail = last_len != S;
    d.slot0 = nfull_slots;
    nfull_slots += d.nblocks - (tail ? 1 : 0);
    d.tail_slot = tail ? ntails : -1;  // provisional: offset by the number of full slots below
    if (tail) ntails++;
    max_nb = std::max(max_nb, d.nblocks);
    rq->get_nblocks = d.nblocks;
  }
  static const int64_t dma_blocks = getenv("MEC_BATCHER_DMA_BLOCKS") ? atoll(getenv("MEC_BATCHER_DMA_BLOCKS")) : 1;
  for (size_t q = 0; q < reqs.size(); q++) {
    tab[q].dma = tab[q].nblocks >= dma_blocks ? 1 : 0;
    if (tab[q].tail_slot >= 0) tab[q].tail_slot += nfull_slots;
    reqs[q]->get_slot0 = tab[q].slot0;
    reqs[q]->get_tail_slot = tab[q].tail_slot;
    reqs[q]->ch = EncChunk();
    reqs[q]->ch.s = &s;
  }
  const int64_t nslots = nfull_slots + ntails;
  const int64_t stride = round_up(nslots * P + 512, 256);
  if ((rc = s.aux.ensure(static_cast<size_t>(k * stride)))) return rc;
  if ((rc = s.out.ensure(static_cast<size_t>(nslots * std::max(r, 1) * pitch)))) return rc;
  if ((rc = s.dig.ensure(static_cast<size_t>(nslots * (k + r) * 32)))) return rc;
  if ((rc = s.flags.ensure(static_cast<size_t>(nslots * k)))) return rc;
  if ((rc = s.hflags.ensure(static_cast<size_t>(nslots * k)))) return rc;
  // writeDataBlocks: 0 = the scatter kernel assembles every object in device memory and ONE copy per request takes it to the caller
  // (fewest driver-side copy descriptors), 1 = the scatter kernel stores straight into the callers' buffers over PCIe, 2 = the copy
  // engine moves every (request, shard, block) piece itself — measured in profiles/r2_concurrency.md
  static const int out_mode = getenv("MEC_BATCHER_GET_OUT") ? atoi(getenv("MEC_BATCHER_GET_OUT")) : 0;
  std::vector<int64_t> asm_off(reqs.size(), 0);
  if (out_mode == 0) {
    int64_t total = 0;
    for (size_t q = 0; q < reqs.size(); q++) { asm_off[q] = total; total += round_up(tab[q].length, 256); }
    if ((rc = s.src.ensure(static_cast<size_t>(total + 256)))) return rc;
    for (size_t q = 0; q < reqs.size(); q++) tab[q].dst = static_cast<uint8_t*>(s.src.p) + asm_off[q];
  }
  const int phase_slot = static_cast<int>(&s - c->slots);
  if (b->trace) cudaEventRecord(b->ev_ph[phase_slot][0], s.st);
  MEC_CUDA_OK(cudaMemcpyAsync(s.dreq.p, tab, reqs.size() * sizeof(BatchGetDesc), cudaMemcpyHostToDevice, s.st));
  BatchGetParams gp;
  memset(&gp, 0, sizeof(gp));
  gp.reqs = static_cast<const BatchGetDesc*>(s.dreq.p);
  gp.nreq = static_cast<int>(reqs.size()); gp.k = k; gp.r = r;
  for (int t = 0; t < k; t++) gp.chosen[t] = chosen[t];
  for (int q = 0; q < r; q++) gp.targets[q] = targets[q];
  gp.bs = bs; gp.S = S; gp.P = P;
  gp.arena = static_cast<uint8_t*>(s.aux.p); gp.arena_stride = stride;
  gp.rebuilt = static_cast<const uint8_t*>(s.out.p); gp.pitch = pitch;
  const unsigned gx = static_cast<unsigned>(std::min<int64_t>(std::max<int64_t>(max_nb, 1), 16));
  // long ranges are staged by the copy engines (k strided copies per request are cheap next to megabytes of frames); short ones by
  // the gather kernel (no per-request driver call at all)
  const int64_t fstride = 32 + S;
  {
    // one 1-D copy per (request, survivor, block) frame — the arena pads frames to a 16-byte pitch — all of them in ONE batched call
    std::vector<void*> dsts, srcs;
    std::vector<size_t> sizes;
    for (size_t q = 0; q < reqs.size(); q++) {
      const BatchGetDesc& d = tab[q];
      if (!d.dma) continue;
      const int64_t nf = d.nblocks - (d.tail_slot >= 0 ? 1 : 0);
      for (int t = 0; t < k; t++) {
        const uint8_t* f = d.files[chosen[t]] + d.start_block * fstride;
        uint8_t* a = gp.arena + t * stride;
        for (int64_t j = 0; j < nf; j++) {
          dsts.push_back(a + (d.slot0 + j) * P); srcs.push_back(const_cast<uint8_t*>(f) + j * fstride); sizes.push_back(static_cast<size_t>(fstride));
        }
        if (d.tail_slot >= 0) {
          dsts.push_back(a + d.tail_slot * P); srcs.push_back(const_cast<uint8_t*>(f) + nf * fstride); sizes.push_back(static_cast<size_t>(32 + d.last_len));
        }
      }
    }
    if ((rc = copy_batch(dsts, srcs, sizes, s.st))) return rc;
  }
  bool any_gather = false;
  for (size_t q = 0; q < reqs.size(); q++) any_gather |= !tab[q].dma;
  if (any_gather) {
    batch_get_gather_kernel<<<dim3(gx, static_cast<unsigned>(k), static_cast<unsigned>(reqs.size())), 128, 0, s.st>>>(gp);
    MEC_CUDA_OK(cudaGetLastError());
    c->eng->count_launch();
  }
  if (b->trace) cudaEventRecord(b->ev_ph[phase_slot][1], s.st);
  MEC_CUDA_OK(cudaMemsetAsync(s.flags.p, 0, static_cast<size_t>(nslots * k), s.st));
  const int32_t* block_len = nullptr;
  auto launch = [&](int64_t slot0, int64_t cnt, int64_t shard_len) -> int {
    FusedDesc d;
    d.block_len = block_len;
    d.k = k; d.r = r; d.coef = rows.data(); d.static_encode = false; d.contiguous = false; d.hash_outputs = false;
    d.key = kMagicKey; d.out_pitch = pitch; d.expect_block_stride = P; d.in_block_stride = P;
    d.nblocks = cnt; d.S = static_cast<int32_t>(shard_len);
    for (int t = 0; t < k; t++) {
      const uint8_t* base = gp.arena + t * stride + slot0 * P;
      d.map_base[t] = base; d.map_len[t] = stride - slot0 * P; d.expect_ptr[t] = base; d.in_ptr[t] = base + 32;
    }
    d.out = static_cast<uint8_t*>(s.out.p) + slot0 * r * pitch;
    d.digests = static_cast<uint8_t*>(s.dig.p) + slot0 * (k + r) * 32;
    d.corrupt = static_cast<uint8_t*>(s.flags.p) + slot0 * k;
    return c->eng->launch_fused(d, c->opt, s.st);
  };
  if (ntails > 0 && c->eng->small_ok(c->opt, nslots)) {
    // ONE launch for every block of every caller, the short last ones included (per-block shard lengths in the latency kernel)
    if ((rc = s.hblk.ensure(static_cast<size_t>(nslots) * sizeof(int32_t)))) return rc;
    if ((rc = s.blk.ensure(static_cast<size_t>(nslots) * sizeof(int32_t)))) return rc;
    int32_t* lens = static_cast<int32_t*>(s.hblk.p);
    for (int64_t q = 0; q < nfull_slots; q++) lens[q] = static_cast<int32_t>(S);
    for (size_t q = 0; q < reqs.size(); q++)
      if (tab[q].tail_slot >= 0) lens[tab[q].tail_slot] = static_cast<int32_t>(tab[q].last_len);
    MEC_CUDA_OK(cudaMemcpyAsync(s.blk.p, lens, static_cast<size_t>(nslots) * sizeof(int32_t), cudaMemcpyHostToDevice, s.st));
    block_len = static_cast<const int32_t*>(s.blk.p);
    if ((rc = launch(0, nslots, S))) return rc;
  } else {
    if (nfull_slots > 0 && (rc = launch(0, nfull_slots, S))) return rc;  // ONE launch for every caller's full blocks
    for (size_t q = 0; q < reqs.size(); q++)
      if (tab[q].tail_slot >= 0 && (rc = launch(tab[q].tail_slot, 1, tab[q].last_len))) return rc;
  }
  MEC_CUDA_OK(cudaMemcpyAsync(s.hflags.p, s.flags.p, static_cast<size_t>(nslots * k), cudaMemcpyDeviceToHost, s.st));
  if (b->trace) cudaEventRecord(b->ev_ph[phase_slot][2], s.st);
  if (out_mode == 2) {
    // writeDataBlocks by the copy engine: one 1-D copy per (request, data shard, block) piece, all in one batched call
    std::vector<void*> dsts, srcs;
    std::vector<size_t> sizes;
    int tpos[kMaxShards], qpos[kMaxShards];
    for (int i = 0; i < k; i++) { tpos[i] = qpos[i] = -1; }
    for (int t = 0; t < k; t++) if (chosen[t] < k) tpos[chosen[t]] = t;
    for (int q = 0; q < r; q++) qpos[targets[q]] = q;
    for (size_t rq = 0; rq < reqs.size(); rq++) {
      const BatchGetDesc& d = tab[rq];
      const int64_t hi = d.offset + d.length;
      for (int64_t bb = 0; bb < d.nblocks; bb++) {
        const bool is_tail = d.tail_slot >= 0 && bb == d.nblocks - 1;
        const int64_t cur = is_tail ? d.last_len : S, slot = is_tail ? d.tail_slot : d.slot0 + bb;
        const int64_t B = d.start_block + bb, blo = B * bs, bhi = std::min(blo + bs, d.total);
        for (int i = 0; i < k; i++) {
          const int64_t slo = blo + static_cast<int64_t>(i) * cur, shi = std::min(slo + cur, bhi);
          const int64_t a = std::max(slo, d.offset), e = std::min(shi, hi);
          if (e <= a) continue;
          const uint8_t* src = tpos[i] >= 0 ? gp.arena + tpos[i] * stride + slot * P + 32 : gp.rebuilt + (slot * r + qpos[i]) * pitch;
          dsts.push_back(d.dst + (a - d.offset)); srcs.push_back(const_cast<uint8_t*>(src) + (a - slo)); sizes.push_back(static_cast<size_t>(e - a));
        }
      }
    }
    if ((rc = copy_batch(dsts, srcs, sizes, s.st))) return rc;
  } else {
    batch_get_scatter_kernel<<<dim3(gx, static_cast<unsigned>(k), static_cast<unsigned>(reqs.size())), 128, 0, s.st>>>(gp);
    MEC_CUDA_OK(cudaGetLastError());
    c->eng->count_launch();
    if (out_mode == 0) {
      std::vector<void*> dsts, srcs;
      std::vector<size_t> sizes;
      for (size_t q = 0; q < reqs.size(); q++) {
        if (tab[q].length <= 0) continue;
        dsts.push_back(reqs[q]->dst); srcs.push_back(static_cast<uint8_t*>(s.src.p) + asm_off[q]); sizes.push_back(static_cast<size_t>(tab[q].length));
      }
      if ((rc = copy_batch(dsts, srcs, sizes, s.st))) return rc;
    }
  }
  if (b->trace) cudaEventRecord(b->ev_ph[phase_slot][3], s.st);
  c->st_blocks_read += nslots;
  c->st_shards_rebuilt += nslots * r;
  b->st_batches++;
  b->st_kernel_batches++;
  b->st_requests += static_cast<int64_t>(reqs.size());
  b->st_blocks += nslots;
  return MEC_OK;
}

static void batcher_main(mec_batcher* b) {
  mec_codec* c = b->codec;
  cudaSetDevice(c->device);
  InflightBatch inflight[kSlots];
  int head = 0, tail = 0, nbusy = 0;
  using clk = std::chrono::steady_clock;
  auto us_since = [](clk::time_point t0) { return std::chrono::duration_cast<std::chrono::microseconds>(clk::now() - t0).count(); };
  for (;;) {
    std::vector<BatchReq*> reqs;
    {
      std::unique_lock<std::mutex> lk(b->mu);
      const auto t_idle = clk::now();
      if (nbusy == 0) b->cv_work.wait(lk, [&] { return b->stop || !b->queue.empty(); });
      b->us_idle += us_since(t_idle);
      if (b->stop && b->queue.empty() && nbusy == 0) return;
      if (!b->queue.empty() && nbusy < kSlots) {
        // gather: whatever is queued now; when nothing is in flight, give concurrent callers max_wait_us to join
        auto req_blocks = [&](const BatchReq* r) -> int64_t {
          return r->kind == 0 ? ceil_frac(r->len, c->block_size) : (r->length / c->block_size + 2);
        };
        auto blocks_queued = [&] {
          int64_t nb = 0;
          for (BatchReq* r : b->queue) nb += req_blocks(r);
          return nb;
        };
        // an idle GPU gives the first caller up to max_wait_us of company before its batch is cut
        // (measured, profiles/r2_concurrency.md: waiting while batches are in flight only adds latency to a closed loop of callers)
        if (nbusy == 0 && b->max_wait_us > 0 && blocks_queued() < b->max_blocks)
          b->cv_work.wait_for(lk, std::chrono::microseconds(b->max_wait_us), [&] { return b->stop || blocks_queued() >= b->max_blocks; });
        // a batch = the oldest request plus everything queued behind it of the same kind (and, for GETs, the same reader set)
        // a shallow pipeline is filled with several smaller batches rather than one big one: the fused kernel takes ~0.6 ms however few
        // blocks it is given (one warp walks one erasure block), so the copy of one batch has to overlap the kernel of another —
        // otherwise a closed loop of callers falls into a convoy (all of them in one batch, the GPU idle while they wake up)
        static const int depth = getenv("MEC_BATCHER_DEPTH") ? atoi(getenv("MEC_BATCHER_DEPTH")) : 4;
        int64_t cap = b->max_blocks;
        if (depth > 0 && nbusy < depth - 1) {
          const int64_t q = blocks_queued();
          if (q > 8) cap = std::min(cap, std::max<int64_t>(4, ceil_frac(q, depth - nbusy)));
        }
        int64_t nb = 0;
        const BatchReq* lead = b->queue.front();
        for (auto it = b->queue.begin(); it != b->queue.end();) {
          BatchReq* r = *it;
          if (r->kind != lead->kind || (r->kind == 1 && r->alive_mask != lead->alive_mask)) { ++it; continue; }
          const int64_t rb = req_blocks(r);
          if (!reqs.empty() && nb + rb > cap) break;
          reqs.push_back(r);
          it = b->queue.erase(it);
          nb += rb;
        }
      }
    }
    if (!reqs.empty()) {
      Slot& s = c->slots[head];
      int rc;
      const auto t_submit = clk::now();
      {
        std::lock_guard<std::mutex> lk(c->mu);
        inflight[head].is_get = reqs.front()->kind == 1;
        inflight[head].get_k = c->k;
        if (inflight[head].is_get) {
          rc = batcher_submit_get(b, s, reqs);
          if (rc != MEC_OK)
            for (BatchReq* r : reqs) r->ch.s = &s;  // batcher_finish looks the slot up through the first request
        } else {
          rc = batcher_submit(b, s, reqs, &inflight[head].needs_retire);
        }
      }
      b->us_submit += us_since(t_submit);
      inflight[head].reqs = reqs;
      inflight[head].busy = true;
      inflight[head].traced = b->trace && rc == MEC_OK && (inflight[head].is_get || !inflight[head].needs_retire);
      if (rc != MEC_OK) {  // drain what was enqueued before the failure, then fail the whole batch
        cudaStreamSynchronize(s.st);
        batcher_finish(b, inflight[head], rc);
      } else {
        head = (head + 1) % kSlots;
        nbusy++;
      }
      bool more;
      {
        std::lock_guard<std::mutex> lk(b->mu);
        more = !b->queue.empty();
      }
      if (more && nbusy < kSlots) continue;  // keep the slots full before waiting on the oldest batch
    }
    if (nbusy > 0) {
      const auto t_sync = clk::now();
      const cudaError_t e = cudaStreamSynchronize(c->slots[tail].st);
      b->us_sync += us_since(t_sync);
      if (inflight[tail].traced && e == cudaSuccess) {
        float a = 0, k2 = 0, sc = 0;
        cudaEventElapsedTime(&a, b->ev_ph[tail][0], b->ev_ph[tail][1]);
        cudaEventElapsedTime(&k2, b->ev_ph[tail][1], b->ev_ph[tail][2]);
        cudaEventElapsedTime(&sc, b->ev_ph[tail][2], b->ev_ph[tail][3]);
        b->us_stage += static_cast<int64_t>(a * 1000); b->us_kernel += static_cast<int64_t>(k2 * 1000); b->us_scatter += static_cast<int64_t>(sc * 1000);
      }
      const auto t_fin = clk::now();
      batcher_finish(b, inflight[tail], e == cudaSuccess ? MEC_OK : MEC_ERR_CUDA);
      b->us_finish += us_since(t_fin);
      tail = (tail + 1) % kSlots;
      nbusy--;
    }
  }
}
}  // namespace

extern "C" int mec_batcher_new(int k, int m, int64_t block_size, int device, int64_t max_batch_blocks, int max_wait_us, mec_batcher** out) {
  if (!out) return MEC_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  mec_codec* c = nullptr;
  int rc = mec_codec_new(k, m, block_size, MEC_HIGHWAYHASH256S, device, &c);
  if (rc) return rc;
  {
    std::lock_guard<std::mutex> lk(c->mu);
    rc = ensure_engine(c);
  }
  if (rc) { mec_codec_free(c); return rc; }
  if (c->S() >= (1ll << 31)) { mec_codec_free(c); return MEC_ERR_UNSUPPORTED; }
  mec_batcher* b = new mec_batcher;
  b->codec = c;
  if (max_batch_blocks > 0) b->max_blocks = max_batch_blocks;
  // a merged batch stages at most 1 GiB of object bytes: six slots are sized for a full batch up front (below)
  b->max_blocks = std::max<int64_t>(1, std::min<int64_t>(b->max_blocks, (1ll << 30) / std::max<int64_t>(block_size, 1)));
  if (max_wait_us >= 0) b->max_wait_us = max_wait_us;
  {
    // size every slot for a full batch up front: growing a buffer later means cudaFree + cudaMalloc, which stall the whole device
    // (all six batches in flight) for tens of milliseconds — seen as p99 spikes of 150-450 ms before
    std::lock_guard<std::mutex> lk(c->mu);
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(device);
    const int64_t S = c->S(), pitch = round_up(S, 16), P = round_up(32 + S, 16), slots = b->max_blocks + std::min<int64_t>(64, b->max_blocks / 8 + 4);
    for (Slot& s : c->slots) {
      if (rc == MEC_OK) rc = s.src.ensure(static_cast<size_t>(slots * block_size));
      if (rc == MEC_OK) rc = s.out.ensure(static_cast<size_t>(slots * std::max(m, 1) * pitch));
      if (rc == MEC_OK) rc = s.dig.ensure(static_cast<size_t>(slots * c->n * 32));
      if (rc == MEC_OK) rc = s.hdig.ensure(static_cast<size_t>(slots * c->n * 32));
      if (rc == MEC_OK) rc = s.aux.ensure(static_cast<size_t>(k * round_up(slots * P + 512, 256)));
      if (rc == MEC_OK) rc = s.flags.ensure(static_cast<size_t>(slots * k));
      if (rc == MEC_OK) rc = s.hflags.ensure(static_cast<size_t>(slots * k));
      if (rc == MEC_OK) rc = s.dreq.ensure(static_cast<size_t>(slots) * std::max(sizeof(BatchReqDesc), sizeof(BatchGetDesc)));
      if (rc == MEC_OK) rc = s.hreq.ensure(static_cast<size_t>(slots) * std::max(sizeof(BatchReqDesc), sizeof(BatchGetDesc)));
    }
    cudaSetDevice(cur);
  }
  if (rc) { mec_codec_free(c); delete b; return rc; }
  b->trace = getenv("MEC_BATCHER_TRACE") != nullptr;
  if (b->trace) {
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(device);
    for (int i = 0; i < kSlots; i++)
      for (int j = 0; j < 4; j++)
        if (cudaEventCreate(&b->ev_ph[i][j]) != cudaSuccess) { cudaGetLastError(); b->trace = false; }
    cudaSetDevice(cur);
  }
  if (const char* e = getenv("MEC_BATCHER_GATHER_PCT")) b->gather_pct = std::max(0, std::min(100, atoi(e)));
  if (b->gather_pct > 0) {
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(device);
    bool ok = cudaStreamCreateWithFlags(&b->aux, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < kSlots; i++)
      ok = cudaEventCreateWithFlags(&b->ev_table[i], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&b->ev_gather[i], cudaEventDisableTiming) == cudaSuccess;
    cudaSetDevice(cur);
    if (!ok) { cudaGetLastError(); b->gather_pct = 0; }
  }
  b->worker = std::thread(batcher_main, b);
  *out = b;
  return MEC_OK;
}

extern "C" void mec_batcher_free(mec_batcher* b) {
  if (!b) return;
  {
    std::lock_guard<std::mutex> lk(b->mu);
    b->stop = true;
  }
  b->cv_work.notify_all();
  if (b->worker.joinable()) b->worker.join();
  for (int i = 0; i < kSlots; i++) {
    if (b->ev_table[i]) cudaEventDestroy(b->ev_table[i]);
    if (b->ev_gather[i]) cudaEventDestroy(b->ev_gather[i]);
    for (int j = 0; j < 4; j++)
      if (b->ev_ph[i][j]) cudaEventDestroy(b->ev_ph[i][j]);
  }
  if (b->aux) cudaStreamDestroy(b->aux);
  mec_codec_free(b->serial);
  mec_codec_free(b->codec);
  delete b;
}

static int64_t batcher_call(mec_batcher* b, const uint8_t* src, int64_t len, uint8_t* const* files, uint8_t* data_digests, bool with_data,
                            int write_quorum) {
  if (!b || len < 0 || !files || (len > 0 && !src)) return MEC_ERR_INVALID_ARGUMENT;
  int online = 0;
  for (int i = 0; i < b->codec->n; i++) online += files[i] != nullptr;
  if (online < write_quorum) return MEC_ERR_WRITE_QUORUM;  // cmd/erasure-encode.go:59-65
  if (len == 0) return 0;
  BatchReq r;
  r.src = src; r.len = len; r.files = files; r.data_digests = data_digests; r.with_data = with_data;
  // page-locked buffers (mec_alloc_pinned*, or anything cudaHostRegister'ed) are moved by the gather / scatter kernels; checked here,
  // on the caller's thread, so the worker does not pay for it
  r.mapped = mec_is_pinned(src) && (!data_digests || mec_is_pinned(data_digests));
  for (int i = 0; i < b->codec->n && r.mapped; i++)
    if (files[i] && (i >= b->codec->k || with_data)) r.mapped = mec_is_pinned(files[i]) != 0;
  std::unique_lock<std::mutex> lk(b->mu);
  if (b->stop) return MEC_ERR_INVALID_ARGUMENT;
  b->queue.push_back(&r);
  if (b->queue.size() == 1 || static_cast<int64_t>(b->queue.size()) * std::max<int64_t>(1, len / b->codec->block_size) >= b->max_blocks)
    b->cv_work.notify_one();  // the worker is woken by the first arrival and by the one that fills the batch, not by every caller
  r.cv.wait(lk, [&] { return r.done; });
  return r.result;
}
extern "C" int64_t mec_batcher_encode(mec_batcher* b, const uint8_t* src, int64_t len, uint8_t* const* files, int write_quorum) {
  return batcher_call(b, src, len, files, nullptr, true, write_quorum);
}
extern "C" int64_t mec_batcher_encode_sg(mec_batcher* b, const uint8_t* src, int64_t len, uint8_t* const* files, uint8_t* data_digests,
                                         int write_quorum) {
  if (!data_digests) return MEC_ERR_INVALID_ARGUMENT;
  return batcher_call(b, src, len, files, data_digests, false, write_quorum);
}
// Erasure.Decode through the coalescer: GETs that are queued together and see the same drives online are merged into one launch.
// Calls whose buffers are not page-locked, and calls that meet bitrot, run on their own through a private handle.
extern "C" int64_t mec_batcher_decode(mec_batcher* b, const uint8_t* const* files, int64_t offset, int64_t length, int64_t total,
                                      uint8_t* dst, int* heal_hint) {
  if (heal_hint) *heal_hint = 0;
  if (!b || !files) return MEC_ERR_INVALID_ARGUMENT;
  if (offset < 0 || length < 0 || offset + length > total) return MEC_ERR_INVALID_ARGUMENT;  // cmd/erasure-decode.go:240-245
  if (length == 0) return 0;
  if (!dst) return MEC_ERR_INVALID_ARGUMENT;
  mec_codec* c = b->codec;
  const int n = c->n, k = c->k;
  BatchReq r;
  r.kind = 1; r.gfiles = files; r.offset = offset; r.length = length; r.total = total; r.dst = dst; r.heal_hint = heal_hint;
  bool mapped = n <= kBatchMaxFiles && n <= 64 && k <= 32 && mec_is_pinned(dst) != 0;
  int online = 0;
  for (int i = 0; i < n; i++) {
    if (!files[i]) continue;
    online++;
    if (i < 64) r.alive_mask |= 1ull << i;
  }
  if (online < k) return MEC_ERR_READ_QUORUM;
  {  // only the k readers that will be used have to be mapped
    int cnt = 0;
    for (int i = 0; i < n && cnt < k && mapped; i++)
      if (files[i]) { mapped = mec_is_pinned(files[i]) != 0; cnt++; }
  }
  if (!mapped || getenv("MEC_BATCHER_MEMCPY") != nullptr) {
    mec_codec* sc = batcher_serial(b);
    return sc ? mec_decode_prefer(sc, files, nullptr, offset, length, total, dst, heal_hint) : MEC_ERR_UNEXPECTED;
  }
  std::unique_lock<std::mutex> lk(b->mu);
  if (b->stop) return MEC_ERR_INVALID_ARGUMENT;
  b->queue.push_back(&r);
  b->cv_work.notify_one();
  r.cv.wait(lk, [&] { return r.done; });
  return r.result;
}

extern "C" int64_t mec_batcher_stat(const mec_batcher* b, const char* name) {
  if (!b || !name) return -1;
  if (!strcmp(name, "batches")) return b->st_batches.load();
  if (!strcmp(name, "requests")) return b->st_requests.load();
  if (!strcmp(name, "blocks")) return b->st_blocks.load();
  if (!strcmp(name, "kernel_batches")) return b->st_kernel_batches.load();
  if (!strcmp(name, "launches")) return b->codec->eng ? b->codec->eng->launches() : 0;
  if (!strcmp(name, "us_submit")) return b->us_submit.load();
  if (!strcmp(name, "us_sync")) return b->us_sync.load();
  if (!strcmp(name, "us_finish")) return b->us_finish.load();
  if (!strcmp(name, "us_idle")) return b->us_idle.load();
  if (!strcmp(name, "us_stage")) return b->us_stage.load();
  if (!strcmp(name, "us_kernel")) return b->us_kernel.load();
  if (!strcmp(name, "us_scatter")) return b->us_scatter.load();
  return -1;
}
