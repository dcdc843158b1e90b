/*
 * minio_ec.h — C ABI of the B200 erasure-code + bitrot library (libminio_ec.so).
 *
 * Drop-in boundary for MinIO's erasure hot path.  The reference has no FFI here — the path sits
 * behind Go interfaces (reedsolomon.Encoder held by `Erasure.encoder`, hash.Hash from
 * BitrotAlgorithm.New, io.Writer / io.ReaderAt per drive).  Each entry point names the reference
 * function whose body a cgo binding would replace (file:line under /root/reference); the Go-side
 * stub is written out in INTEGRATION.md.  Plain pointers and sizes only; no retained caller memory.
 *
 * Return codes: 0 / non-negative = success, negative = one of MEC_ERR_*.
 */
#ifndef MINIO_EC_H
#define MINIO_EC_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- errors: reedsolomon.Err* (go.mod:49) and MinIO storage errors ------------------------- */
enum {
  MEC_OK = 0,
  MEC_ERR_INV_SHARD_NUM = -1,     /* reedsolomon.ErrInvShardNum   cmd/erasure-coding.go:45 */
  MEC_ERR_MAX_SHARD_NUM = -2,     /* reedsolomon.ErrMaxShardNum   cmd/erasure-coding.go:49 */
  MEC_ERR_TOO_FEW_SHARDS = -3,    /* reedsolomon.ErrTooFewShards */
  MEC_ERR_SHARD_NO_DATA = -4,     /* reedsolomon.ErrShardNoData */
  MEC_ERR_SHARD_SIZE = -5,        /* reedsolomon.ErrShardSize */
  MEC_ERR_SHORT_DATA = -6,        /* reedsolomon.ErrShortData */
  MEC_ERR_FILE_CORRUPT = -7,      /* errFileCorrupt        cmd/storage-errors.go:104 */
  MEC_ERR_LESS_DATA = -8,         /* errLessData           cmd/storage-errors.go:114 */
  MEC_ERR_UNEXPECTED = -9,        /* errUnexpected         cmd/storage-errors.go:29 */
  MEC_ERR_READ_QUORUM = -10,      /* errErasureReadQuorum  cmd/erasure-errors.go:23 */
  MEC_ERR_WRITE_QUORUM = -11,     /* errErasureWriteQuorum cmd/erasure-errors.go:26 */
  MEC_ERR_INVALID_ARGUMENT = -12, /* errInvalidArgument */
  MEC_ERR_CUDA = -100,            /* CUDA runtime / driver failure (see mec_last_error) */
  MEC_ERR_NO_DEVICE = -101,       /* no CUDA device: the library never falls back to the CPU */
  MEC_ERR_UNSUPPORTED = -102      /* geometry or algorithm outside the GPU path */
};

/* ---- bitrot algorithms: cmd/xl-storage-format-v1.go:142-159 -------------------------------- */
enum { MEC_SHA256 = 1, MEC_HIGHWAYHASH256 = 2, MEC_HIGHWAYHASH256S = 3, MEC_BLAKE2B512 = 4 };

typedef struct mec_codec mec_codec;

/* ---- lifecycle ----------------------------------------------------------------------------- */
/* NewErasure (cmd/erasure-coding.go:42): validates k>0, m>=0, k+m<=256 with the same errors.
 * `device` is the CUDA ordinal.  One codec may be shared by threads; calls serialise per codec. */
int mec_codec_new(int k, int m, int64_t block_size, int bitrot_algo, int device, mec_codec** out);
void mec_codec_free(mec_codec* c);
int mec_device_count(void);
const char* mec_last_error(void); /* thread-local description of the last MEC_ERR_CUDA */
const char* mec_version(void);

/* ---- size helpers -------------------------------------------------------------------------- */
int64_t mec_shard_size(const mec_codec* c);                                  /* Erasure.ShardSize       cmd/erasure-coding.go:116 */
int64_t mec_shard_file_size(const mec_codec* c, int64_t total_length);       /* Erasure.ShardFileSize   cmd/erasure-coding.go:121 */
int64_t mec_shard_file_offset(const mec_codec* c, int64_t start_offset, int64_t length,
                              int64_t total_length);                         /* Erasure.ShardFileOffset cmd/erasure-coding.go:135 */
int64_t mec_bitrot_shard_file_size(int64_t size, int64_t shard_size, int algo); /* bitrotShardFileSize cmd/bitrot.go:156 */
int64_t mec_ceil_frac(int64_t numerator, int64_t denominator);               /* ceilFrac cmd/utils.go:689 */

/* ---- pinned host memory for the byte pool (internal/bpool/bpool.go:25-93; AllocAligned :52,68) ------
 * mec_alloc_pinned: page-locked memory, wherever the calling thread runs.  mec_alloc_pinned_on: page-locked memory on the
 * NUMA node of CUDA device `device` (/sys/bus/pci/devices/<bdf>/numa_node; mbind + first touch + cudaHostRegister) — on a
 * multi-socket box every GPU should stream from the memory behind its own root port.  Both are released with
 * mec_free_pinned.  mec_device_numa_node returns that node (-1 unknown); mec_bind_thread_to_device restricts the calling
 * thread to the CPUs of that node (never beyond its current affinity mask) and returns the node, or -1 when it could not.
 * mec_is_pinned: 1 when `p` is page-locked memory known to CUDA (the host-buffer entry points DMA straight from / into it). */
void* mec_alloc_pinned(size_t bytes);
void* mec_alloc_pinned_on(int device, size_t bytes);
void mec_free_pinned(void* p);
int mec_device_numa_node(int device);
int mec_bind_thread_to_device(int device);
int mec_is_pinned(const void* p);

/* ---- fused block encode: the body of Erasure.EncodeData (cmd/erasure-coding.go:77-91) plus
 * the HighwayHash256 of every shard that streamingBitrotWriter.Write computes
 * (cmd/bitrot-streaming.go:57-59), for ALL erasure blocks of `src` in one call.
 *   src      : len object bytes (host).  Blocks are block_size bytes; the last may be short.
 *   parity   : receives parity shard j of block b at parity + (b*m + j)*shard_size  (host)
 *   digests  : receives 32-byte digest of shard i of block b at digests + (b*(k+m)+i)*32 (host)
 * Data shards alias `src` exactly as Split does (shard i of block b = src[b*block_size + i*S_b ..),
 * zero padded), so they are not copied back.  len == 0 is a no-op (EncodeData returns nil shards). */
int mec_encode_blocks(mec_codec* c, const uint8_t* src, int64_t len, uint8_t* parity, uint8_t* digests);

/* Same, device-resident: d_src/d_parity/d_digests are device pointers on the codec's device,
 * d_src 16-byte aligned, parity row pitch `parity_pitch` (multiple of 16, >= shard_size):
 * parity shard (b, j) at d_parity + (b*m + j)*parity_pitch.  Asynchronous on `cuda_stream`
 * (a cudaStream_t, may be NULL).  d_digests == NULL computes parity only (whole-file bitrot algorithms). */
int mec_encode_blocks_device(mec_codec* c, const uint8_t* d_src, int64_t len, uint8_t* d_parity,
                             int64_t parity_pitch, uint8_t* d_digests, void* cuda_stream);

/* ---- fused block reconstruct: Erasure.DecodeDataBlocks / DecodeDataAndParityBlocks
 * (cmd/erasure-coding.go:94-113 -> reedsolomon ReconstructData / Reconstruct) fused with the
 * digest check of streamingBitrotReader.ReadAt (cmd/bitrot-streaming.go:186-196).
 *   frames[i]   : host pointer to shard file i in streaming-bitrot layout ([32B digest][shard])*,
 *                 or NULL when the drive is offline.  All `nblocks` blocks are full blocks except
 *                 that the last block's shard length is `last_shard_len` (0 => shard_size).
 *   want[i]     : 1 = rebuild shard i into out[i] (same frame layout, digests recomputed)
 *   data_only   : ReconstructData semantics (parity never rebuilt)
 *   corrupt[i]  : set to 1 when any frame of file i fails its digest
 * The first k readable files in index order are used (parallelReader.Read, cmd/erasure-decode.go:127);
 * a file with a corrupt frame is dropped for the whole call and the next one is tried. */
int mec_reconstruct_frames(mec_codec* c, const uint8_t* const* frames, int64_t nblocks,
                           int64_t last_shard_len, const uint8_t* want, int data_only,
                           uint8_t* const* out, uint8_t* corrupt);

/* Device-resident variant of the fused reconstruct (BASELINE configs 3 and 4): d_frames[i] is a device
 * pointer to shard file i in frame layout with a 16-byte aligned `frame_pitch` (>= 32 + shard_size;
 * digest at +0, shard bytes at +32 of every frame), or NULL when the shard is unavailable.  The first k
 * non-NULL files are read; every shard with want[i] != 0 that was not read is rebuilt into
 * d_out + (b*r + q)*out_pitch (q = rank of i among the rebuilt shards) and its digest written to
 * d_digests[(b*(k+r) + k + q)*32]; digests of the k shards read land in d_digests[(b*(k+r) + t)*32] and
 * d_corrupt[b*k + t] is set when frame t of block b fails its stored digest.  All nblocks are full blocks.
 * `flags`: MEC_RECONSTRUCT_DATA_ONLY = ReconstructData semantics (parity never rebuilt);
 * MEC_RECONSTRUCT_NO_OUTPUT_DIGESTS = the rebuilt shards are not hashed (GetObject consumes their bytes only,
 * cmd/erasure-decode.go:283-289; the k shards read are still hashed and checked), their digest slots stay untouched.
 * d_corrupt != NULL requires d_digests != NULL (the digest check runs in the hash threads; MEC_ERR_INVALID_ARGUMENT otherwise);
 * with d_digests == NULL nothing is hashed or verified.
 * Asynchronous on `cuda_stream`; the fail-over policy stays with the caller (see mec_reconstruct_frames). */
#define MEC_RECONSTRUCT_DATA_ONLY 1
#define MEC_RECONSTRUCT_NO_OUTPUT_DIGESTS 2
int mec_reconstruct_device(mec_codec* c, const uint8_t* const* d_frames, int64_t frame_pitch, int64_t nblocks,
                           const uint8_t* want, int flags, uint8_t* d_out, int64_t out_pitch,
                           uint8_t* d_digests, uint8_t* d_corrupt, void* cuda_stream);

/* ---- whole-part drivers ------------------------------------------------------------------- */
/* Erasure.Encode (cmd/erasure-encode.go:69) with one streaming bitrot writer per shard
 * (cmd/bitrot.go:105, cmd/bitrot-streaming.go:44): files[i] (host, or NULL = offline writer)
 * receives exactly the bytes of part.N on drive i; each must hold
 * mec_bitrot_shard_file_size(mec_shard_file_size(len)) bytes.  Fails with MEC_ERR_WRITE_QUORUM
 * when fewer than `write_quorum` writers are online.  Returns bytes consumed. */
int64_t mec_encode(mec_codec* c, const uint8_t* src, int64_t len, uint8_t* const* files, int write_quorum);
/* The same call in the shape streamingBitrotWriter.Write actually has (cmd/bitrot-streaming.go:57-69: w.Write(hash), then
 * w.Write(p) — the shard bytes are never copied next to their digest): the k data shards stay inside `src`, where Split
 * aliased them (cmd/erasure-coding.go:81), and only what the GPU produced comes back —
 *   files[k..n)   : complete part.N images of the parity drives (NULL = offline), written by DMA straight from device memory
 *   data_digests  : digest of data shard i of block b at data_digests + (b*k + i)*32; writer i emits, per block,
 *                   {data_digests[b][i], src + b*block_size + i*S_b (zero padded to S_b)}  (writev / the cgo writeHashed stub)
 *   files[0..k)   : only tested for NULL (offline writers count against write_quorum)
 * No assembly copy exists anywhere on this path.  Returns bytes consumed. */
int64_t mec_encode_sg(mec_codec* c, const uint8_t* src, int64_t len, uint8_t* const* files, uint8_t* data_digests,
                      int write_quorum);

/* Erasure.Decode (cmd/erasure-decode.go:239) over streaming bitrot readers: writes object bytes
 * [offset, offset+length) of a part of `total_length` bytes to dst.  files[i] NULL = offline.
 * Returns bytes written; *heal_hint is set to MEC_ERR_FILE_CORRUPT when a frame failed its digest
 * but the read still succeeded (the errFileCorrupt side-band of cmd/erasure-decode.go:288-293). */
int64_t mec_decode(mec_codec* c, const uint8_t* const* files, int64_t offset, int64_t length,
                   int64_t total_length, uint8_t* dst, int* heal_hint);

/* Same with parallelReader.preferReaders (cmd/erasure-decode.go:92-123; `prefer` as passed by
 * cmd/erasure-object.go:387 for local drives): readers with prefer[i] != 0 are tried before the others, each
 * class in index order.  prefer == NULL is mec_decode. */
int64_t mec_decode_prefer(mec_codec* c, const uint8_t* const* files, const uint8_t* prefer, int64_t offset,
                          int64_t length, int64_t total_length, uint8_t* dst, int* heal_hint);

/* Erasure.Heal (cmd/erasure-decode.go:317): rebuilds every shard file with out_files[i] != NULL
 * from the readable files (NULL = offline / stale). */
int mec_heal(mec_codec* c, const uint8_t* const* files, int64_t total_length, uint8_t* const* out_files);
/* Heal with its `prefer` argument and the bitrot side-band.  Returns MEC_OK, or MEC_ERR_FILE_CORRUPT when the stale shards
 * WERE rebuilt and written but a source reader failed its digest on the way — Heal's derr (cmd/erasure-decode.go:338-341,366;
 * healObject aborts the part on it, cmd/erasure-healing.go:603-608); corrupt[i] (optional, n bytes) names those readers.
 * mec_heal is this call with prefer = NULL, corrupt = NULL and returns the same codes. */
int mec_heal_prefer(mec_codec* c, const uint8_t* const* files, const uint8_t* prefer, int64_t total_length,
                    uint8_t* const* out_files, uint8_t* corrupt);
/* Heal of many objects (BASELINE config 4; healObject fan-out of cmd/global-heal.go:152): object o has files[o][0..n),
 * totals[o], out_files[o][0..n) with mec_heal's meaning.  `pool` holds npool codec handles of the same geometry (each owns
 * its streams and staging buffers); one host thread per handle pulls objects, so staging, kernels and copy-back of
 * different objects overlap.  rcs (optional) receives the per-object result (MEC_ERR_FILE_CORRUPT = healed, a source reader
 * failed its digest); the return value is the first result that is neither MEC_OK nor MEC_ERR_FILE_CORRUPT. */
int mec_heal_batch(mec_codec* const* pool, int npool, int64_t nobjects, const uint8_t* const* const* files,
                   const int64_t* total_lengths, uint8_t* const* const* out_files, int* rcs);

/* ---- cross-request coalescer: many concurrent PutObject calls, one launch ---------------------------------------------------
 * MinIO encodes one block per loop iteration on every request goroutine (cmd/erasure-encode.go:76-108, one Erasure per request,
 * cmd/erasure-object.go:1371); a launch that small leaves 147 of 148 SMs idle.  A batcher owns one codec (k, m, block_size,
 * HighwayHash256S) and a worker thread: callers block in mec_batcher_encode / mec_batcher_encode_sg (same arguments and results
 * as mec_encode / mec_encode_sg), the worker merges everything that is queued — up to max_batch_blocks erasure blocks, waiting
 * at most max_wait_us for company when the GPU is idle — into one staged buffer (one batched copy call) and ONE fused launch over
 * all full blocks, then writes every caller's frames straight into that caller's buffers.  Up to six merged batches are in flight;
 * device staging for all six is allocated here (about 2.4 x max_batch_blocks x block_size each).
 * mec_batcher_stat: "batches", "requests", "blocks", "launches"; with MEC_BATCHER_TRACE=1 in the environment also the worker's and the
 * streams' accumulated phase times "us_submit", "us_sync", "us_finish", "us_idle", "us_stage", "us_kernel", "us_scatter". */
typedef struct mec_batcher mec_batcher;
int mec_batcher_new(int k, int m, int64_t block_size, int device, int64_t max_batch_blocks, int max_wait_us, mec_batcher** out);
void mec_batcher_free(mec_batcher* b);
int64_t mec_batcher_encode(mec_batcher* b, const uint8_t* src, int64_t len, uint8_t* const* files, int write_quorum);
int64_t mec_batcher_encode_sg(mec_batcher* b, const uint8_t* src, int64_t len, uint8_t* const* files, uint8_t* data_digests,
                              int write_quorum);
/* Erasure.Decode through the coalescer (arguments and results of mec_decode): GETs queued together that see the same drives online
 * share one reader set and are merged into one fused launch (gather kernel, reconstruct, scatter into every caller's dst).  A request
 * that meets a corrupt frame is redone on its own with parallelReader's exact fail-over; pageable buffers take the same private path. */
int64_t mec_batcher_decode(mec_batcher* b, const uint8_t* const* files, int64_t offset, int64_t length, int64_t total_length,
                           uint8_t* dst, int* heal_hint);
int64_t mec_batcher_stat(const mec_batcher* b, const char* name);

/* ---- object checksums of the PutObject stream (internal/hash/checksum.go:64-73, internal/hash/crc.go) ---------------------
 * CRC32 (IEEE), CRC32C (Castagnoli) and CRC64NVME — the checksum types hash.Reader can merge (ChecksumType.CanMerge) — of a
 * byte stream, as the finalized values Go's hash/crc32 and hash/crc64 return (big-endian encode them for Checksum.Raw).
 * `which` is a mask of MEC_CRC*; out3[0] = CRC32, out3[1] = CRC32C, out3[2] = CRC64NVME (0 where not asked for).
 *   mec_checksums_device : the stream is already in device memory (e.g. staged for an encode)
 *   mec_checksums        : host buffer, staged through the codec's slots
 *   mec_checksum_combine : Checksum.AddPart (crc.go:32-73): checksum of A || B from those of A and B and |B|, on the host
 * mec_set_option(c, "checksums", mask) makes mec_encode / mec_encode_sg / mec_encode_blocks compute them on the bytes they stage
 * anyway (the object crosses PCIe once); mec_last_checksums returns the values of the codec's last such call and its length. */
#define MEC_CRC32 1
#define MEC_CRC32C 2
#define MEC_CRC64NVME 4
int mec_checksums_device(mec_codec* c, const uint8_t* d_src, int64_t len, int which, uint64_t* out3, void* cuda_stream);
int mec_checksums(mec_codec* c, const uint8_t* src, int64_t len, int which, uint64_t* out3);
uint64_t mec_checksum_combine(int type, uint64_t crc1, uint64_t crc2, int64_t len2);
int64_t mec_last_checksums(const mec_codec* c, uint64_t* out3);

/* ---- legacy whole-file bitrot (cmd/bitrot-whole.go, BitrotAlgorithm SHA256 / BLAKE2b512 / HighwayHash256) ----
 * Erasure.Encode with wholeBitrotWriters (cmd/bitrot-whole.go:35-45): files[i] receives the raw shard file
 * (shards of all blocks back to back, mec_shard_file_size(len) bytes) and sums + i*64 the digest over the whole
 * file that bitrotWriterSum (cmd/bitrot.go:148) would return (32 or 64 bytes).  The codec's algorithm must be
 * one of the whole-file algorithms.  Returns bytes consumed. */
int64_t mec_encode_whole(mec_codec* c, const uint8_t* src, int64_t len, uint8_t* const* files, uint8_t* sums,
                         int write_quorum);
/* BitrotAlgorithm.New() hash.Hash one-shot (cmd/bitrot.go:47-64) for a whole-file algorithm: digests of
 * `count` equal-length messages laid out back to back; digest i at digests + i*mec_digest_size(algo).
 * Also bitrotVerify's non-streaming branch (cmd/bitrot.go:165-175): hash the file, compare with `want`. */
int mec_whole_hash(mec_codec* c, int algo, const uint8_t* msgs, int64_t msg_len, int64_t count, uint8_t* digests);
int mec_bitrot_verify_whole(mec_codec* c, int algo, const uint8_t* file, int64_t file_len, const uint8_t* want);
/* Device-resident mec_whole_hash: message i at d_msgs + i*pitch (pitch a multiple of 16, >= msg_len, with 128 readable
 * bytes after the last message), digest i at d_digests + i*64.  Asynchronous on `cuda_stream`. */
int mec_whole_hash_device(mec_codec* c, int algo, const uint8_t* d_msgs, int64_t pitch, int64_t msg_len, int64_t count,
                          uint8_t* d_digests, void* cuda_stream);
/* Erasure.Decode / Erasure.Heal over whole-file bitrot readers (wholeBitrotReader.ReadAt, cmd/bitrot-whole.go:66-81): files[i] is
 * the RAW shard file of drive i (mec_shard_file_size(total) bytes, NULL = offline), sums + i*64 its expected digest (the
 * checksum xl.meta keeps for whole-file algorithms).  A reader whose whole-file digest does not match is errFileCorrupt: it is
 * dropped and the next drive in index order takes its place (cmd/xl-storage.go:1931-1950, cmd/erasure-decode.go:196-199).
 * mec_decode_whole returns bytes written (*heal_hint = MEC_ERR_FILE_CORRUPT when a reader was dropped but the read succeeded).
 * mec_heal_whole rebuilds every shard file with out_files[i] != NULL, writes its digest to out_sums + i*64 (optional) and returns
 * MEC_OK, or MEC_ERR_FILE_CORRUPT when it healed but met bitrot in a source (corrupt[i], optional, names the readers).
 * The codec's algorithm must be SHA256, BLAKE2b512 or HighwayHash256. */
int64_t mec_decode_whole(mec_codec* c, const uint8_t* const* files, const uint8_t* sums, int64_t offset, int64_t length,
                         int64_t total_length, uint8_t* dst, int* heal_hint);
int mec_heal_whole(mec_codec* c, const uint8_t* const* files, const uint8_t* sums, int64_t total_length,
                   uint8_t* const* out_files, uint8_t* out_sums, uint8_t* corrupt);
int mec_digest_size(int algo);

/* bitrotVerify (cmd/bitrot.go:164) for the streaming algorithm: scans a whole shard file. */
int mec_bitrot_verify(mec_codec* c, const uint8_t* file, int64_t file_len, int64_t part_len);
/* The same for many shard files at once (the deep scan of the data scanner visits every part on every drive): results[f] = MEC_OK or
 * MEC_ERR_FILE_CORRUPT (wrong length or a frame that fails its digest); the frames of all files stream through the codec's slots and
 * consecutive full frames — across files — share one hash-only launch.  Returns MEC_OK unless the call itself failed. */
int mec_bitrot_verify_batch(mec_codec* c, int64_t nfiles, const uint8_t* const* files, const int64_t* file_lens,
                            const int64_t* part_lens, int* results);

/* ---- shard-shaped low-level calls (keep Erasure.EncodeData / reedsolomon.Encoder shapes) ---- */
/* reedsolomon.Encoder.Encode: shards[0..k) in, shards[k..k+m) out, each shard_len bytes (host). */
int mec_rs_encode_shards(mec_codec* c, uint8_t* const* shards, int64_t shard_len);
/* reedsolomon.Encoder.Reconstruct / ReconstructData: shards[i] with present[i]==0 are rebuilt in
 * place (buffers must exist). */
int mec_rs_reconstruct_shards(mec_codec* c, uint8_t* const* shards, const uint8_t* present,
                              int64_t shard_len, int data_only);
/* hash.Hash one-shot for BitrotAlgorithm.New() == HighwayHash256(S) (cmd/bitrot.go:55-58): digests
 * of `count` equal-length messages laid out back to back (host). */
int mec_hh256_batch(mec_codec* c, const uint8_t* msgs, int64_t msg_len, int64_t count, uint8_t* digests);

/* erasureSelfTest + bitrotSelfTest (cmd/erasure-coding.go:149, cmd/bitrot.go:224) on the GPU. */
int mec_selftest(int device);

/* tuning knobs for benchmarks/tests (erasure blocks per CTA, loader, GF specialisation).
 * "jit": run-time specialisation of decode matrices / uncompiled (k, m) geometries with NVRTC, cached per process:
 *   -1 (default) a pattern that has been seen with 32 MiB of input is compiled on a background thread; calls that
 *      arrive before it is ready run the generic runtime-matrix kernel — no request ever waits for the compiler;
 *    1 compile (or wait for the background compile) inside the call — tests and benchmarks;  0 never specialise.
 * Compiled kernels are also kept on disk (MEC_JIT_CACHE_DIR, default $HOME/.cache/minio_b200, "off" disables), keyed by the kernel
 * sources, options and matrix: a restarted process loads the kernel of a pattern it has met before instead of recompiling.
 * "small_blocks": launches of at most this many erasure blocks run the latency form of the kernel (one CTA per block, hash warps
 *   decoupled from the GF warps: ~0.17 ms for a 1 MiB block instead of ~0.54 ms); -1 (default) = five per SM, 0 = never. */
int mec_set_option(mec_codec* c, const char* name, int64_t value);
/* Boundary counters (SURVEY §5 metrics row): name is one of "launches", "blocks_encoded", "blocks_read",
 * "shards_rebuilt", "corrupt_shards", "bytes_h2d", "bytes_d2h", "jit_compiles", "jit_ms", "jit_disk_hits" (process-wide), "jit_launches",
 * "small_launches" (launches that took the latency form).  -1 for unknown names.
 * Every ABI call is also wrapped in an NVTX range (visible in Nsight Systems) — the tracing hook of SURVEY §5. */
int64_t mec_get_stat(const mec_codec* c, const char* name);
/* Stops background kernel specialisation and waits for a compile in flight.  Call before the process exits (exit()
 * during an NVRTC compile lets libnvrtc's exit handlers run under the compiling thread); later calls of the library
 * keep working with the kernels already in the cache.  Also registered with atexit() after the first background compile. */
void mec_shutdown(void);
/* Queues, on the background compiler, the specialised kernels of every single-missing-data-shard pattern of the codec's geometry
 * (what a drive failure turns the next GETs and the heal into).  Returns the number of kernels queued; nothing blocks.  With the
 * on-disk cache this is a one-time cost per (k, m) and library version. */
int mec_jit_prewarm(mec_codec* c);
/* Build check of the run-time specialisation (no device needed): NVRTC-compiles the fused kernel for the r x k matrix `coef`
 * (align = S mod 16 of the inputs, eb = erasure blocks per CTA or 0, rows3d / hash_outputs as the engine would pass them)
 * without loading it.  Returns the cubin size, 0 when the compile fails (log in mec_last_error), -1 without libnvrtc. */
int64_t mec_jit_compile_check(int k, int r, const uint8_t* coef, int align, int eb, int rows3d, int hash_outputs);
/* number of kernels launched by this codec so far (bench.py's gpu_launches) */
int64_t mec_launch_count(const mec_codec* c);

#ifdef __cplusplus
}
#endif
#endif
