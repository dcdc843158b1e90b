// batch_copy.cuh — gather / scatter kernels of the cross-request coalescer (mec_batcher).
// A merged batch of hundreds of small PutObject calls cannot afford a handful of cudaMemcpy* calls per request: at ~4 us of
// driver time each, the worker thread tops out near 25 k requests/s.  Page-locked host memory is mapped into the device's
// address space (UVA), so two kernels move everything instead, driven by one descriptor table per batch:
//   gather : object bytes of every request, host -> the merged staging buffer (full blocks contiguous, tails behind them)
//   scatter: digest + shard of every (request, drive, block) straight into the caller's part-file images — the frame
//            assembly of streamingBitrotWriter.Write (cmd/bitrot-streaming.go:57-69), Split's zero padding included —
//            or, for the scatter-gather form, the data-shard digests into the caller's digest array.
// Host frames start at arbitrary byte offsets (a frame is 32 + S bytes); stores are issued 16-byte aligned on the DESTINATION
// (whole 128-byte PCIe writes per warp) and the source is re-aligned in registers.
#pragma once
#include <cstdint>

namespace mec {

constexpr int kBatchMaxFiles = 32;   // drives per erasure set the batcher accepts (MinIO sets have <= 16)

struct BatchReqDesc {
  const uint8_t* src;                // host (mapped): the caller's object bytes
  int64_t len;
  uint8_t* files[kBatchMaxFiles];    // host (mapped): part-file images, nullptr = offline / not wanted
  uint8_t* data_digests;             // host (mapped) or nullptr: [block][k][32]
  int64_t full0;                     // merged layout: slot of the request's first full block
  int64_t tail_slot;                 // slot of its tail block (parity rows / digests), -1 = none
  int64_t tail_src_off;              // byte offset of the tail block in the staging buffer
  int32_t with_data;                 // 1: data drives get complete frames too
  int32_t pad;
};

struct BatchCopyParams {
  const BatchReqDesc* reqs;          // host (mapped) table
  int nreq, k, m;
  int64_t bs, S, pitch;
  uint8_t* staged;                   // device: merged object bytes
  const uint8_t* parity;             // device: parity rows, slot s row j at (s*m + j)*pitch
  const uint8_t* digests;            // device: [slot][k+m][32]
};

// 16 bytes from an arbitrarily aligned source (4-byte aligned loads + funnel shifts)
__device__ __forceinline__ uint4 load16_any(const uint8_t* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  if ((a & 15) == 0) return *reinterpret_cast<const uint4*>(p);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
  const uint32_t sh = static_cast<uint32_t>(a & 3) * 8;
  const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
  if (sh == 0) return make_uint4(w0, w1, w2, w3);
  const uint32_t w4 = w[4];
  return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
}

// copy n bytes; dst may be misaligned: byte head up to the first 16-byte boundary of dst, 16-byte body, byte tail.
// The source must be readable up to 4 bytes past its end rounded up to a word (device staging buffers are padded; host
// sources are read exactly, see `exact_src`).
__device__ __forceinline__ void cta_copy(uint8_t* dst, const uint8_t* src, int64_t n, bool exact_src) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int64_t head = n < 16 ? n : ((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
  for (int64_t i = tid; i < head; i += nt) dst[i] = src[i];
  int64_t body = (n - head) / 16;
  if (exact_src && body > 0 && (reinterpret_cast<uintptr_t>(src + head) & 3) != 0) body--;  // the last funnel would read 1-3 bytes past the source
  for (int64_t q = tid; q < body; q += nt) *reinterpret_cast<uint4*>(dst + head + q * 16) = load16_any(src + head + q * 16);
  for (int64_t i = head + body * 16 + tid; i < n; i += nt) dst[i] = src[i];
}
// Host -> device copy for sources in mapped host memory.  SM loads over PCIe are worth their latency only as whole 128-byte
// requests, and host frames start at arbitrary (even) offsets: the source is therefore read in 16-byte vectors aligned on the
// SOURCE (every warp load = 512 contiguous bytes) into a shared-memory stage, and the stage is written out with the destination-
// aligned copy above.  Vectors stay inside the 16-byte units that hold the first and the last source byte (same pages).
constexpr int kStageBytes = 8192;
__device__ __forceinline__ void cta_copy_from_host(uint8_t* dst, const uint8_t* src, int64_t n, uint8_t* stage /* kStageBytes + 32, 16-byte aligned */) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int64_t o = 0; o < n; o += kStageBytes) {
    const int64_t len = n - o < kStageBytes ? n - o : kStageBytes;
    const uintptr_t a = reinterpret_cast<uintptr_t>(src + o);
    const uint4* base = reinterpret_cast<const uint4*>(a & ~static_cast<uintptr_t>(15));
    const int lead = static_cast<int>(a & 15);
    const int nvec = static_cast<int>((lead + len + 15) >> 4);
    __syncthreads();  // the previous stage has been consumed
    for (int v = tid; v < nvec; v += nt) reinterpret_cast<uint4*>(stage)[v] = base[v];
    __syncthreads();
    cta_copy(dst + o, stage + lead, len, true);
  }
}

__device__ __forceinline__ void cta_zero(uint8_t* dst, int64_t n) {
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = 0;
}

// grid (pieces, nreq): object bytes -> staging buffer
__global__ void __launch_bounds__(256) batch_gather_kernel(const BatchCopyParams p) {
  const BatchReqDesc& r = p.reqs[blockIdx.y];
  const int64_t full_bytes = r.len / p.bs * p.bs, tail = r.len - full_bytes;
  constexpr int64_t kPiece = 32 << 10;
  __shared__ __align__(16) uint8_t stage[kStageBytes + 32];
  for (int64_t o = static_cast<int64_t>(blockIdx.x) * kPiece; o < full_bytes; o += static_cast<int64_t>(gridDim.x) * kPiece)
    cta_copy_from_host(p.staged + r.full0 * p.bs + o, r.src + o, full_bytes - o < kPiece ? full_bytes - o : kPiece, stage);
  for (int64_t o = static_cast<int64_t>(blockIdx.x) * kPiece; o < tail; o += static_cast<int64_t>(gridDim.x) * kPiece)
    cta_copy_from_host(p.staged + r.tail_src_off + o, r.src + full_bytes + o, tail - o < kPiece ? tail - o : kPiece, stage);
}

// grid (block stride, k + m, nreq): frames (or data digests) of one drive of one request
__global__ void __launch_bounds__(128) batch_scatter_kernel(const BatchCopyParams p) {
  const BatchReqDesc& r = p.reqs[blockIdx.z];
  const int i = blockIdx.y, k = p.k, m = p.m, n = k + m;
  const int64_t nfull = r.len / p.bs, tail = r.len % p.bs, nb = nfull + (tail ? 1 : 0);
  const int64_t fstride = 32 + p.S;
  uint8_t* f = r.files[i];
  const bool frames = f != nullptr && (i >= k || r.with_data);
  const bool ddig = i < k && r.data_digests != nullptr;
  if (!frames && !ddig) return;
  for (int64_t b = blockIdx.x; b < nb; b += gridDim.x) {
    const bool is_tail = b >= nfull;
    const int64_t slot = is_tail ? r.tail_slot : r.full0 + b;
    const uint8_t* dg = p.digests + (slot * n + i) * 32;
    if (ddig) cta_copy(r.data_digests + (b * k + i) * 32, dg, 32, false);
    if (!frames) continue;
    const int64_t per = is_tail ? (tail + k - 1) / k : p.S, blen = is_tail ? tail : p.bs;
    uint8_t* fr = f + b * fstride;
    cta_copy(fr, dg, 32, false);  // hash first, then the shard (cmd/bitrot-streaming.go:60,65)
    if (i >= k) {
      cta_copy(fr + 32, p.parity + (slot * m + (i - k)) * p.pitch, per, false);
    } else {
      int64_t have = blen - static_cast<int64_t>(i) * per;
      have = have < 0 ? 0 : (have > per ? per : have);
      const uint8_t* sb = is_tail ? p.staged + r.tail_src_off : p.staged + slot * p.bs;
      if (have > 0) cta_copy(fr + 32, sb + static_cast<int64_t>(i) * per, have, false);
      if (have < per) cta_zero(fr + 32 + have, per - have);  // Split's zero padding (cmd/erasure-coding.go:81)
    }
  }
}

}  // namespace mec

// ---- degraded / plain GETs of many requests that share one reader set ------------------------------------------------------
namespace mec {

struct BatchGetDesc {
  const uint8_t* files[kBatchMaxFiles];  // host (mapped) part files, from their first byte; nullptr = offline
  uint8_t* dst;                          // host (mapped): object bytes [offset, offset + length)
  int64_t offset, length, total;
  int64_t start_block, nblocks;          // blocks of the part that are read
  int64_t slot0;                         // merged layout: slot of the request's first FULL block
  int64_t tail_slot;                     // slot of the part's short last block when the range reaches it, -1 otherwise
  int64_t last_len;                      // shard length of that block
  int32_t dma;                           // 1: the copy engines stage this request's frames (long ranges), the gather kernel skips it
  int32_t pad;
};

struct BatchGetParams {
  const BatchGetDesc* reqs;
  int nreq, k, r;
  int32_t chosen[32];                    // reader set (ascending shard indices), position t -> shard
  int32_t targets[16];                   // rebuilt data shards, output row q -> shard
  int64_t bs, S, P;                      // block size, full shard size, device frame pitch
  uint8_t* arena;                        // survivors: reader position t at arena + t*arena_stride, slot s at + s*P
  int64_t arena_stride;
  const uint8_t* rebuilt;                // rows (slot*r + q)*pitch
  int64_t pitch;
};

__device__ __forceinline__ int64_t get_slot(const BatchGetDesc& r, int64_t b) {
  return (r.tail_slot >= 0 && b == r.nblocks - 1) ? r.tail_slot : r.slot0 + b;
}

// grid (block stride, k reader positions, nreq): frames of the chosen readers -> arena
__global__ void __launch_bounds__(128) batch_get_gather_kernel(const BatchGetParams p) {
  const BatchGetDesc& r = p.reqs[blockIdx.z];
  if (r.dma) return;
  const int t = blockIdx.y;
  const uint8_t* f = r.files[p.chosen[t]];
  const int64_t fstride = 32 + p.S;
  __shared__ __align__(16) uint8_t stage[kStageBytes + 32];
  for (int64_t b = blockIdx.x; b < r.nblocks; b += gridDim.x) {
    const bool is_tail = r.tail_slot >= 0 && b == r.nblocks - 1;
    const int64_t cur = is_tail ? r.last_len : p.S;
    cta_copy_from_host(p.arena + t * p.arena_stride + get_slot(r, b) * p.P, f + (r.start_block + b) * fstride, 32 + cur, stage);
  }
}

// grid (block stride, k data shards, nreq): writeDataBlocks (cmd/erasure-utils.go:42) — the wanted bytes of every data shard, read
// or rebuilt, into the caller's buffer
__global__ void __launch_bounds__(128) batch_get_scatter_kernel(const BatchGetParams p) {
  const BatchGetDesc& r = p.reqs[blockIdx.z];
  const int i = blockIdx.y;
  int t = -1, q = -1;
  for (int x = 0; x < p.k; x++) if (p.chosen[x] == i) t = x;
  for (int x = 0; x < p.r; x++) if (p.targets[x] == i) q = x;
  if (t < 0 && q < 0) return;
  const int64_t hi = r.offset + r.length;
  for (int64_t b = blockIdx.x; b < r.nblocks; b += gridDim.x) {
    const bool is_tail = r.tail_slot >= 0 && b == r.nblocks - 1;
    const int64_t cur = is_tail ? r.last_len : p.S, slot = get_slot(r, b);
    const int64_t B = r.start_block + b, blo = B * p.bs, bhi = blo + p.bs < r.total ? blo + p.bs : r.total;
    const int64_t slo = blo + static_cast<int64_t>(i) * cur, shi = slo + cur < bhi ? slo + cur : bhi;
    const int64_t a = slo > r.offset ? slo : r.offset, e = shi < hi ? shi : hi;
    if (e <= a) continue;
    const uint8_t* src = t >= 0 ? p.arena + t * p.arena_stride + slot * p.P + 32 : p.rebuilt + (slot * p.r + q) * p.pitch;
    cta_copy(r.dst + (a - r.offset), src + (a - slo), e - a, false);
  }
}

}  // namespace mec
