// ec_engine.h — host-side engine behind the C ABI: kernel dispatch, tensor maps, device staging.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>
#include "ec_kernel.cuh"

namespace mec {

void set_last_error(const std::string& s);
const char* get_last_error();

#define MEC_CUDA_OK(expr)                                                                    \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::mec::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e));             \
      return MEC_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

// One fused launch over `nblocks` equally shaped erasure blocks (all pointers are device pointers).
struct FusedDesc {
  int k = 0, r = 0;
  const uint8_t* coef = nullptr;   // r x k row-major runtime matrix (ignored when static_encode)
  bool static_encode = false;      // matrix == parity rows of (k, r): compile-time kernels eligible
  int64_t nblocks = 0;
  int32_t S = 0;
  // input addressing
  bool contiguous = true;          // true: Split layout inside contiguous object bytes
  const uint8_t* in_base = nullptr;  // contiguous: first byte of block 0
  int64_t in_block_stride = 0;     // bytes between consecutive blocks (both kinds)
  int64_t in_block_len = 0;        // contiguous: valid object bytes per block (<= k*S)
  const uint8_t* in_ptr[kMaxK] = {};   // !contiguous: shard t, block 0, byte 0
  const uint8_t* map_base[kMaxK] = {}; // !contiguous: 16B-aligned base of the allocation holding input t
  int64_t map_len[kMaxK] = {};         // !contiguous: addressable bytes from map_base[t]
  // outputs
  uint8_t* out = nullptr;
  int64_t out_pitch = 0;
  uint8_t* digests = nullptr;      // nullptr: skip hashing entirely
  bool hash_outputs = true;        // false: only the k inputs are hashed (GET path: reconstructed data needs no digest)
  const uint8_t* expect_ptr[kMaxK] = {};
  int64_t expect_block_stride = 0;
  uint8_t* corrupt = nullptr;
  const uint8_t* key = nullptr;    // 32-byte HighwayHash key (host)
  const int32_t* block_len = nullptr;  // device: per-block shard bytes (<= S); only launches that take the latency kernel (see small_ok)
  const SmallBlock* blocks = nullptr;  // device, contiguous encode: per-block geometry relative to in_base; S = the largest shard length
  // one block of the launch is an object's short last block (latency kernel only, see small_ok): no table needed
  int64_t tail_block = -1, tail_in_off = -1;  // index; contiguous encode: byte offset from in_base (rows tail_S apart)
  int32_t tail_S = 0, tail_bytes = 0;
};

struct EngineOptions {
  int eb = 0;              // erasure blocks per CTA (0 = auto)
  int force_bytewise = 0;  // 1: never use TMA
  int force_dynamic = 0;   // 1: never use the compile-time specialised GF kernels
  int balance_grid = 0;  // shrink the persistent grid so that every CTA makes the same number of passes
  int grid_mult = 0;       // CTAs per SM (0 = occupancy)
  int jit = -1;            // decode-matrix kernels specialised at run time with NVRTC: -1 auto (large launches), 0 never, 1 always
  int no_rows3d = 0;       // 1: never use the one-request-per-tile 3-D TMA fetch
  int use_auto = 0;        // 1: warp-autonomous pipeline (no CTA barriers) when k + r == 16; measured slower, off by default
  int64_t chunk_blocks = 0; // host pipeline chunk (0 = auto)
  int64_t small_blocks = -1; // launches of at most this many erasure blocks take the latency kernel (ec_small.cuh); -1: 5 per SM, 0: never
  int static_groups = 1;   // 1 (default): erasure-block groups are dealt to CTAs statically (g += gridDim.x); 0: through a claim counter —
                           // measured equal on a dedicated GPU (profiles/r2_kernel_ab.md), useful when SMs are shared or uneven
};

class Engine {
 public:
  explicit Engine(int device);
  ~Engine();
  int init();
  int device() const { return device_; }
  int launch_fused(const FusedDesc& d, const EngineOptions& opt, cudaStream_t st);
  int64_t launches() const { return launches_; }
  void count_launch() { launches_++; }
  int64_t jit_compiles() const;  // process-wide
  int64_t small_limit(const EngineOptions& opt) const { return opt.small_blocks >= 0 ? opt.small_blocks : 5ll * num_sms_; }  // most blocks a latency-kernel launch takes
  bool small_ok(const EngineOptions& opt, int64_t nblocks) const {  // would a launch of nblocks take the latency kernel?
    return nblocks <= small_limit(opt) && opt.eb <= 0 && !opt.force_bytewise && opt.jit != 1 && nblocks < (1ll << 31);
  }
  int64_t small_launches() const { return small_launches_; }  // launches that took the latency kernel (ec_small.cuh)
  int64_t jit_launches() const { return jit_launches_; }  // launches of this engine that ran a specialised kernel
  double jit_seconds() const;
  int64_t jit_disk_hits() const;  // specialised kernels loaded from the on-disk cache instead of compiled (process-wide)
  int num_sms() const { return num_sms_; }
  // queue the specialisation of a reconstruct matrix (r x k rows) on the background compiler, as if the pattern were already warm
  void jit_prewarm(int k, int r, const uint8_t* coef, int eb_t, bool hash_out) {
    jit_kernel(k, r, coef, 0, eb_t, false, hash_out, -1, int64_t{1} << 40);
  }

 private:
  int device_;
  int num_sms_ = 0;
  void* encode_tiled_ = nullptr;  // cuTensorMapEncodeTiled
  int64_t launches_ = 0;
  int64_t jit_launches_ = 0;
  int64_t small_launches_ = 0;
  struct LaunchMemo { const void* fn; int threads; size_t smem; int per_sm; };
  std::vector<LaunchMemo> launch_memo_;
  int raise_smem_limit(const void* kfn, size_t smem);
  uint32_t* claim_slots_ = nullptr;  // ring of per-launch group-claim counters (device)
  uint32_t claim_next_ = 0;
  // run-time specialised kernels, keyed by (k, r, matrix bytes)
  void* jit_kernel(int k, int r, const uint8_t* coef, int align, int eb_t, bool rows3d, bool hash_out, int mode, int64_t in_bytes);
};

// stop background specialisation and wait for a compile in flight (mec_shutdown)
void jit_shutdown();
// NVRTC-compile (not load) the kernel specialised for `coef`; cubin bytes, 0 on a compile error, -1 without libnvrtc
int64_t jit_compile_check(int k, int r, const uint8_t* coef, int align, int eb_t, bool rows3d, bool hash_out);

// grow-only device / pinned buffers
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t n);
  void release();
};

}  // namespace mec
