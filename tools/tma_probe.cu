// tma_probe.cu — does cp.async.bulk.tensor accept box starts that are not 16-byte aligned?  (u16 / u8 elements)
// build: nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, int box_bytes, int rows, uint8_t* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t b = static_cast<uint32_t>(__cvta_generic_to_shared(&bar));
  const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(box_bytes * rows) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(&map), "r"(x), "r"(y), "r"(b)
                 : "memory");
  }
  __syncthreads();
  uint32_t done = 0;
  for (int spin = 0; !done && spin < (1 << 22); spin++)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(b) : "memory");
  if (!done) { if (threadIdx.x == 0) printf("timeout\n"); return; }
  for (int i = threadIdx.x; i < box_bytes * rows; i += blockDim.x) out[i] = smem[i];
}

int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;  // an illegal instruction is sticky: one case per process
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return 2;
  const size_t row = 1 << 20, nrows = 8;
  std::vector<uint8_t> h(row * nrows);
  for (size_t i = 0; i < h.size(); i++) h[i] = static_cast<uint8_t>((i * 2654435761u) >> 13);
  uint8_t *d, *o;
  cudaMalloc(&d, h.size());
  cudaMalloc(&o, 1 << 16);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  struct Case { CUtensorMapDataType dt; int esz; int box_el; int x_el; const char* name; } cases[] = {
      {CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 144, 87382 / 2, "u16 box 288B at byte 87382 (mod 16 = 6)"},
      {CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 144, 3, "u16 box 288B at byte 6"},
      {CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 144, 8, "u16 box 288B at byte 16 (aligned)"},
      {CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 256, 87382 * 5, "u8 box 256B at byte 436910 (mod 16 = 14)"},
      {CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 256, 7, "u8 box 256B at byte 7"},
      {CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 72, 3, "u32 box 288B at byte 12"},
  };
  int idx = -1;
  for (auto& c : cases) {
    idx++;
    if (only >= 0 && idx != only) continue;
    CUtensorMap m;
    cuuint64_t dims[2] = {row / c.esz, nrows};
    cuuint64_t strides[1] = {row};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(c.box_el), 4};
    cuuint32_t es[2] = {1, 1};
    CUresult r = reinterpret_cast<EncodeFn>(fn)(&m, c.dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%-48s encode failed %d\n", c.name, static_cast<int>(r)); continue; }
    const int bb = c.box_el * c.esz;
    cudaMemset(o, 0xEE, 1 << 16);
    probe<<<1, 128, bb * 4>>>(m, c.x_el, 2, bb, 4, o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-48s kernel error: %s\n", c.name, cudaGetErrorString(e)); return 1; }
    std::vector<uint8_t> got(bb * 4);
    cudaMemcpy(got.data(), o, got.size(), cudaMemcpyDeviceToHost);
    bool ok = true;
    for (int rr = 0; rr < 4 && ok; rr++)
      for (int i = 0; i < bb; i++)
        if (got[rr * bb + i] != h[(2 + rr) * row + static_cast<size_t>(c.x_el) * c.esz + i]) { ok = false; break; }
    printf("%-48s %s\n", c.name, ok ? "OK: bytes land re-aligned" : "MISMATCH");
  }
  return 0;
}
