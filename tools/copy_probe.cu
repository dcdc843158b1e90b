// copy_probe: what the copy engine sustains for a train of small pinned H2D copies (the coalescer's staging pattern), alone and
// with a D2H train on another stream.  nvcc -O2 -arch=sm_100a -o copy_probe copy_probe.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
int main() {
  const size_t total = 512ull << 20;
  char *h = nullptr, *h2 = nullptr, *d = nullptr, *d2 = nullptr;
  cudaHostAlloc(&h, total, cudaHostAllocPortable);
  cudaHostAlloc(&h2, total, cudaHostAllocPortable);
  cudaMalloc(&d, total);
  cudaMalloc(&d2, total);
  cudaStream_t s1, s2, s3;
  cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&s3, cudaStreamNonBlocking);
  for (size_t piece : {64u << 10, 256u << 10, 1u << 20, 4u << 20, 16u << 20, 512u << 20}) {
    for (int mode = 0; mode < 5; mode++) {  // 0: H2D alone, 1: H2D + D2H (1/3 of the bytes), 2: H2D split over two streams, 3/4: cudaMemcpyBatchAsync (40 per call) alone / with D2H
      const size_t n = total / piece;
      double best = 0, best_issue = 0;
      for (int rep = 0; rep < 3; rep++) {
        cudaDeviceSynchronize();
        auto t0 = std::chrono::steady_clock::now();
        if (mode >= 3) {
          std::vector<void*> dsts(n), srcs(n);
          std::vector<size_t> sizes(n, piece);
          for (size_t i = 0; i < n; i++) { dsts[i] = d + i * piece; srcs[i] = h + ((i * 7) % n) * piece; }
          cudaMemcpyAttributes at = {};
          at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
          size_t idx0 = 0, fail = 0;
          for (size_t i = 0; i < n; i += 40) {
            const size_t cnt = n - i < 40 ? n - i : 40;
            cudaError_t e = cudaMemcpyBatchAsync(dsts.data() + i, srcs.data() + i, sizes.data() + i, cnt, &at, &idx0, 1, &fail, s1);
            if (e != cudaSuccess) { printf("batch copy: %s\n", cudaGetErrorString(e)); return 1; }
            if (mode == 4)
              for (size_t j = i; j < i + cnt; j += 3) cudaMemcpyAsync(h2 + j * piece, d2 + j * piece, piece, cudaMemcpyDeviceToHost, s2);
          }
        } else
        for (size_t i = 0; i < n; i++) {
          cudaMemcpyAsync(d + i * piece, h + ((i * 7) % n) * piece, piece, cudaMemcpyHostToDevice, (mode == 2 && (i & 1)) ? s3 : s1);
          if (mode == 1 && i % 3 == 0) cudaMemcpyAsync(h2 + i * piece, d2 + i * piece, piece, cudaMemcpyDeviceToHost, s2);
        }
        auto t1 = std::chrono::steady_clock::now();
        cudaStreamSynchronize(s1);
        cudaStreamSynchronize(s3);
        auto t2 = std::chrono::steady_clock::now();
        cudaStreamSynchronize(s2);
        const double sec = std::chrono::duration<double>(t2 - t0).count(), issue = std::chrono::duration<double>(t1 - t0).count();
        if (total / sec > best) { best = total / sec; best_issue = issue / n * 1e6; }
      }
      printf("{\"piece_KiB\": %zu, \"mode\": \"%s\", \"h2d_GBps\": %.1f, \"issue_us_per_copy\": %.2f}\n", piece >> 10,
             mode == 0 ? "h2d" : mode == 1 ? "h2d+d2h/3" : mode == 2 ? "h2d 2 streams" : mode == 3 ? "batch40" : "batch40+d2h/3", best / 1e9, best_issue);
    }
  }
  return 0;
}
