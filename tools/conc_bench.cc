// conc_bench.cc — the many-small-concurrent-requests regime (cmd/erasure-object.go:1371-1441: one Erasure per request, one block
// per loop iteration, thousands of request goroutines): T host threads each PUT (or degraded-GET) objects of a fixed size.
//   mode batcher : every PUT goes through ONE mec_batcher (cross-request coalescing into merged launches)
//   mode pool    : every thread picks a codec handle from a pool of P (no coalescing: one launch per call)
//   mode get     : degraded GETs (4 data drives offline) through the pool
//   mode cpu     : the C oracle's SIMD encode + HighwayHash on the calling thread (the reference's per-request CPU path)
// Prints one JSON line: aggregate GiB/s, calls/s, p50/p99 latency.  Buffers are pinned (mec_alloc_pinned_on) for the GPU modes.
// Build: g++ -O2 -std=c++17 tools/conc_bench.cc -o tools/conc_bench -Iinclude -Lminio_b200 -lminio_ec -Loracle -loracle -lpthread
#include <algorithm>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "minio_ec.h"
extern "C" void orc_encode_hash_blocks_st(int k, int m, int64_t bs, const uint8_t* src, int64_t nblocks, uint8_t* parity, uint8_t* digests);

static void on_segv(int sig) {
  void* bt[48];
  const int n = backtrace(bt, 48);
  fprintf(stderr, "signal %d, backtrace:\n", sig);
  backtrace_symbols_fd(bt, n, 2);
  _exit(139);
}

int main(int argc, char** argv) {
  signal(SIGSEGV, on_segv);
  signal(SIGABRT, on_segv);
  std::string mode = argc > 1 ? argv[1] : "batcher";
  const int T = argc > 2 ? atoi(argv[2]) : 256;
  // object size: MiB, or KiB with a 'k' suffix (objects smaller than one erasure block: "256k")
  const std::string osz = argc > 3 ? argv[3] : "1";
  const int64_t osize = (!osz.empty() && (osz.back() == 'k' || osz.back() == 'K')) ? atoll(osz.c_str()) << 10 : atoll(osz.c_str()) << 20;
  const int calls_per_thread = argc > 4 ? atoi(argv[4]) : 32;
  const int P = argc > 5 ? atoi(argv[5]) : 8;
  const int device = argc > 6 ? atoi(argv[6]) : 0;
  const int k = 12, m = 4, n = k + m;
  const int64_t bs = 1 << 20;
  const bool gpu = mode != "cpu";
  mec_codec* probe = nullptr;
  mec_codec_new(k, m, bs, MEC_HIGHWAYHASH256S, device, &probe);
  const int64_t S = mec_shard_size(probe), fsz = mec_bitrot_shard_file_size(mec_shard_file_size(probe, osize), S, MEC_HIGHWAYHASH256S);
  const int64_t nb = (osize + bs - 1) / bs;
  if (gpu) mec_bind_thread_to_device(device);
  auto alloc = [&](size_t b) -> uint8_t* { return gpu ? static_cast<uint8_t*>(mec_alloc_pinned_on(device, b)) : static_cast<uint8_t*>(malloc(b)); };
  // per-thread buffers: object, 16 part files, digests, destination
  struct TB { uint8_t* obj; std::vector<uint8_t*> files; uint8_t* dd; uint8_t* dst; uint8_t* parity; };
  std::vector<TB> tb(T);
  uint8_t* arena = alloc(static_cast<size_t>(T) * (osize * 2 + n * fsz + nb * n * 32 + nb * m * S + 4096 * 24));
  if (!arena) { fprintf(stderr, "allocation failed\n"); return 1; }
  uint8_t* p = arena;
  auto take = [&](size_t b) { uint8_t* r = p; p += (b + 4095) / 4096 * 4096; return r; };
  uint64_t x = 88172645463325252ull;
  for (int t = 0; t < T; t++) {
    tb[t].obj = take(osize);
    for (int64_t i = 0; i < osize / 8; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; reinterpret_cast<uint64_t*>(tb[t].obj)[i] = x; }
    for (int i = 0; i < n; i++) tb[t].files.push_back(take(fsz));
    tb[t].dd = take(nb * n * 32);
    tb[t].dst = take(osize);
    tb[t].parity = take(nb * m * S);
  }
  mec_batcher* bat = nullptr;
  std::vector<mec_codec*> pool;
  if (mode == "batcher") {
    int rc = mec_batcher_new(k, m, bs, device, getenv("MAX_BATCH") ? atoll(getenv("MAX_BATCH")) : 512, getenv("MAX_WAIT_US") ? atoi(getenv("MAX_WAIT_US")) : 100, &bat);
    if (rc) { fprintf(stderr, "batcher: %d %s\n", rc, mec_last_error()); return 1; }
  } else if (gpu && mode != "bget") {
    for (int i = 0; i < P; i++) { mec_codec* c = nullptr; mec_codec_new(k, m, bs, MEC_HIGHWAYHASH256S, device, &c); if (getenv("POOL_JIT")) mec_set_option(c, "jit", 1); pool.push_back(c); }
  }
  if (mode == "bget") {  // coalesced GETs: part files through a temporary handle, then everything goes through the batcher
    mec_codec* c0 = nullptr; mec_codec_new(k, m, bs, MEC_HIGHWAYHASH256S, device, &c0);
    for (int t = 0; t < T; t++) {
      std::vector<uint8_t*> f(tb[t].files);
      if (mec_encode(c0, tb[t].obj, osize, f.data(), k) != osize) { fprintf(stderr, "prep encode failed\n"); return 1; }
    }
    mec_codec_free(c0);
    int rc = mec_batcher_new(k, m, bs, device, getenv("MAX_BATCH") ? atoll(getenv("MAX_BATCH")) : 512, getenv("MAX_WAIT_US") ? atoi(getenv("MAX_WAIT_US")) : 100, &bat);
    if (rc) { fprintf(stderr, "batcher: %d\n", rc); return 1; }
  }
  if (mode == "get") {  // every thread's part files must exist first
    for (int t = 0; t < T; t++) {
      std::vector<uint8_t*> f(tb[t].files);
      if (mec_encode(pool[t % P], tb[t].obj, osize, f.data(), k) != osize) { fprintf(stderr, "prep encode failed\n"); return 1; }
    }
  }
  std::atomic<int> ready{0}, go{0};
  std::atomic<int64_t> errors{0};
  std::vector<std::vector<double>> lat(T);
  auto one_call = [&](int t) -> int64_t {
    TB& b = tb[t];
    if (mode == "batcher") return mec_batcher_encode_sg(bat, b.obj, osize, b.files.data(), b.dd, k + 1);
    if (mode == "pool") return mec_encode_sg(pool[t % P], b.obj, osize, b.files.data(), b.dd, k + 1);
    if (mode == "bget") {
      const uint8_t* f[64];
      for (int i = 0; i < n; i++) f[i] = i < 4 ? nullptr : b.files[i];
      int hint = 0;
      return mec_batcher_decode(bat, f, 0, osize, osize, b.dst, &hint);
    }
    if (mode == "get") {
      const uint8_t* f[64];
      for (int i = 0; i < n; i++) f[i] = i < 4 ? nullptr : b.files[i];
      int hint = 0;
      return mec_decode(pool[t % P], f, 0, osize, osize, b.dst, &hint);
    }
    orc_encode_hash_blocks_st(k, m, bs, b.obj, nb, b.parity, b.dd);
    return osize;
  };
  std::vector<std::thread> th;
  for (int t = 0; t < T; t++)
    th.emplace_back([&, t] {
      for (int w = 0; w < 2; w++) if (one_call(t) != osize) errors++;   // warm-up
      ready++;
      while (!go.load()) std::this_thread::yield();
      lat[t].reserve(calls_per_thread);
      for (int i = 0; i < calls_per_thread; i++) {
        const auto t0 = std::chrono::steady_clock::now();
        if (one_call(t) != osize) errors++;
        lat[t].push_back(std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
      }
    });
  while (ready.load() < T) std::this_thread::yield();
  const auto t0 = std::chrono::steady_clock::now();
  go = 1;
  for (auto& t : th) t.join();
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::vector<double> all;
  for (auto& v : lat) all.insert(all.end(), v.begin(), v.end());
  std::sort(all.begin(), all.end());
  const double total = static_cast<double>(T) * calls_per_thread * osize;
  printf("{\"mode\": \"%s\", \"threads\": %d, \"object_MiB\": %lld, \"object_bytes\": %lld, \"calls\": %zu, \"seconds\": %.4f, \"GiB_per_s\": %.2f, \"calls_per_s\": %.0f, "
         "\"p50_us\": %.0f, \"p99_us\": %.0f, \"errors\": %lld",
         mode.c_str(), T, static_cast<long long>(osize >> 20), static_cast<long long>(osize), all.size(), sec, total / sec / (1 << 30), all.size() / sec, all[all.size() / 2],
         all[static_cast<size_t>(all.size() * 0.99)], static_cast<long long>(errors.load()));
  if (bat) printf(", \"batches\": %lld, \"blocks_per_batch\": %.1f, \"launches\": %lld", static_cast<long long>(mec_batcher_stat(bat, "batches")),
                  static_cast<double>(mec_batcher_stat(bat, "blocks")) / std::max<int64_t>(1, mec_batcher_stat(bat, "batches")),
                  static_cast<long long>(mec_batcher_stat(bat, "launches")));
  if (bat && getenv("MEC_BATCHER_TRACE"))
    printf(", \"worker_us\": {\"submit\": %lld, \"sync\": %lld, \"finish\": %lld, \"idle\": %lld}, \"stream_us\": {\"stage\": %lld, \"kernel\": %lld, \"scatter\": %lld}",
           (long long)mec_batcher_stat(bat, "us_submit"), (long long)mec_batcher_stat(bat, "us_sync"), (long long)mec_batcher_stat(bat, "us_finish"),
           (long long)mec_batcher_stat(bat, "us_idle"), (long long)mec_batcher_stat(bat, "us_stage"), (long long)mec_batcher_stat(bat, "us_kernel"),
           (long long)mec_batcher_stat(bat, "us_scatter"));
  if (mode == "pool" || mode == "get") printf(", \"pool\": %d", P);
  printf("}\n");
  if (bat) mec_batcher_free(bat);
  for (auto c : pool) mec_codec_free(c);
  mec_codec_free(probe);
  mec_shutdown();  // no NVRTC compile may be in flight when the C runtime runs its exit handlers (a degraded-GET pattern warms up during the run)
  return errors.load() ? 2 : 0;
}
