// ec_numa.cc — NUMA-local pinned host memory and thread placement for the host boundary.
//
// MinIO's PutObject/GetObject buffers come from a byte pool of aligned slabs (internal/bpool/bpool.go:25-93,
// AllocAligned at :52/:68).  For a GPU behind PCIe the pool wants two more properties: the pages must be pinned
// (page-locked DMA; pageable memory costs a staging copy inside the driver) and they must live on the NUMA node
// the GPU's root port hangs off — on an 8-GPU box with all buffers on one socket, eight links pull through one memory
// controller set and half of them cross the socket interconnect (round-1 SCALE: 0.52 efficiency at N = 8).
//
// No libnuma dependency: mbind(2) / sched_setaffinity(2) directly, topology from sysfs.
#include <cuda_runtime.h>
#include <dirent.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "../../include/minio_ec.h"

namespace {
struct PinnedAlloc {
  size_t bytes;
  bool mmapped;  // mmap + cudaHostRegister (NUMA path) vs cudaHostAlloc
};
std::mutex g_mu;
std::map<void*, PinnedAlloc>& registry() {
  static std::map<void*, PinnedAlloc>* m = new std::map<void*, PinnedAlloc>;
  return *m;
}

int read_int_file(const std::string& path, int dflt) {
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return dflt;
  int v = dflt;
  if (fscanf(f, "%d", &v) != 1) v = dflt;
  fclose(f);
  return v;
}

// "0-31,64-95" -> cpu list
std::vector<int> parse_cpulist(const std::string& path) {
  std::vector<int> cpus;
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return cpus;
  char buf[4096];
  if (fgets(buf, sizeof(buf), f)) {
    const char* p = buf;
    while (*p) {
      while (*p && !isdigit(static_cast<unsigned char>(*p))) p++;
      if (!*p) break;
      char* e;
      long a = strtol(p, &e, 10), b = a;
      p = e;
      if (*p == '-') { b = strtol(p + 1, &e, 10); p = e; }
      for (long c = a; c <= b && c < 4096; c++) cpus.push_back(static_cast<int>(c));
    }
  }
  fclose(f);
  return cpus;
}
}  // namespace

extern "C" int mec_device_numa_node(int device) {
  char bdf[32] = {0};
  if (cudaDeviceGetPCIBusId(bdf, sizeof(bdf), device) != cudaSuccess) { cudaGetLastError(); return -1; }
  for (char* p = bdf; *p; p++) *p = static_cast<char>(tolower(static_cast<unsigned char>(*p)));
  int node = read_int_file(std::string("/sys/bus/pci/devices/") + bdf + "/numa_node", -1);
  if (node < 0) {
    // single-node hosts (and some VMs) report -1: node 0 if it exists
    if (read_int_file("/sys/devices/system/node/node0/meminfo", -2) != -2 || access("/sys/devices/system/node/node0", F_OK) == 0) {
      int nodes = 0;
      if (DIR* d = opendir("/sys/devices/system/node")) {
        while (dirent* e = readdir(d))
          if (!strncmp(e->d_name, "node", 4) && isdigit(static_cast<unsigned char>(e->d_name[4]))) nodes++;
        closedir(d);
      }
      if (nodes == 1) node = 0;
    }
  }
  return node;
}

extern "C" int mec_bind_thread_to_device(int device) {
  const int node = mec_device_numa_node(device);
  if (node < 0) return -1;
  const std::vector<int> cpus = parse_cpulist("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
  if (cpus.empty()) return -1;
  cpu_set_t cur, want;
  CPU_ZERO(&want);
  if (sched_getaffinity(0, sizeof(cur), &cur) != 0) return -1;
  int n = 0;
  for (int c : cpus)
    if (c < CPU_SETSIZE && CPU_ISSET(c, &cur)) { CPU_SET(c, &want); n++; }  // never widen a cgroup / taskset restriction
  if (n == 0) return -1;
  if (sched_setaffinity(0, sizeof(want), &want) != 0) return -1;
  return node;
}

extern "C" void* mec_alloc_pinned_on(int device, size_t bytes) {
  if (bytes == 0) bytes = 1;
  const int node = mec_device_numa_node(device);
  const size_t page = 2u << 20;
  const size_t len = (bytes + page - 1) / page * page;
  void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (p == MAP_FAILED) return nullptr;
  madvise(p, len, MADV_HUGEPAGE);
#ifdef SYS_mbind
  if (node >= 0 && node < 1024) {
    unsigned long mask[16] = {0};
    mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
    // MPOL_PREFERRED (1): fall back to other nodes instead of failing when the node is short of memory
    syscall(SYS_mbind, p, len, 1 /*MPOL_PREFERRED*/, mask, 1024ul, 0u);
  }
#endif
  // first touch (places the pages under the policy above), in parallel: page faults of a multi-GiB buffer are slow from one thread
  {
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt == 0 ? 1 : (nt > 16 ? 16 : nt);
    if (len < (64u << 20)) nt = 1;
    std::vector<std::thread> th;
    const size_t per = (len / nt + page - 1) / page * page;
    for (unsigned t = 0; t < nt; t++) {
      const size_t a = static_cast<size_t>(t) * per, b = a + per < len ? a + per : len;
      if (a >= b) break;
      th.emplace_back([=] {
        for (size_t o = a; o < b; o += 4096) static_cast<volatile char*>(p)[o] = 0;
      });
    }
    for (auto& t : th) t.join();
  }
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(device);
  const cudaError_t e = cudaHostRegister(p, len, cudaHostRegisterPortable);
  cudaSetDevice(cur);
  if (e != cudaSuccess) {
    cudaGetLastError();
    munmap(p, len);
    // registration refused (locked-memory limit, container policy): plain cudaHostAlloc still gives pinned memory
    return mec_alloc_pinned(bytes);
  }
  std::lock_guard<std::mutex> lk(g_mu);
  registry()[p] = PinnedAlloc{len, true};
  return p;
}

extern "C" void* mec_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  std::lock_guard<std::mutex> lk(g_mu);
  registry()[p] = PinnedAlloc{bytes, false};
  return p;
}

extern "C" void mec_free_pinned(void* p) {
  if (!p) return;
  PinnedAlloc a{0, false};
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = registry().find(p);
    if (it == registry().end()) { cudaFreeHost(p); cudaGetLastError(); return; }
    a = it->second;
    registry().erase(it);
  }
  if (a.mmapped) {
    cudaHostUnregister(p);
    munmap(p, a.bytes);
  } else {
    cudaFreeHost(p);
  }
}

// 1 when [p, p + bytes) is page-locked memory known to CUDA (DMA engines copy straight from / into it)
extern "C" int mec_is_pinned(const void* p) {
  {  // buffers from mec_alloc_pinned* are answered from the registry: no driver call (and no driver lock) on the request path
    std::lock_guard<std::mutex> lk(g_mu);
    auto& reg = registry();
    auto it = reg.upper_bound(const_cast<void*>(p));
    if (it != reg.begin()) {
      --it;
      if (static_cast<const char*>(p) < static_cast<const char*>(it->first) + it->second.bytes) return 1;
    }
  }
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return 0; }
  return at.type == cudaMemoryTypeHost ? 1 : 0;
}
