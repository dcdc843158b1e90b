// rtc_compat.h — lets the kernel headers compile both under nvcc and under NVRTC (run-time specialisation of
// the fused kernel for a decode matrix, ec_jit.cc).  NVRTC has no C/C++ standard library and no <cuda.h>.
#pragma once
#ifdef __CUDACC_RTC__
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef int int32_t;
typedef unsigned long uint64_t;
typedef long int64_t;
typedef unsigned long uintptr_t;
namespace std {
template <class T, T v>
struct integral_constant {
  static constexpr T value = v;
  typedef T value_type;
  constexpr operator T() const noexcept { return v; }
};
template <class T, T... I>
struct integer_sequence {};
template <class T, int N, T... I>
struct mec_make_seq : mec_make_seq<T, N - 1, static_cast<T>(N - 1), I...> {};
template <class T, T... I>
struct mec_make_seq<T, 0, I...> { typedef integer_sequence<T, I...> type; };
template <class T, T N>
using make_integer_sequence = typename mec_make_seq<T, static_cast<int>(N)>::type;
}  // namespace std
struct alignas(64) CUtensorMap_st { unsigned long long opaque[16]; };
typedef CUtensorMap_st CUtensorMap;
#else
#include <cuda.h>
#include <cstdint>
#include <utility>
#endif
