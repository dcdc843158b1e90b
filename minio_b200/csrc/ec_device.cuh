// ec_device.cuh — device building blocks of the fused erasure-code + bitrot kernel (sm_100a).
//
//   * packed GF(2^8) arithmetic on 4 bytes per 32-bit register (poly 0x11D — the field of
//     klauspost/reedsolomon, used by cmd/erasure-coding.go:63,85,106,112)
//   * HighwayHash-256 (minio/highwayhash, keyed as cmd/bitrot.go:37,55-58) with TWO threads per
//     hash stream: a thread owns 64-bit lanes {0,1} or {2,3}; the zipper-merge only mixes lanes
//     inside such a pair, so the per-packet update needs no cross-thread traffic at all.  Only the
//     10 finalisation rounds exchange lanes (one shuffle pair per round).
//   * mbarrier / TMA (cp.async.bulk.tensor) wrappers.
#pragma once
#include "rtc_compat.h"
#include "gf256.h"

namespace mec {

#ifndef MEC_XTIME
#define MEC_XTIME 0
#endif
#ifndef MEC_HH_VARIANT
#define MEC_HH_VARIANT 0
#endif

// ---------------------------------------------------------------- small helpers
template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) {
  (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  static_for_impl(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
}

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

// ---------------------------------------------------------------- packed GF(2^8)
// multiply each of the 4 packed field elements by x (i.e. by 2) modulo x^8+x^4+x^3+x^2+1.
// Three instruction mixes with different ALU-pipe / FMA-pipe weight (measured in tools/ubench2.cu):
//   V=0  PRMT sign-replicate + 2 AND (+ shift)      : 3 ALU ops, FMA pipe almost idle
//   V=1  AND + integer SUB + ADD + high-half multiply: 1 ALU op, ~8 FMA-pipe cycles
//   V=2  AND + XOR + shift + high-half multiply      : 2 ALU ops, ~6 FMA-pipe cycles
template <int V>
__device__ __forceinline__ uint32_t gf_xtime4_v(uint32_t a) {
  if constexpr (V == 0) {
    const uint32_t m = prmt(a, 0u, 0xba98u);  // 0xff per byte whose msb is set
    return ((a & 0x7f7f7f7fu) << 1) ^ (m & 0x1d1d1d1du);
  } else if constexpr (V == 1) {
    const uint32_t t = a & 0x7f7f7f7fu;
    uint32_t h;
    asm("sub.u32 %0, %1, %2;" : "=r"(h) : "r"(a), "r"(t));  // == a & 0x80808080, kept as an integer op on purpose
    return (t + t) ^ __umulhi(h, 0x1du << 25);
  } else {
    const uint32_t h = a & 0x80808080u;
    return ((a ^ h) << 1) ^ __umulhi(h, 0x1du << 25);
  }
}
// Division by x: a * x^-1 = (a >> 1) ^ (lsb ? 0x8e : 0), x^-1 = x^7+x^3+x^2+x.  The low bits are peeled off with one
// AND, removed with an integer subtract and turned into the reduction term with a plain 32-bit multiply (both cheap
// FMA-pipe ops; 0/1 bytes times 0x8e cannot carry), so a Horner step in x^-1 costs 2 ALU-pipe ops where the
// doubling above costs 3.  Used with the coefficient decomposition c = sum_b c'_b x^-b (see GfStaticApply).
__device__ __forceinline__ uint32_t gf_xdiv4(uint32_t a) {
  const uint32_t l = a & 0x01010101u;
  uint32_t e;
  asm("sub.u32 %0, %1, %2;" : "=r"(e) : "r"(a), "r"(l));  // clears the low bit of every byte (kept as an integer op)
  return (e >> 1) ^ (l * 0x8eu);
}
#ifndef MEC_GF_DIV
#define MEC_GF_DIV 0   // 1: Horner in x^-1 (gf_xdiv4), 0: Horner in x (gf_xtime4).  Measured: the x^-1 form saves an ALU op per step but c*x^7 is denser than the (sparse) RS coefficients, so it loses overall.
#endif

// One Horner step  acc*x ^ x_terms  in 3 ALU-pipe ops + 1 FMA-pipe op: the byte-msb mask comes from PRMT, the
// doubling is a plain 32-bit add whose cross-byte carry-in bit is masked INSIDE the final LOP3
// ((a+a) & 0xfefefefe == (a & 0x7f7f7f7f) << 1), and the reduction mask is folded into the LOP3 that adds the terms.
__device__ __forceinline__ uint32_t gf_xtime_add4(uint32_t a, uint32_t x_terms) {
  const uint32_t m = prmt(a, 0u, 0xba98u);  // 0xff per byte whose msb is set
  const uint32_t a2 = a + a;
  const uint32_t y = (m & 0x1d1d1d1du) ^ x_terms;
  return (a2 & 0xfefefefeu) ^ y;
}
#ifndef MEC_FUSED_STEP
#define MEC_FUSED_STEP 1   // 1: gf_xtime_add4 (shipped), 0: separate doubling + XOR accumulation
#endif

#ifndef MEC_XMIX_NUM
#define MEC_XMIX_NUM 0   // of every MEC_XMIX_DEN Horner steps, this many use the FMA-heavy variant 1
#endif
#ifndef MEC_XMIX_DEN
#define MEC_XMIX_DEN 4
#endif
__device__ __forceinline__ uint32_t gf_xtime4(uint32_t a) { return gf_xtime4_v<MEC_XTIME>(a); }

// Compile-time specialised  out[j] = XOR_t  coef(j,t) (x) in[t]   on packed words.
// Horner over the bit planes of the coefficients: 7 doublings per OUTPUT word (not per input), and
// the plane sums are assembled from "four Russians" combinations of input triples, so a
// (12 -> 4) product costs ~210 integer ops per 4-byte column instead of 48 table multiplies.
// bit of coefficient c that multiplies Horner plane `plane`: in the x^-1 scheme plane b carries x^-b and
// c = sum_b c'_b x^-b with c'_b = bit (7-b) of c*x^7; in the x scheme it is simply bit `plane` of c.
__host__ __device__ constexpr int plane_bit(uint8_t c, int plane) {
#if MEC_GF_DIV
  return (gf_mul(c, 0x80) >> (7 - plane)) & 1;
#else
  return (c >> plane) & 1;
#endif
}

#ifndef MEC_GF_GROUP
#define MEC_GF_GROUP 4   // inputs per "four Russians" group (3 or 4); 4 needs ~13 % fewer XORs for RS(12,4)
#endif

template <class MAT>  // MAT::K, MAT::R, static constexpr uint8_t MAT::coef(j, t)
struct GfStaticApply {
  static constexpr int K = MAT::K, R = MAT::R, GS = MEC_GF_GROUP, G = (K + GS - 1) / GS, NC = 1 << GS;
  // index of the XOR combination of group g that feeds output j at Horner plane `plane`
  __host__ __device__ static constexpr int combo_index(int j, int g, int plane) {
    int idx = 0;
    for (int q = 0; q < GS; q++)
      if (GS * g + q < K && plane_bit(MAT::coef(j, GS * g + q), plane)) idx |= 1 << q;
    return idx;
  }
  __device__ __forceinline__ static void run(const uint32_t (&in)[K], uint32_t (&out)[R]) {
    uint32_t cmb[G][NC];
    static_for<G>([&](auto g_) {
      constexpr int g = decltype(g_)::value;
      cmb[g][0] = 0u;
      static_for<NC - 1>([&](auto i_) {
        constexpr int idx = decltype(i_)::value + 1;
        constexpr int low = idx & -idx;              // lowest set bit
        constexpr int q = (low == 1) ? 0 : (low == 2) ? 1 : (low == 4) ? 2 : 3;
        constexpr int t = GS * g + q;
        const uint32_t v = (t < K) ? in[t < K ? t : 0] : 0u;
        cmb[g][idx] = cmb[g][idx & (idx - 1)] ^ v;   // unused combinations are dead code
      });
    });
    static_for<R>([&](auto j_) {
      constexpr int j = decltype(j_)::value;
      uint32_t acc = 0u;
      static_for<8>([&](auto bb_) {
        constexpr int step = decltype(bb_)::value;  // 0 = innermost plane of the Horner scheme
        constexpr int plane = 7 - step;
#if MEC_FUSED_STEP && !MEC_GF_DIV
        uint32_t terms = 0u;
        static_for<G>([&](auto g_) {
          constexpr int g = decltype(g_)::value;
          constexpr int idx = combo_index(j, g, plane);
          if constexpr (idx != 0) terms ^= cmb[g][idx];
        });
        if constexpr (step != 0) acc = gf_xtime_add4(acc, terms);
        else acc = terms;
#else
#if MEC_GF_DIV
        if constexpr (step != 0) acc = gf_xdiv4(acc);
#else
        if constexpr (step != 0) {
          constexpr bool fma_heavy = ((plane * R + j) % MEC_XMIX_DEN) < MEC_XMIX_NUM;
          acc = fma_heavy ? gf_xtime4_v<1>(acc) : gf_xtime4_v<MEC_XTIME>(acc);
        }
#endif
        static_for<G>([&](auto g_) {
          constexpr int g = decltype(g_)::value;
          constexpr int idx = combo_index(j, g, plane);
          if constexpr (idx != 0) acc ^= cmb[g][idx];
        });
#endif
      });
      out[j] = acc;
    });
  }
};

// parity rows of reedsolomon.New(K, M): coef(j, t) = M[K + j][t]
template <int K_, int M_>
struct EncodeMatrix {
  static constexpr int K = K_, R = M_;
  static constexpr CodingMatrix<K_, M_> mat = build_coding_matrix<K_, M_>();
  __host__ __device__ static constexpr uint8_t coef(int j, int t) { return mat.v[K_ + j][t]; }
};

// ---------------------------------------------------------------- HighwayHash-256, half state
struct HHHalf {
  uint64_t v0[2], v1[2], m0[2], m1[2];
};

__device__ __forceinline__ uint64_t rot32(uint64_t x) { return (x >> 32) | (x << 32); }
__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi) {
  return static_cast<uint64_t>(lo) | (static_cast<uint64_t>(hi) << 32);
}

__device__ __forceinline__ void hh_init(HHHalf& s, const uint64_t (&key)[4], int h) {
  const uint64_t i0a = h ? 0x13198a2e03707344ull : 0xdbe6d5d5fe4cce2full;
  const uint64_t i0b = h ? 0x243f6a8885a308d3ull : 0xa4093822299f31d0ull;
  const uint64_t i1a = h ? 0xbe5466cf34e90c6cull : 0x3bd39e10cb0ef593ull;
  const uint64_t i1b = h ? 0x452821e638d01377ull : 0xc0acf169b5f18a8cull;
  const uint64_t ka = h ? key[2] : key[0], kb = h ? key[3] : key[1];
  s.m0[0] = i0a; s.m0[1] = i0b; s.m1[0] = i1a; s.m1[1] = i1b;
  s.v0[0] = i0a ^ ka; s.v0[1] = i0b ^ kb;
  s.v1[0] = i1a ^ rot32(ka); s.v1[1] = i1b ^ rot32(kb);
}

// ZipperMergeAndAdd(v1 = hi lane, v0 = lo lane): the two 64-bit addends; byte shuffles as PRMTs
__device__ __forceinline__ void hh_zipper(uint64_t hi, uint64_t lo, uint64_t& z1, uint64_t& z0) {
  const uint32_t v0l = static_cast<uint32_t>(lo), v0h = static_cast<uint32_t>(lo >> 32);
  const uint32_t v1l = static_cast<uint32_t>(hi), v1h = static_cast<uint32_t>(hi >> 32);
  // add0 += [v0.b3, v1.b4, v0.b2, v0.b5 | v1.b6, v0.b1, v1.b7, v0.b0]
  // add1 += [v1.b3, v0.b4, v1.b2, v1.b5 | v1.b1, v0.b6, v1.b0, v0.b7]
  // the four bytes both low words take from the high halves are gathered once: 5 PRMTs instead of 6
  const uint32_t x = prmt(v0h, v1h, 0x5041u);  // [v0.b5, v1.b4, v0.b4, v1.b5]
  const uint32_t a0l = prmt(v0l, x, 0x4253u);
  const uint32_t a0h = prmt(v0l, v1h, 0x0716u);
  const uint32_t a1l = prmt(v1l, x, 0x7263u);
  const uint32_t a1h = prmt(v1l, v0h, 0x7061u);
  z0 = pack64(a0l, a0h);
  z1 = pack64(a1l, a1h);
}

// 32 x 32 -> 64 multiply.  MEC_HH_MUL: 0 = mul.wide (IMAD.WIDE), 1 = mul.lo + mul.hi, 2 = mul.wide via PTX with
// an explicit unpack (the plain C form makes ptxas copy the high word through the ALU before the XOR)
#ifndef MEC_HH_MUL
#define MEC_HH_MUL 1
#endif
__device__ __forceinline__ uint64_t hh_mul32(uint32_t a, uint32_t b) {
#if MEC_HH_MUL == 1
  return pack64(a * b, __umulhi(a, b));
#elif MEC_HH_MUL == 2
  uint32_t lo, hi;
  asm("{\n.reg .b64 t;\nmul.wide.u32 t, %2, %3;\nmov.b64 {%0, %1}, t;\n}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
  return pack64(lo, hi);
#else
  return static_cast<uint64_t>(a) * b;
#endif
}

// 64-bit add routed through the FMA pipe: x + y as IMAD.WIDE(y.lo, one, x) + (y.hi << 32).  `one` is
// the value 1 loaded from kernel parameters, so the compiler cannot fold the multiply back into IADD3.
__device__ __forceinline__ uint64_t add64_fma(uint64_t x, uint64_t y, uint32_t one) {
  return static_cast<uint64_t>(static_cast<uint32_t>(y)) * one + x + ((y >> 32) << 32);
}

// HighwayHash Update for the two lanes a thread owns.  MEC_HH_VARIANT moves 64-bit additions from the
// ALU pipe (IADD3/IADD3.X) to the FMA pipe: 0 = none, 1 = the zipper->v1 add, 2 = also v1 += mul0 + packet.
__device__ __forceinline__ void hh_update(HHHalf& s, uint64_t a0, uint64_t a1, uint32_t one = 1u) {
  const uint64_t a[2] = {a0, a1};
  uint64_t m1old[2];
#pragma unroll
  for (int i = 0; i < 2; i++) {
#if MEC_HH_VARIANT >= 2
    s.v1[i] = add64_fma(add64_fma(s.v1[i], a[i], one), s.m0[i], one);
#else
    s.v1[i] += s.m0[i] + a[i];
#endif
    s.m0[i] ^= hh_mul32(static_cast<uint32_t>(s.v1[i]), static_cast<uint32_t>(s.v0[i] >> 32));
    // only the low word of (v0 += mul1) feeds the multiply; the full 64-bit add is merged with the
    // zipper addend below into one 3-input carry chain
    m1old[i] = s.m1[i];
    const uint32_t v0lo = static_cast<uint32_t>(s.v0[i]) + static_cast<uint32_t>(m1old[i]);
    s.m1[i] ^= hh_mul32(v0lo, static_cast<uint32_t>(s.v1[i] >> 32));
  }
  uint64_t z1, z0;
  hh_zipper(s.v1[1], s.v1[0], z1, z0);
  s.v0[0] = s.v0[0] + m1old[0] + z0;
  s.v0[1] = s.v0[1] + m1old[1] + z1;
  hh_zipper(s.v0[1], s.v0[0], z1, z0);
#if MEC_HH_VARIANT >= 1
  s.v1[0] = add64_fma(s.v1[0], z0, one);
  s.v1[1] = add64_fma(s.v1[1], z1, one);
#else
  s.v1[0] += z0;
  s.v1[1] += z1;
#endif
}

// byte `pos` (0..31) of the padded remainder packet for a tail of n (1..31) bytes
template <class GetByte>
__device__ __forceinline__ uint32_t hh_rem_byte(int pos, int n, GetByte&& tail) {
  const int n4 = n & ~3, m4 = n & 3;
  if (pos < n4) return tail(pos);
  if (n & 16) {
    if (pos >= 28) return tail(n - 4 + (pos - 28));
  } else if (m4) {
    if (pos == 16) return tail(n4);
    if (pos == 17) return tail(n4 + (m4 >> 1));
    if (pos == 18) return tail(n4 + m4 - 1);
  }
  return 0u;
}

template <class GetByte>
__device__ __forceinline__ void hh_remainder(HHHalf& s, int h, int n, GetByte&& tail) {
  const uint64_t inc = (static_cast<uint64_t>(n) << 32) + static_cast<uint64_t>(n);
#pragma unroll
  for (int i = 0; i < 2; i++) {
    s.v0[i] += inc;
    const uint32_t lo = static_cast<uint32_t>(s.v1[i]), hi = static_cast<uint32_t>(s.v1[i] >> 32);
    s.v1[i] = pack64(__funnelshift_l(lo, lo, n), __funnelshift_l(hi, hi, n));
  }
  uint64_t a[2] = {0, 0};
  for (int i = 0; i < 16; i++) {
    const uint64_t b = hh_rem_byte(16 * h + i, n, tail);
    a[i >> 3] |= b << (8 * (i & 7));
  }
  hh_update(s, a[0], a[1]);
}

// 10 permute-and-update rounds + modular reduction; thread h==0 returns digest bytes 0..15,
// h==1 bytes 16..31 (as two little-endian u64).  Lanes of a pair must be adjacent (xor 1).
__device__ __forceinline__ void hh_finalize(HHHalf& s, uint64_t& d0, uint64_t& d1) {
#pragma unroll 1
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = rot32(__shfl_xor_sync(0xffffffffu, s.v0[0], 1));
    const uint64_t p1 = rot32(__shfl_xor_sync(0xffffffffu, s.v0[1], 1));
    hh_update(s, p0, p1);
  }
  const uint64_t a3 = (s.v1[1] + s.m1[1]) & 0x3FFFFFFFFFFFFFFFull;
  const uint64_t a2 = s.v1[0] + s.m1[0];
  const uint64_t a1 = s.v0[1] + s.m0[1];
  const uint64_t a0 = s.v0[0] + s.m0[0];
  d1 = a1 ^ ((a3 << 1) | (a2 >> 63)) ^ ((a3 << 2) | (a2 >> 62));
  d0 = a0 ^ (a2 << 1) ^ (a2 << 2);
}

// ---------------------------------------------------------------- mbarrier + TMA
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // try_wait suspends the thread for a hardware-bounded time; a transfer that never completes (bad
  // tensor map) traps after ~2^22 retries instead of hanging the GPU.
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; spin++) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.b32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 22)) __trap();
  }
}
// one elected lane of a fully converged warp (keeps TMA operands in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "elect.sync _|P1, 0xffffffff;\n"
      "selp.b32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// 2-D tiled TMA load: box lands at `dst` (shared, 128-byte aligned), completion on `bar`
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

// 3-D tiled TMA load (x, block, shard row)
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1, int32_t c2,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

}  // namespace mec
