// test_gfplan.cu — runs the compile-time GF(2^8) plans of ec_device.cuh (GfStaticApply: bit-plane Horner, four-Russians
// groups, subset-sum change of basis) ON THE HOST and compares every output word with a plain table multiply.
// Built by nvcc as a host-only executable (no GPU needed): the same template code the kernels instantiate, with PRMT
// restated in C.  Covers every compile-time (k, m) of ec_engine.cu plus decode matrices of several erasure patterns,
// at every transform level and group size, and prints the op-count estimate the level choice is based on.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "../../minio_b200/csrc/ec_device.cuh"
#include "gfplan_cases.inc"
using namespace mec;

static uint32_t rng_state = 0x12345678u;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 17; rng_state ^= rng_state << 5; return rng_state; }

template <class MAT>
struct CoefTable {
  uint8_t v[MAT::R][MAT::K];
};
template <class MAT>
constexpr CoefTable<MAT> make_table() {
  CoefTable<MAT> t{};
  for (int j = 0; j < MAT::R; j++)
    for (int c = 0; c < MAT::K; c++) t.v[j][c] = MAT::coef(j, c);
  return t;
}
template <class MAT>
static void reference(const uint32_t* in, uint32_t* out) {
  static constexpr CoefTable<MAT> tbl = make_table<MAT>();
  for (int j = 0; j < MAT::R; j++) {
    uint32_t w = 0;
    for (int b = 0; b < 4; b++) {
      uint8_t acc = 0;
      for (int t = 0; t < MAT::K; t++) acc ^= gf_mul(tbl.v[j][t], static_cast<uint8_t>(in[t] >> (8 * b)));
      w |= static_cast<uint32_t>(acc) << (8 * b);
    }
    out[j] = w;
  }
}

template <class MAT, int L, int GS>
static int check_one(const char* name) {
  if constexpr ((1 << L) > MAT::K || (1 << L) > MAT::R) return 0;
  else {
    int bad = 0;
    for (int it = 0; it < 2000 && !bad; it++) {
      uint32_t in[MAT::K], want[MAT::R], got[MAT::R];
      for (int t = 0; t < MAT::K; t++) in[t] = it == 0 ? 0u : (it == 1 ? 0xffffffffu : (it < 40 ? (0x80u << (8 * (it & 3))) * ((it >> 2) == t) : rnd()));
      reference<MAT>(in, want);
      GfXform<MAT, L>::template run<GS>(in, got);
      for (int j = 0; j < MAT::R; j++)
        if (want[j] != got[j]) { printf("MISMATCH %s L=%d GS=%d iter %d row %d: want %08x got %08x\n", name, L, GS, it, j, want[j], got[j]); bad = 1; }
    }
    return bad;
  }
}

template <class MAT>
static int check(const char* name) {
  int bad = 0;
  bad |= check_one<MAT, 0, 3>(name); bad |= check_one<MAT, 0, 4>(name);
  bad |= check_one<MAT, 1, 3>(name); bad |= check_one<MAT, 1, 4>(name);
  bad |= check_one<MAT, 2, 3>(name); bad |= check_one<MAT, 2, 4>(name);
  bad |= check_one<MAT, 3, 3>(name); bad |= check_one<MAT, 3, 4>(name);
  // the shipped choice
  for (int it = 0; it < 2000; it++) {
    uint32_t in[MAT::K], want[MAT::R], got[MAT::R];
    for (int t = 0; t < MAT::K; t++) in[t] = rnd();
    reference<MAT>(in, want);
    GfStaticApply<MAT>::run(in, got);
    for (int j = 0; j < MAT::R; j++) bad |= want[j] != got[j];
  }
  printf("%-16s k=%2d r=%2d  level %d group %d  ~%3d ALU ops per word column (plain: %3d)  %s\n", name, MAT::K, MAT::R, GfStaticApply<MAT>::kLevel,
         GfStaticApply<MAT>::kGroup, GfStaticApply<MAT>::kOps,
         GfXform<MAT, 0>::template cost<3>() < GfXform<MAT, 0>::template cost<4>() ? GfXform<MAT, 0>::template cost<3>() : GfXform<MAT, 0>::template cost<4>(),
         bad ? "FAIL" : "ok");
  return bad;
}

int main() {
  int bad = 0;
#define X(K, M) bad |= check<EncodeMatrix<K, M>>("encode(" #K "," #M ")");
  X(12, 4) X(4, 2) X(16, 4) X(8, 8) X(8, 4) X(6, 2) X(2, 2) X(10, 4) X(7, 5) X(5, 3) X(14, 2) X(3, 1) X(1, 1)
#undef X
#define X(NAME) bad |= check<NAME>(#NAME);
  GFPLAN_DECODE_CASES(X)
#undef X
  if (bad) { printf("FAILED\n"); return 1; }
  printf("all GF plans match the table multiply\n");
  return 0;
}
