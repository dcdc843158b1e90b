// tests/cpp/test_erasure.cc — the reference's Go test tables, ported as data, run against the C++ mirror
// (include/minio_erasure.hpp) which drives the GPU through the C ABI.  Results are compared bit-exactly with
// the CPU oracle (test infrastructure) where the Go tests only check round trips.
//   TestErasureEncodeDecode        cmd/erasure_test.go:33-118
//   TestErasureEncode              cmd/erasure-encode_test.go:54-161   (badDisk :30-50)
//   TestErasureDecode              cmd/erasure-decode_test.go:44-197
//   TestErasureDecodeRandomOffsetLength  cmd/erasure-decode_test.go:200-  (reduced iteration count)
//   TestErasureHeal                cmd/erasure-heal_test.go:42-155
//   TestAllBitrotAlgorithms        cmd/bitrot_test.go:25-77 (HighwayHash256S — the GPU algorithm)
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <string>
#include "../../include/minio_erasure.hpp"
#include "../../oracle/oracle.h"

using namespace minio;
static int g_fail = 0;
#define CHECK(cond, ...)                                                       \
  do { if (!(cond)) { g_fail++; fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)

constexpr int64_t MiB = 1 << 20;
static std::mt19937_64 rng(42);
static std::vector<uint8_t> randbytes(size_t n) {
  std::vector<uint8_t> v(n);
  for (size_t i = 0; i + 8 <= n; i += 8) { uint64_t r = rng(); memcpy(&v[i], &r, 8); }
  for (size_t i = n & ~size_t(7); i < n; i++) v[i] = static_cast<uint8_t>(rng());
  return v;
}

// ---- in-memory drives (newErasureTestSetup / xlStorage stand-ins) and fault injection (badDisk) ----
struct MemFile : Writer, ReaderAt {
  std::vector<uint8_t> data;
  bool faulty = false;  // badDisk: every I/O fails with errFaultyDisk
  Err Write(const uint8_t* p, size_t n, size_t* w) override {
    if (faulty) return Err::errFaultyDisk;
    data.insert(data.end(), p, p + n);
    if (w) *w = n;
    return Err::nil;
  }
  Err ReadAt(uint8_t* buf, size_t n, int64_t off, size_t* nread) override {
    if (faulty) return Err::errFaultyDisk;
    if (off < 0 || static_cast<size_t>(off) + n > data.size()) return Err::errFileNotFound;
    memcpy(buf, data.data() + off, n);
    if (nread) *nread = n;
    return Err::nil;
  }
};
struct BytesReader : Reader {
  const uint8_t* p; size_t n, pos = 0;
  BytesReader(const uint8_t* p_, size_t n_) : p(p_), n(n_) {}
  size_t Read(uint8_t* buf, size_t want) override { size_t t = std::min(want, n - pos); memcpy(buf, p + pos, t); pos += t; return t; }
};
struct BufWriter : Writer {
  std::vector<uint8_t> data;
  Err Write(const uint8_t* p, size_t n, size_t* w) override { data.insert(data.end(), p, p + n); if (w) *w = n; return Err::nil; }
};

struct Setup {
  std::unique_ptr<Erasure> e;
  int k, m, n;
  std::vector<std::unique_ptr<MemFile>> files;
  std::vector<std::unique_ptr<StreamingBitrotWriter>> bw;
  std::vector<std::unique_ptr<StreamingBitrotReader>> br;
  Setup(int k_, int m_, int64_t bs) : k(k_), m(m_), n(k_ + m_) {
    Err er = Erasure::New(k, m, bs, &e);
    if (er != Err::nil) { fprintf(stderr, "NewErasure failed: %d\n", static_cast<int>(er)); exit(2); }
    for (int i = 0; i < n; i++) files.emplace_back(new MemFile);
  }
  std::vector<Writer*> writers(const std::vector<bool>& offline = {}) {
    bw.clear();
    std::vector<Writer*> w(n, nullptr);
    for (int i = 0; i < n; i++) {
      files[i]->data.clear();
      bw.emplace_back(new StreamingBitrotWriter(files[i].get(), e->ShardSize(), e->codec()));
      if (offline.empty() || !offline[i]) w[i] = bw.back().get();
    }
    return w;
  }
  std::vector<StreamingBitrotReader*> readers(int64_t total, const std::vector<bool>& offline = {}) {
    br.clear();
    std::vector<StreamingBitrotReader*> r(n, nullptr);
    const int64_t till = e->ShardFileOffset(0, total, total);
    for (int i = 0; i < n; i++) {
      br.emplace_back(new StreamingBitrotReader(files[i].get(), till, e->ShardSize(), e->codec()));
      if (offline.empty() || !offline[i]) r[i] = br.back().get();
    }
    return r;
  }
};

static void check_files_vs_oracle(Setup& s, int64_t bs, const std::vector<uint8_t>& data, const char* what) {
  const int n = s.n;
  const int64_t fsz = orc_bitrot_shard_file_size(orc_shard_file_size(bs, s.k, static_cast<int64_t>(data.size())), orc_shard_size(bs, s.k), ORC_HIGHWAYHASH256S);
  std::vector<std::vector<uint8_t>> want(n, std::vector<uint8_t>(static_cast<size_t>(fsz)));
  std::vector<uint8_t*> wp(n);
  for (int i = 0; i < n; i++) wp[i] = want[i].data();
  int64_t rc = orc_erasure_encode(s.k, s.m, bs, ORC_HIGHWAYHASH256S, data.data(), static_cast<int64_t>(data.size()), wp.data(), nullptr);
  CHECK(rc == static_cast<int64_t>(data.size()), "%s: oracle encode rc %lld", what, static_cast<long long>(rc));
  for (int i = 0; i < n; i++)
    if (!s.files[i]->faulty && !s.files[i]->data.empty())
      CHECK(s.files[i]->data == want[i], "%s: shard file %d differs from the oracle", what, i);
}

// ---------------------------------------------------------------------------------------------------
static void TestErasureEncodeDecode() {
  struct T { int k, m, missingData, missingParity; bool reconstructParity, shouldFail; } tests[] = {
      {2, 2, 0, 0, true, false}, {3, 3, 1, 0, true, false}, {4, 4, 2, 0, false, false}, {5, 5, 0, 1, true, false},
      {6, 6, 0, 2, true, false}, {7, 7, 1, 1, false, false}, {8, 8, 3, 2, false, false}, {2, 2, 2, 1, true, true},
      {4, 2, 2, 2, false, true}, {8, 4, 2, 2, false, false}};
  auto data = randbytes(256);
  int idx = 0;
  for (auto& t : tests) {
    std::unique_ptr<Erasure> e;
    CHECK(Erasure::New(t.k, t.m, MiB, &e) == Err::nil, "test %d: NewErasure", idx);
    std::vector<std::vector<uint8_t>> enc;
    CHECK(e->EncodeData(data.data(), data.size(), &enc) == Err::nil, "test %d: EncodeData", idx);
    auto orig = enc;
    for (int j = 0; j < t.missingData; j++) enc[j].clear();
    for (int j = t.k; j < t.k + t.missingParity; j++) enc[j].clear();
    Err er = t.reconstructParity ? e->DecodeDataAndParityBlocks(enc) : e->DecodeDataBlocks(enc);
    CHECK((er != Err::nil) == t.shouldFail, "test %d: reconstruct err=%d shouldFail=%d", idx, static_cast<int>(er), t.shouldFail);
    if (er == Err::nil) {
      const int upto = t.reconstructParity ? t.k + t.m : t.k;
      for (int j = 0; j < upto; j++) CHECK(enc[j] == orig[j], "test %d: shard %d mismatch", idx, j);
    }
    idx++;
  }
  // NewErasure argument validation (cmd/erasure-coding.go:44-50)
  std::unique_ptr<Erasure> e;
  CHECK(Erasure::New(0, 2, MiB, &e) == Err::ErrInvShardNum, "k=0");
  CHECK(Erasure::New(2, -1, MiB, &e) == Err::ErrInvShardNum, "m<0");
  CHECK(Erasure::New(200, 57, MiB, &e) == Err::ErrMaxShardNum, "k+m>256");
}

static void TestErasureEncode() {
  struct T { int dataBlocks, onDisks, offDisks; int64_t blocksize, data; int64_t offset; bool shouldFail, shouldFailQuorum; } tests[] = {
      {2, 4, 0, MiB, MiB, 0, false, false},       {3, 6, 0, MiB, MiB, 1, false, false},        {4, 8, 2, MiB, MiB, 2, false, false},
      {5, 10, 3, MiB, MiB, MiB, false, false},    {6, 12, 4, MiB, MiB, MiB, false, false},     {7, 14, 5, MiB, 0, 0, false, false},
      {8, 16, 7, MiB, 0, 0, false, false},        {2, 4, 2, MiB, MiB, 0, false, true},         {4, 8, 4, MiB, MiB, 0, false, true},
      {7, 14, 7, MiB, MiB, 0, false, true},       {8, 16, 8, MiB, MiB, 0, false, true},        {5, 10, 3, MiB, MiB, 0, false, false},
      {3, 6, 1, MiB, MiB, MiB / 2, false, false}, {2, 4, 0, MiB / 2, MiB, MiB / 2 + 1, false, false}, {4, 8, 0, MiB - 1, MiB, MiB - 1, false, false},
      {8, 12, 2, MiB, MiB, 2, false, false},      {8, 10, 1, MiB, MiB, 0, false, false},       {10, 14, 0, MiB, MiB, 17, false, false},
      {2, 6, 2, MiB, MiB, MiB / 2, false, false}, {10, 16, 8, MiB, MiB, 0, false, true}};
  int idx = 0;
  for (auto& t : tests) {
    Setup s(t.dataBlocks, t.onDisks - t.dataBlocks, t.blocksize);
    auto data = randbytes(static_cast<size_t>(t.data));
    std::vector<uint8_t> part(data.begin() + t.offset, data.end());
    auto w = s.writers();
    BytesReader src(part.data(), part.size());
    int64_t total = 0;
    Err er = s.e->Encode(src, w, t.dataBlocks + 1, &total);
    CHECK((er != Err::nil) == t.shouldFail, "test %d: first encode err=%d", idx, static_cast<int>(er));
    if (er == Err::nil) {
      CHECK(total == static_cast<int64_t>(part.size()), "test %d: wrote %lld want %zu", idx, static_cast<long long>(total), part.size());
      check_files_vs_oracle(s, t.blocksize, part, "TestErasureEncode");
      // second pass: the first offDisks writers are faulty, writer 0 is offline when offDisks > 0
      w = s.writers();
      for (int j = 0; j < t.offDisks; j++) s.files[j]->faulty = true;
      if (t.offDisks > 0) w[0] = nullptr;
      BytesReader src2(part.data(), part.size());
      er = s.e->Encode(src2, w, t.dataBlocks + 1, &total);
      CHECK((er != Err::nil) == t.shouldFailQuorum, "test %d: quorum encode err=%d shouldFailQuorum=%d", idx, static_cast<int>(er), t.shouldFailQuorum);
      if (er == Err::nil) {
        CHECK(total == static_cast<int64_t>(part.size()), "test %d: second pass length", idx);
        check_files_vs_oracle(s, t.blocksize, part, "TestErasureEncode/faulty");
      } else {
        CHECK(er == Err::errErasureWriteQuorum, "test %d: expected errErasureWriteQuorum got %d", idx, static_cast<int>(er));
      }
    }
    idx++;
  }
}

static void TestErasureDecode() {
  // cmd/erasure-decode_test.go:44-83, verbatim minus the algorithm column (the GPU path is HighwayHash256S)
  struct T { int dataBlocks, onDisks, offDisks; int64_t blocksize, data, offset, length; bool shouldFail, shouldFailQuorum; } tests[] = {
      {2, 4, 0, MiB, MiB, 0, MiB, false, false},            {3, 6, 0, MiB, MiB, 0, MiB, false, false},
      {4, 8, 0, MiB, MiB, 0, MiB, false, false},            {5, 10, 0, MiB, MiB, 1, MiB - 1, false, false},
      {6, 12, 0, MiB, MiB, MiB, 0, false, false},           {7, 14, 0, MiB, MiB, 3, 1024, false, false},
      {8, 16, 0, MiB, MiB, 4, 8 * 1024, false, false},      {7, 14, 7, MiB, MiB, MiB, 1, true, false},
      {6, 12, 6, MiB, MiB, 0, MiB, false, false},           {5, 10, 5, MiB, MiB, 0, MiB, false, false},
      {4, 8, 4, MiB, MiB, 0, MiB, false, false},            {3, 6, 3, MiB, MiB, 0, MiB, false, false},
      {2, 4, 2, MiB, MiB, 0, MiB, false, false},            {2, 4, 1, MiB, MiB, 0, MiB, false, false},
      {3, 6, 2, MiB, MiB, 0, MiB, false, false},            {4, 8, 3, 2 * MiB, MiB, 0, MiB, false, false},
      {5, 10, 6, MiB, MiB, 0, MiB, false, true},            {5, 10, 2, MiB, 2 * MiB, MiB, MiB, false, false},
      {5, 10, 1, MiB, MiB, 0, MiB, false, false},           {6, 12, 3, MiB, MiB, 0, MiB, false, false},
      {6, 12, 7, MiB, MiB, 0, MiB, false, true},            {8, 16, 8, MiB, MiB, 0, MiB, false, false},
      {8, 16, 9, MiB, MiB, 0, MiB, false, true},            {8, 16, 7, MiB, MiB, 0, MiB, false, false},
      {2, 4, 1, MiB, MiB, 0, MiB, false, false},            {2, 4, 0, MiB, MiB, 0, MiB, false, false},
      {2, 4, 0, MiB, MiB + 1, 0, MiB + 1, false, false},    {2, 4, 0, MiB, 2 * MiB, 12, MiB + 17, false, false},
      {3, 6, 0, MiB, 2 * MiB, 1023, MiB + 1024, false, false}, {4, 8, 0, MiB, 2 * MiB, 11, MiB + 2 * 1024, false, false},
      {6, 12, 0, MiB, 2 * MiB, 512, MiB + 8 * 1024, false, false}, {8, 16, 0, MiB, 2 * MiB, MiB, MiB - 1, false, false},
      {2, 4, 0, MiB, MiB, -1, 3, true, false},              {2, 4, 0, MiB, MiB, 1024, -1, true, false},
      {4, 6, 0, MiB, MiB, 0, MiB, false, false},            {4, 6, 1, MiB, 2 * MiB, 12, MiB + 17, false, false},
      {4, 6, 3, MiB, 2 * MiB, 1023, MiB + 1024, false, true}, {8, 12, 4, MiB, 2 * MiB, 11, MiB + 2 * 1024, false, false},
      // beyond the reference table: the north-star geometry with a range crossing four blocks
      {12, 16, 4, MiB, 5 * MiB + 77, MiB - 3, 3 * MiB + 50, false, false}};
  int idx = 0;
  for (auto& t : tests) {
    Setup s(t.dataBlocks, t.onDisks - t.dataBlocks, t.blocksize);
    auto data = randbytes(static_cast<size_t>(t.data));
    auto w = s.writers();
    BytesReader src(data.data(), data.size());
    int64_t total = 0;
    CHECK(s.e->Encode(src, w, t.dataBlocks + 1, &total) == Err::nil, "test %d: encode", idx);
    // pass 1: all disks online
    {
      auto r = s.readers(t.data);
      BufWriter out;
      int64_t written = 0;
      Err er = s.e->Decode(out, r, t.offset, t.length, t.data, &written);
      CHECK((er != Err::nil) == t.shouldFail, "test %d: decode err=%d shouldFail=%d", idx, static_cast<int>(er), t.shouldFail);
      if (er == Err::nil) CHECK(out.data == std::vector<uint8_t>(data.begin() + t.offset, data.begin() + t.offset + t.length), "test %d: decoded bytes differ", idx);
    }
    // pass 2: the first offDisks drives are faulty (badDisk), drive 0 offline
    if (!t.shouldFail) {
      for (int j = 0; j < t.offDisks; j++) s.files[j]->faulty = true;
      auto r = s.readers(t.data);
      if (t.offDisks > 0) r[0] = nullptr;
      BufWriter out;
      int64_t written = 0;
      Err er = s.e->Decode(out, r, t.offset, t.length, t.data, &written);
      CHECK((er != Err::nil) == t.shouldFailQuorum, "test %d: degraded decode err=%d shouldFailQuorum=%d", idx, static_cast<int>(er), t.shouldFailQuorum);
      if (er == Err::nil) CHECK(out.data == std::vector<uint8_t>(data.begin() + t.offset, data.begin() + t.offset + t.length), "test %d: degraded bytes differ", idx);
      else CHECK(er == Err::errErasureReadQuorum, "test %d: expected errErasureReadQuorum got %d", idx, static_cast<int>(er));
    }
    idx++;
  }
}

static void TestErasureDecodeRandomOffsetLength() {
  // cmd/erasure-decode_test.go:200: (7,7), 5 MiB object, random ranges; 10000 iterations there, 60 here per run
  const int64_t bs = MiB, length = 5 * MiB;
  Setup s(7, 7, bs);
  auto data = randbytes(static_cast<size_t>(length));
  auto w = s.writers();
  BytesReader src(data.data(), data.size());
  int64_t total = 0;
  CHECK(s.e->Encode(src, w, 8, &total) == Err::nil && total == length, "random: encode");
  for (int it = 0; it < 60; it++) {
    int64_t off = static_cast<int64_t>(rng() % length), len = static_cast<int64_t>(rng() % (length - off));
    auto r = s.readers(length);
    BufWriter out;
    int64_t written = 0;
    Err er = s.e->Decode(out, r, off, len, length, &written);
    CHECK(er == Err::nil && written == len, "random %d: err %d", it, static_cast<int>(er));
    CHECK(out.data == std::vector<uint8_t>(data.begin() + off, data.begin() + off + len), "random %d: bytes differ off=%lld len=%lld", it, static_cast<long long>(off), static_cast<long long>(len));
  }
}

static void TestErasureHeal() {
  // cmd/erasure-heal_test.go:42-61, verbatim minus the algorithm column
  struct T { int dataBlocks, disks, offDisks, badDisks, badStaleDisks; int64_t blocksize, size; bool shouldFail; } tests[] = {
      {2, 4, 1, 0, 0, MiB, MiB, false},      {3, 6, 2, 0, 0, MiB, MiB, false},      {4, 8, 2, 1, 0, MiB, MiB, false},
      {5, 10, 3, 1, 0, MiB, MiB, false},     {6, 12, 2, 3, 0, MiB, MiB, false},     {7, 14, 4, 1, 0, MiB, MiB, false},
      {8, 16, 6, 1, 1, MiB, MiB, false},     {7, 14, 2, 3, 0, MiB / 2, MiB, false}, {6, 12, 1, 0, 1, MiB - 1, MiB, true},
      {5, 10, 3, 0, 3, MiB / 2, MiB, true},  {4, 8, 1, 1, 0, MiB, MiB, false},      {2, 4, 1, 0, 1, MiB, MiB, true},
      {6, 12, 8, 3, 0, MiB, MiB, true},      {7, 14, 3, 4, 0, MiB, MiB, false},     {7, 14, 6, 1, 0, MiB, MiB, false},
      {8, 16, 4, 5, 0, MiB, MiB, true},      {2, 4, 1, 0, 0, MiB, MiB, false},      {12, 16, 2, 1, 0, MiB, MiB, false},
      {6, 8, 1, 0, 0, MiB, MiB, false},      {2, 4, 1, 0, 0, MiB, 64 * MiB, false},
      // beyond the reference table: RS(12,4) with four stale drives and a short last block
      {12, 16, 4, 0, 0, MiB, 8 * MiB + 333, false}};
  int idx = 0;
  for (auto& t : tests) {
    Setup s(t.dataBlocks, t.disks - t.dataBlocks, t.blocksize);
    auto data = randbytes(static_cast<size_t>(t.size));
    auto w = s.writers();
    BytesReader src(data.data(), data.size());
    int64_t total = 0;
    CHECK(s.e->Encode(src, w, t.dataBlocks + 1, &total) == Err::nil, "heal %d: encode", idx);
    std::vector<std::vector<uint8_t>> golden;
    for (auto& f : s.files) golden.push_back(f->data);
    // stale = first offDisks (their readers are nil, writers are fresh files); bad = next badDisks faulty readers
    auto readers = s.readers(t.size);
    std::vector<std::unique_ptr<MemFile>> healed;
    std::vector<std::unique_ptr<StreamingBitrotWriter>> hw;
    std::vector<Writer*> writers(s.n, nullptr);
    for (int j = 0; j < t.offDisks; j++) {
      readers[j] = nullptr;
      healed.emplace_back(new MemFile);
      if (j < t.badStaleDisks) healed.back()->faulty = true;
      hw.emplace_back(new StreamingBitrotWriter(healed.back().get(), s.e->ShardSize(), s.e->codec()));
      writers[j] = hw.back().get();
    }
    for (int j = t.offDisks; j < t.offDisks + t.badDisks; j++) s.files[j]->faulty = true;
    Err er = s.e->Heal(writers, readers, t.size);
    CHECK((er != Err::nil) == t.shouldFail, "heal %d: err=%d shouldFail=%d", idx, static_cast<int>(er), t.shouldFail);
    if (er == Err::nil)
      for (int j = 0; j < t.offDisks; j++)
        if (!healed[j]->faulty) CHECK(healed[j]->data == golden[j], "heal %d: healed shard file %d differs", idx, j);
    idx++;
  }
}

static void TestBitrot() {
  // cmd/bitrot_test.go:25-77 shape: 35 bytes written in 10-byte shards through writer -> "disk" -> reader
  std::unique_ptr<Erasure> e;
  CHECK(Erasure::New(1, 1, 10, &e) == Err::nil, "bitrot: NewErasure");   // blockSize 10, k=1 => shardSize 10
  MemFile disk;
  StreamingBitrotWriter w(&disk, 10, e->codec());
  const char* msgs[] = {"aaaaaaaaaa", "aaaaaaaaaa", "aaaaaaaaaa", "aaaaa"};
  for (auto mtxt : msgs) {
    size_t wr = 0;
    CHECK(w.Write(reinterpret_cast<const uint8_t*>(mtxt), strlen(mtxt), &wr) == Err::nil && wr == strlen(mtxt), "bitrot: write");
  }
  CHECK(static_cast<int64_t>(disk.data.size()) == bitrotShardFileSize(35, 10), "bitrot: file size %zu", disk.data.size());
  size_t wr = 0;
  CHECK(w.Write(reinterpret_cast<const uint8_t*>("x"), 1, &wr) == Err::errBitrotWriteNotAllowed, "bitrot: write after short shard must fail");
  StreamingBitrotReader r(&disk, 35, 10, e->codec());
  uint8_t b[10];
  size_t nr = 0;
  CHECK(r.ReadAt(b, 10, 0, &nr) == Err::nil && memcmp(b, "aaaaaaaaaa", 10) == 0, "bitrot: read 0");
  CHECK(r.ReadAt(b, 10, 10, &nr) == Err::nil, "bitrot: read 10");
  CHECK(r.ReadAt(b, 10, 20, &nr) == Err::nil, "bitrot: read 20");
  CHECK(r.ReadAt(b, 5, 30, &nr) == Err::nil && memcmp(b, "aaaaa", 5) == 0, "bitrot: read 30");
  CHECK(r.ReadAt(b, 5, 3, &nr) == Err::errUnexpected, "bitrot: unaligned offset");
  CHECK(e->BitrotVerify(disk.data.data(), static_cast<int64_t>(disk.data.size()), 35) == Err::nil, "bitrotVerify ok");
  disk.data[40] ^= 1;
  CHECK(r.ReadAt(b, 10, 0, &nr) == Err::errFileCorrupt, "bitrot: corrupt frame must fail");
  CHECK(e->BitrotVerify(disk.data.data(), static_cast<int64_t>(disk.data.size()), 35) == Err::errFileCorrupt, "bitrotVerify corrupt");
}

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void on_segv(int sig) {  // a backtrace instead of a silent exit code when something dies during teardown
  void* frames[64];
  const int n = backtrace(frames, 64);
  const char msg[] = "fatal signal, backtrace:\n";
  (void)!write(2, msg, sizeof(msg) - 1);
  backtrace_symbols_fd(frames, n, 2);
  _exit(128 + sig);
}

int main(int argc, char** argv) {
  signal(SIGSEGV, on_segv);
  signal(SIGABRT, on_segv);
  std::string only = argc > 1 ? argv[1] : "";
  if (mec_device_count() < 1) { fprintf(stderr, "no CUDA device\n"); return 3; }
  struct { const char* name; void (*fn)(); } all[] = {{"TestErasureEncodeDecode", TestErasureEncodeDecode}, {"TestErasureEncode", TestErasureEncode},
                                                      {"TestErasureDecode", TestErasureDecode}, {"TestErasureDecodeRandomOffsetLength", TestErasureDecodeRandomOffsetLength},
                                                      {"TestErasureHeal", TestErasureHeal}, {"TestBitrot", TestBitrot}};
  for (auto& t : all) {
    if (!only.empty() && only != t.name) continue;
    int before = g_fail;
    t.fn();
    printf("%s %s\n", g_fail == before ? "ok  " : "FAIL", t.name);
  }
  mec_shutdown();  // background specialisation must be idle before exit()
  return g_fail ? 1 : 0;
}
