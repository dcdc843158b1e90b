// minio_erasure.hpp — C++17 host-side mirror of MinIO's erasure / bitrot surfaces above the C ABI
// (include/minio_ec.h).  The image has no Go toolchain, so this header plays the role of the Go
// methods a cgo build would keep: same names, argument meaning and error behaviour.
//
//   Erasure                  cmd/erasure-coding.go:35-141
//   Erasure::Encode          cmd/erasure-encode.go:69-110   (+ multiWriter.Write :34-67)
//   Erasure::Decode / Heal   cmd/erasure-decode.go:239-364  (+ parallelReader.Read :127-235,
//                                                             writeDataBlocks cmd/erasure-utils.go:42-105)
//   StreamingBitrotWriter    cmd/bitrot-streaming.go:32-75, newStreamingBitrotWriter :108
//   StreamingBitrotReader    cmd/bitrot-streaming.go:141-213
//   bitrotVerify             cmd/bitrot.go:164-216, bitrotShardFileSize :156
//
// All arithmetic happens on the GPU through libminio_ec.so; this layer only moves bytes and applies the
// reference's quorum / fail-over rules.  Blocks are batched (kBatchBlocks per GPU call) — the shard files
// produced are byte-identical to the block-at-a-time loop of the reference.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#include "minio_ec.h"

namespace minio {

enum class Err : int {  // the sentinel errors the reference compares with errors.Is
  nil = 0,
  ErrInvShardNum = MEC_ERR_INV_SHARD_NUM,
  ErrMaxShardNum = MEC_ERR_MAX_SHARD_NUM,
  ErrTooFewShards = MEC_ERR_TOO_FEW_SHARDS,
  ErrShardNoData = MEC_ERR_SHARD_NO_DATA,
  ErrShardSize = MEC_ERR_SHARD_SIZE,
  ErrShortData = MEC_ERR_SHORT_DATA,
  errFileCorrupt = MEC_ERR_FILE_CORRUPT,      // cmd/storage-errors.go:104
  errLessData = MEC_ERR_LESS_DATA,            // :114
  errUnexpected = MEC_ERR_UNEXPECTED,         // :29
  errErasureReadQuorum = MEC_ERR_READ_QUORUM,   // cmd/erasure-errors.go:23
  errErasureWriteQuorum = MEC_ERR_WRITE_QUORUM, // :26
  errInvalidArgument = MEC_ERR_INVALID_ARGUMENT,
  errDiskNotFound = -20,   // cmd/storage-errors.go:53
  errFaultyDisk = -21,     // :65
  errFileNotFound = -22,   // :71
  ErrShortWrite = -23,     // io.ErrShortWrite
  errBitrotWriteNotAllowed = -24,  // "bitrot write not allowed" bitrot-streaming.go:49
  errUnexpectedBitrotSize = -25,   // "unexpected bitrot buffer size" :52
  errCuda = MEC_ERR_CUDA,
  errNoDevice = MEC_ERR_NO_DEVICE,
  errUnsupported = MEC_ERR_UNSUPPORTED,
};
inline Err fromRc(int64_t rc) { return rc >= 0 ? Err::nil : static_cast<Err>(static_cast<int>(rc)); }

// io.Writer / io.ReaderAt / io.Reader
struct Writer {
  virtual ~Writer() = default;
  virtual Err Write(const uint8_t* p, size_t n, size_t* written) = 0;
};
struct ReaderAt {
  virtual ~ReaderAt() = default;
  virtual Err ReadAt(uint8_t* buf, size_t n, int64_t off, size_t* nread) = 0;
};
struct Reader {  // returns bytes read; 0 => EOF
  virtual ~Reader() = default;
  virtual size_t Read(uint8_t* buf, size_t n) = 0;
};

inline int64_t ceilFrac(int64_t a, int64_t b) { return mec_ceil_frac(a, b); }                       // cmd/utils.go:689
inline int64_t bitrotShardFileSize(int64_t size, int64_t shardSize, int algo = MEC_HIGHWAYHASH256S) { // cmd/bitrot.go:156
  return mec_bitrot_shard_file_size(size, shardSize, algo);
}

// ---- streaming bitrot writer (cmd/bitrot-streaming.go:32-75) ---------------------------------------
// Emits [32-byte HighwayHash256][shard] per Write.  The fused GPU call already produced the digest, so
// Erasure::Encode/Heal use WriteHashed; plain Write hashes through the library (hash.Hash shape).
class StreamingBitrotWriter : public Writer {
 public:
  StreamingBitrotWriter(Writer* iow, int64_t shardSize, mec_codec* c) : iow_(iow), shardSize_(shardSize), codec_(c) {}
  Err WriteHashed(const uint8_t* digest, const uint8_t* p, size_t n, size_t* written) {
    if (written) *written = 0;
    if (n == 0) return Err::nil;                                          // :45-47
    if (finished_) return Err::errBitrotWriteNotAllowed;                  // :48-50
    if (static_cast<int64_t>(n) > shardSize_) return Err::errUnexpectedBitrotSize;  // :51-53
    if (static_cast<int64_t>(n) < shardSize_) finished_ = true;           // :54-56
    size_t w = 0;
    Err e = iow_->Write(digest, 32, &w);                                  // :60
    if (e != Err::nil) return e;
    e = iow_->Write(p, n, &w);                                            // :65
    if (e != Err::nil) return e;
    if (w != n) return Err::ErrShortWrite;
    if (written) *written = n;
    return Err::nil;
  }
  Err Write(const uint8_t* p, size_t n, size_t* written) override {
    if (n == 0) { if (written) *written = 0; return Err::nil; }
    uint8_t dg[32];
    int rc = mec_hh256_batch(codec_, p, static_cast<int64_t>(n), 1, dg);  // h.Reset(); h.Write(p); h.Sum(nil)  :57-59
    if (rc) return fromRc(rc);
    return WriteHashed(dg, p, n, written);
  }

 private:
  Writer* iow_;
  int64_t shardSize_;
  mec_codec* codec_;
  bool finished_ = false;
};

// ---- streaming bitrot reader (cmd/bitrot-streaming.go:141-213) --------------------------------------
// ReadFrameAt hands the raw [digest][shard] frame to the caller; verification happens in the fused kernel.
class StreamingBitrotReader : public ReaderAt {
 public:
  StreamingBitrotReader(ReaderAt* disk, int64_t tillOffset, int64_t shardSize, mec_codec* c)
      : disk_(disk), shardSize_(shardSize), codec_(c) {
    tillOffset_ = ceilFrac(tillOffset, shardSize) * 32 + tillOffset;      // :206
  }
  // offset is a shard-file offset (multiple of shardSize); frame = 32 + n bytes
  Err ReadFrameAt(uint8_t* frame, size_t n, int64_t offset) {
    if (offset % shardSize_ != 0) return Err::errUnexpected;              // :163-167
    const int64_t streamOffset = (offset / shardSize_) * 32 + offset;     // :171
    if (streamOffset + 32 + static_cast<int64_t>(n) > tillOffset_) return Err::errUnexpected;
    size_t got = 0;
    Err e = disk_->ReadAt(frame, 32 + n, streamOffset, &got);
    if (e != Err::nil) return e;
    if (got != 32 + n) return Err::errFileCorrupt;
    return Err::nil;
  }
  Err ReadAt(uint8_t* buf, size_t n, int64_t offset, size_t* nread) override {  // verified read (:161-200)
    std::vector<uint8_t> frame(32 + n);
    Err e = ReadFrameAt(frame.data(), n, offset);
    if (e != Err::nil) return e;
    uint8_t dg[32];
    int rc = mec_hh256_batch(codec_, frame.data() + 32, static_cast<int64_t>(n), 1, dg);
    if (rc) return fromRc(rc);
    if (memcmp(dg, frame.data(), 32) != 0) return Err::errFileCorrupt;    // :194-196
    memcpy(buf, frame.data() + 32, n);
    if (nread) *nread = n;
    return Err::nil;
  }
  int64_t shardSize() const { return shardSize_; }

 private:
  ReaderAt* disk_;
  int64_t tillOffset_, shardSize_;
  mec_codec* codec_;
};

// ---- Erasure (cmd/erasure-coding.go:35) ------------------------------------------------------------
class Erasure {
 public:
  static constexpr int64_t kBatchBlocks = 64;  // erasure blocks per GPU call

  // NewErasure (cmd/erasure-coding.go:42)
  static Err New(int dataBlocks, int parityBlocks, int64_t blockSize, std::unique_ptr<Erasure>* out, int device = 0) {
    mec_codec* c = nullptr;
    int rc = mec_codec_new(dataBlocks, parityBlocks, blockSize, MEC_HIGHWAYHASH256S, device, &c);
    if (rc) return fromRc(rc);
    out->reset(new Erasure(c, dataBlocks, parityBlocks, blockSize));
    return Err::nil;
  }
  ~Erasure() { mec_codec_free(codec_); }
  Erasure(const Erasure&) = delete;
  Erasure& operator=(const Erasure&) = delete;

  mec_codec* codec() const { return codec_; }
  int dataBlocks() const { return k_; }
  int parityBlocks() const { return m_; }
  int64_t ShardSize() const { return mec_shard_size(codec_); }                                   // :116
  int64_t ShardFileSize(int64_t total) const { return mec_shard_file_size(codec_, total); }      // :121
  int64_t ShardFileOffset(int64_t start, int64_t length, int64_t total) const {                  // :135
    return mec_shard_file_offset(codec_, start, length, total);
  }

  // EncodeData (:77): Split + Encode.  Empty input => k+m empty shards.
  Err EncodeData(const uint8_t* data, size_t len, std::vector<std::vector<uint8_t>>* shards) const {
    const int n = k_ + m_;
    shards->assign(n, {});
    if (len == 0) return Err::nil;
    const size_t per = (len + k_ - 1) / k_;
    std::vector<uint8_t*> ptr(n);
    for (int i = 0; i < n; i++) { (*shards)[i].assign(per, 0); ptr[i] = (*shards)[i].data(); }
    for (int i = 0; i < k_; i++) {
      const size_t start = static_cast<size_t>(i) * per;
      if (start < len) memcpy(ptr[i], data + start, std::min(per, len - start));
    }
    return fromRc(mec_rs_encode_shards(codec_, ptr.data(), static_cast<int64_t>(per)));
  }
  // DecodeDataBlocks (:94): no-op when nothing / everything is empty
  Err DecodeDataBlocks(std::vector<std::vector<uint8_t>>& data) const {
    int isZero = 0;
    for (auto& b : data) if (b.empty()) { isZero++; break; }
    if (isZero == 0 || isZero == static_cast<int>(data.size())) return Err::nil;
    return reconstruct(data, true);
  }
  // DecodeDataAndParityBlocks (:111)
  Err DecodeDataAndParityBlocks(std::vector<std::vector<uint8_t>>& data) const { return reconstruct(data, false); }

  // Encode (cmd/erasure-encode.go:69): writers[i] == nullptr is an offline disk.
  Err Encode(Reader& src, std::vector<Writer*>& writers, int quorum, int64_t* total) {
    const int n = k_ + m_;
    const int64_t S = ShardSize();
    std::vector<Err> errs(n, Err::nil);
    std::vector<uint8_t> buf(static_cast<size_t>(kBatchBlocks * blockSize_));
    std::vector<uint8_t> parity(static_cast<size_t>(kBatchBlocks * std::max(m_, 1) * S)), dig(static_cast<size_t>(kBatchBlocks * n * 32));
    std::vector<uint8_t> pad(static_cast<size_t>(S));
    *total = 0;
    bool first = true;
    for (;;) {
      // io.ReadFull over a batch of blocks
      size_t got = 0;
      while (got < buf.size()) {
        size_t r = src.Read(buf.data() + got, buf.size() - got);
        if (r == 0) break;
        got += r;
      }
      const bool eof = got < buf.size();
      if (got == 0 && !first) break;                                       // :89-92
      if (got == 0) {                                                      // empty object: one empty write per writer (:94-100)
        Err e = multiWrite(writers, errs, quorum, [&](int, Writer* w) { size_t wr; return w->Write(nullptr, 0, &wr); });
        if (e != Err::nil) return e;
        break;
      }
      first = false;
      int rc = mec_encode_blocks(codec_, buf.data(), static_cast<int64_t>(got), parity.data(), dig.data());
      if (rc) return fromRc(rc);
      const int64_t nb = ceilFrac(static_cast<int64_t>(got), blockSize_);
      for (int64_t b = 0; b < nb; b++) {
        const int64_t blen = std::min<int64_t>(blockSize_, static_cast<int64_t>(got) - b * blockSize_);
        const int64_t per = ceilFrac(blen, k_);
        Err e = multiWrite(writers, errs, quorum, [&](int i, Writer* w) {
          const uint8_t* shard;
          if (i < k_) {
            const int64_t start = static_cast<int64_t>(i) * per, have = std::max<int64_t>(0, std::min(per, blen - start));
            if (have == per) shard = buf.data() + b * blockSize_ + start;
            else {  // Split's zero padding of the last data shard(s)
              if (have > 0) memcpy(pad.data(), buf.data() + b * blockSize_ + start, static_cast<size_t>(have));
              memset(pad.data() + have, 0, static_cast<size_t>(per - have));
              shard = pad.data();
            }
          } else {
            shard = parity.data() + (b * m_ + (i - k_)) * S;
          }
          size_t wr = 0;
          if (auto* bw = dynamic_cast<StreamingBitrotWriter*>(w)) {
            Err we = bw->WriteHashed(dig.data() + (b * n + i) * 32, shard, static_cast<size_t>(per), &wr);
            if (we == Err::nil && wr != static_cast<size_t>(per)) we = Err::ErrShortWrite;
            return we;
          }
          Err we = w->Write(shard, static_cast<size_t>(per), &wr);
          if (we == Err::nil && wr != static_cast<size_t>(per)) we = Err::ErrShortWrite;
          return we;
        });
        if (e != Err::nil) return e;
        *total += blen;
      }
      if (eof) break;
    }
    return Err::nil;
  }

  // Decode (cmd/erasure-decode.go:239).  readers[i] == nullptr is an offline disk.  *derr receives the
  // errFileCorrupt / errFileNotFound heal hint (:288-293) when the read still succeeded.
  Err Decode(Writer& dst, std::vector<StreamingBitrotReader*> readers, int64_t offset, int64_t length, int64_t totalLength,
             int64_t* written, Err* derr = nullptr) {
    if (written) *written = -1;
    if (derr) *derr = Err::nil;
    if (offset < 0 || length < 0) return Err::errInvalidArgument;         // :240-242
    if (offset + length > totalLength) return Err::errInvalidArgument;    // :243-245
    if (length == 0) { if (written) *written = 0; return Err::nil; }      // :247-249
    const int n = k_ + m_;
    const int64_t S = ShardSize(), sfs = ShardFileSize(totalLength);
    const int64_t startBlock = offset / blockSize_, endBlock = (offset + length) / blockSize_;
    int64_t lastBlock = endBlock;
    if ((offset + length) % blockSize_ == 0) lastBlock = endBlock - 1;
    const int64_t nblk = lastBlock - startBlock + 1;
    const int64_t lastLen = std::min(S, sfs - lastBlock * S);
    std::vector<std::vector<uint8_t>> frames(n);
    std::vector<uint8_t> want(n, 0);
    for (int i = 0; i < k_; i++) want[i] = 1;
    std::vector<std::vector<uint8_t>> out(n);
    Err hint = Err::nil;
    Err e = readAndReconstruct(readers, startBlock, nblk, lastLen, want, true, frames, out, &hint);
    if (e != Err::nil) return e;
    int64_t bytesWritten = 0;
    for (int64_t block = startBlock; block <= lastBlock; block++) {       // :261-305
      int64_t bo, bl;
      if (startBlock == endBlock) { bo = offset % blockSize_; bl = length; }
      else if (block == startBlock) { bo = offset % blockSize_; bl = blockSize_ - bo; }
      else if (block == endBlock) { bo = 0; bl = (offset + length) % blockSize_; }
      else { bo = 0; bl = blockSize_; }
      if (bl == 0) break;
      const int64_t cur = block == lastBlock ? lastLen : S;
      if (static_cast<int64_t>(k_) * cur < bl) return Err::ErrShortData;  // writeDataBlocks (erasure-utils.go:54-56)
      int64_t o = bo, w = bl;
      for (int i = 0; i < k_ && w > 0; i++) {
        if (o >= cur) { o -= cur; continue; }
        const int64_t take = std::min(cur - o, w);
        size_t wr = 0;
        Err we = dst.Write(out[i].data() + (block - startBlock) * (32 + S) + 32 + o, static_cast<size_t>(take), &wr);
        if (we != Err::nil) return we;
        bytesWritten += take; w -= take; o = 0;
      }
    }
    if (written) *written = bytesWritten;
    if (bytesWritten != length) return Err::errLessData;                  // :308-310
    if (derr) *derr = hint;
    return Err::nil;
  }

  // Heal (cmd/erasure-decode.go:317): writers[i] != nullptr marks a stale disk to rebuild.
  Err Heal(std::vector<Writer*>& writers, std::vector<StreamingBitrotReader*> readers, int64_t totalLength) {
    const int n = k_ + m_;
    if (static_cast<int>(writers.size()) != n) return Err::errInvalidArgument;   // :318-320
    if (totalLength <= 0) return Err::nil;
    const int64_t S = ShardSize(), sfs = ShardFileSize(totalLength);
    const int64_t nblk = ceilFrac(totalLength, blockSize_), lastLen = sfs - (nblk - 1) * S;
    std::vector<uint8_t> want(n, 0);
    for (int i = 0; i < n; i++) want[i] = writers[i] != nullptr;
    std::vector<std::vector<uint8_t>> frames(n), out(n);
    Err hint = Err::nil;
    Err e = readAndReconstruct(readers, 0, nblk, lastLen, want, false, frames, out, &hint);
    if (e != Err::nil) return e;
    std::vector<Err> errs(n, Err::nil);
    for (int64_t b = 0; b < nblk; b++) {
      const int64_t cur = b == nblk - 1 ? lastLen : S;
      std::vector<Writer*> ws = writers;
      e = multiWrite(ws, errs, 1, [&](int i, Writer* w) {                 // writeQuorum 1 (:352-358)
        const uint8_t* fr = out[i].data() + b * (32 + S);
        size_t wr = 0;
        if (auto* bw = dynamic_cast<StreamingBitrotWriter*>(w)) return bw->WriteHashed(fr, fr + 32, static_cast<size_t>(cur), &wr);
        return w->Write(fr + 32, static_cast<size_t>(cur), &wr);
      }, /*nil_is_error=*/false);
      if (e != Err::nil) return e;
    }
    return hint == Err::nil ? Err::nil : hint;
  }

  // bitrotVerify (cmd/bitrot.go:164) over an in-memory shard file
  Err BitrotVerify(const uint8_t* file, int64_t fileLen, int64_t partLen) const {
    return fromRc(mec_bitrot_verify(codec_, file, fileLen, partLen));
  }

 private:
  Erasure(mec_codec* c, int k, int m, int64_t bs) : codec_(c), k_(k), m_(m), blockSize_(bs) {}

  Err reconstruct(std::vector<std::vector<uint8_t>>& data, bool dataOnly) const {
    const int n = k_ + m_;
    if (static_cast<int>(data.size()) != n) return Err::ErrTooFewShards;
    size_t per = 0;
    for (auto& b : data) {
      if (b.empty()) continue;
      if (per == 0) per = b.size();
      else if (b.size() != per) return Err::ErrShardSize;
    }
    if (per == 0) return Err::ErrShardNoData;
    std::vector<uint8_t> present(n);
    std::vector<std::vector<uint8_t>> tmp(n);
    std::vector<uint8_t*> ptr(n);
    for (int i = 0; i < n; i++) {
      present[i] = !data[i].empty();
      if (!present[i]) { tmp[i].assign(per, 0); ptr[i] = tmp[i].data(); } else ptr[i] = data[i].data();
    }
    int rc = mec_rs_reconstruct_shards(codec_, ptr.data(), present.data(), static_cast<int64_t>(per), dataOnly ? 1 : 0);
    if (rc) return fromRc(rc);
    for (int i = 0; i < n; i++)
      if (!present[i] && !(dataOnly && i >= k_)) data[i] = std::move(tmp[i]);
    return Err::nil;
  }

  // multiWriter.Write (cmd/erasure-encode.go:34-67)
  template <class F>
  Err multiWrite(std::vector<Writer*>& writers, std::vector<Err>& errs, int quorum, F&& writeOne, bool nil_is_error = true) {
    const int n = static_cast<int>(writers.size());
    for (int i = 0; i < n; i++) {
      if (errs[i] != Err::nil) continue;
      if (writers[i] == nullptr) { if (nil_is_error) errs[i] = Err::errDiskNotFound; else errs[i] = Err::errDiskNotFound; continue; }
      errs[i] = writeOne(i, writers[i]);
      if (errs[i] != Err::nil) writers[i] = nullptr;
    }
    int nilCount = 0;
    for (auto e : errs) nilCount += e == Err::nil;
    if (nilCount >= quorum) return Err::nil;
    return Err::errErasureWriteQuorum;
  }

  // parallelReader.Read over a block range + DecodeData(AndParity)Blocks: reads the first k alive readers in
  // index order, hands the frames to the fused kernel, and on a digest mismatch reads further readers.
  Err readAndReconstruct(std::vector<StreamingBitrotReader*>& readers, int64_t firstBlock, int64_t nblk, int64_t lastLen,
                         const std::vector<uint8_t>& want, bool dataOnly, std::vector<std::vector<uint8_t>>& frames,
                         std::vector<std::vector<uint8_t>>& out, Err* hint) {
    const int n = k_ + m_;
    const int64_t S = ShardSize();
    const size_t fbytes = static_cast<size_t>((nblk - 1) * (32 + S) + 32 + lastLen);
    std::vector<uint8_t> alive(n), have(n, 0);
    for (int i = 0; i < n; i++) alive[i] = readers[i] != nullptr;
    for (int i = 0; i < n; i++) if (want[i]) out[i].assign(fbytes, 0);
    for (;;) {
      int cnt = 0;
      for (int i = 0; i < n && cnt < k_; i++) {
        if (!alive[i]) continue;
        if (!have[i]) {
          frames[i].assign(fbytes, 0);
          Err e = Err::nil;
          for (int64_t b = 0; b < nblk && e == Err::nil; b++)
            e = readers[i]->ReadFrameAt(frames[i].data() + b * (32 + S), static_cast<size_t>(b == nblk - 1 ? lastLen : S), (firstBlock + b) * S);
          if (e != Err::nil) {  // errFileNotFound / errFaultyDisk ...: drop the reader, try the next (:193-212)
            alive[i] = 0;
            if (e == Err::errFileNotFound || e == Err::errFileCorrupt) *hint = e;
            continue;
          }
          have[i] = 1;
        }
        cnt++;
      }
      if (cnt < k_) return Err::errErasureReadQuorum;                      // :234
      std::vector<const uint8_t*> fp(n, nullptr);
      std::vector<uint8_t*> op(n, nullptr);
      int used = 0;
      for (int i = 0; i < n; i++) {
        if (alive[i] && have[i] && used < k_) { fp[i] = frames[i].data(); used++; }
        if (want[i]) op[i] = out[i].data();
      }
      std::vector<uint8_t> corrupt(n, 0);
      int rc = mec_reconstruct_frames(codec_, fp.data(), nblk, lastLen == S ? 0 : lastLen, want.data(), dataOnly ? 1 : 0, op.data(), corrupt.data());
      bool any = false;
      for (int i = 0; i < n; i++) if (corrupt[i]) { alive[i] = 0; any = true; *hint = Err::errFileCorrupt; }
      if (rc == 0) return Err::nil;
      if (rc == MEC_ERR_READ_QUORUM && any) continue;  // more readers may be available
      return fromRc(rc);
    }
  }

  mec_codec* codec_;
  int k_, m_;
  int64_t blockSize_;
};

}  // namespace minio
