"""ctypes binding of include/minio_ec.h.  Host buffers are numpy uint8 arrays; device buffers are
raw pointers (e.g. ``torch.Tensor.data_ptr()``)."""
import ctypes as C
import os

import numpy as np

SHA256, HIGHWAYHASH256, HIGHWAYHASH256S, BLAKE2B512 = 1, 2, 3, 4
ERRORS = {-1: "ErrInvShardNum", -2: "ErrMaxShardNum", -3: "ErrTooFewShards", -4: "ErrShardNoData", -5: "ErrShardSize",
          -6: "ErrShortData", -7: "errFileCorrupt", -8: "errLessData", -9: "errUnexpected", -10: "errErasureReadQuorum",
          -11: "errErasureWriteQuorum", -12: "errInvalidArgument", -100: "CUDA error", -101: "no CUDA device",
          -102: "unsupported on the GPU path"}

_LIB = None


class MecError(RuntimeError):
    def __init__(self, code, what=""):
        self.code = code
        detail = lib().mec_last_error().decode() if code <= -100 else ""
        super().__init__(f"{what}: {ERRORS.get(code, code)} ({code}) {detail}")


def lib_path():
    # MEC_LIB selects an experimental build variant (tools/build_variants.sh); default is the product library
    return os.environ.get("MEC_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libminio_ec.so")


def lib():
    """Load libminio_ec.so.  Fails loudly when the CUDA extension has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    L = C.CDLL(path)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.mec_codec_new.argtypes = [i32, i32, i64, i32, i32, C.POINTER(vp)]
    L.mec_codec_free.argtypes = [vp]
    L.mec_codec_free.restype = None
    L.mec_last_error.restype = C.c_char_p
    L.mec_version.restype = C.c_char_p
    for f, args in (("mec_shard_size", [vp]), ("mec_shard_file_size", [vp, i64]),
                    ("mec_shard_file_offset", [vp, i64, i64, i64]), ("mec_bitrot_shard_file_size", [i64, i64, i32]),
                    ("mec_ceil_frac", [i64, i64]), ("mec_launch_count", [vp])):
        getattr(L, f).restype = i64
        getattr(L, f).argtypes = args
    L.mec_alloc_pinned.restype = vp
    L.mec_alloc_pinned.argtypes = [C.c_size_t]
    L.mec_free_pinned.argtypes = [vp]
    L.mec_free_pinned.restype = None
    L.mec_alloc_pinned_on.restype = vp
    L.mec_alloc_pinned_on.argtypes = [i32, C.c_size_t]
    L.mec_device_numa_node.argtypes = [i32]
    L.mec_bind_thread_to_device.argtypes = [i32]
    L.mec_is_pinned.argtypes = [vp]
    L.mec_encode_sg.restype = i64
    L.mec_encode_sg.argtypes = [vp, vp, i64, vp, vp, i32]
    L.mec_heal_prefer.argtypes = [vp, vp, vp, i64, vp, vp]
    L.mec_batcher_new.argtypes = [i32, i32, i64, i32, i64, i32, C.POINTER(vp)]
    L.mec_batcher_free.argtypes = [vp]
    L.mec_batcher_free.restype = None
    L.mec_batcher_encode.restype = i64
    L.mec_batcher_encode.argtypes = [vp, vp, i64, vp, i32]
    L.mec_batcher_encode_sg.restype = i64
    L.mec_batcher_encode_sg.argtypes = [vp, vp, i64, vp, vp, i32]
    L.mec_batcher_decode.restype = i64
    L.mec_batcher_decode.argtypes = [vp, vp, i64, i64, i64, vp, C.POINTER(i32)]
    L.mec_batcher_stat.restype = i64
    L.mec_batcher_stat.argtypes = [vp, C.c_char_p]
    L.mec_checksums.argtypes = [vp, vp, i64, i32, vp]
    L.mec_checksums_device.argtypes = [vp, vp, i64, i32, vp, vp]
    L.mec_checksum_combine.restype = C.c_uint64
    L.mec_checksum_combine.argtypes = [i32, C.c_uint64, C.c_uint64, i64]
    L.mec_last_checksums.restype = i64
    L.mec_last_checksums.argtypes = [vp, vp]
    L.mec_decode_whole.restype = i64
    L.mec_decode_whole.argtypes = [vp, vp, vp, i64, i64, i64, vp, C.POINTER(i32)]
    L.mec_heal_whole.argtypes = [vp, vp, vp, i64, vp, vp, vp]
    L.mec_encode_blocks.argtypes = [vp, vp, i64, vp, vp]
    L.mec_encode_blocks_device.argtypes = [vp, vp, i64, vp, i64, vp, vp]
    L.mec_reconstruct_frames.argtypes = [vp, vp, i64, i64, vp, i32, vp, vp]
    L.mec_reconstruct_device.argtypes = [vp, vp, i64, i64, vp, i32, vp, i64, vp, vp, vp]
    L.mec_encode.restype = i64
    L.mec_encode.argtypes = [vp, vp, i64, vp, i32]
    L.mec_decode.restype = i64
    L.mec_decode.argtypes = [vp, vp, i64, i64, i64, vp, C.POINTER(i32)]
    L.mec_heal.argtypes = [vp, vp, i64, vp]
    L.mec_heal_batch.argtypes = [vp, i32, i64, vp, vp, vp, vp]
    L.mec_decode_prefer.restype = i64
    L.mec_decode_prefer.argtypes = [vp, vp, vp, i64, i64, i64, vp, C.POINTER(i32)]
    L.mec_jit_compile_check.restype = i64
    L.mec_jit_compile_check.argtypes = [i32, i32, vp, i32, i32, i32, i32]
    L.mec_jit_prewarm.argtypes = [vp]
    L.mec_shutdown.restype = None
    L.mec_shutdown.argtypes = []
    import atexit
    atexit.register(L.mec_shutdown)  # no NVRTC compile may be in flight when the C runtime starts its exit handlers
    L.mec_get_stat.restype = i64
    L.mec_get_stat.argtypes = [vp, C.c_char_p]
    L.mec_bitrot_verify.argtypes = [vp, vp, i64, i64]
    L.mec_bitrot_verify_batch.argtypes = [vp, i64, vp, vp, vp, vp]
    L.mec_encode_whole.restype = i64
    L.mec_encode_whole.argtypes = [vp, vp, i64, vp, vp, i32]
    L.mec_whole_hash.argtypes = [vp, i32, vp, i64, i64, vp]
    L.mec_bitrot_verify_whole.argtypes = [vp, i32, vp, i64, vp]
    L.mec_digest_size.argtypes = [i32]
    L.mec_whole_hash_device.argtypes = [vp, i32, vp, i64, i64, i64, vp, vp]
    L.mec_rs_encode_shards.argtypes = [vp, vp, i64]
    L.mec_rs_reconstruct_shards.argtypes = [vp, vp, vp, i64, i32]
    L.mec_hh256_batch.argtypes = [vp, vp, i64, i64, vp]
    L.mec_selftest.argtypes = [i32]
    L.mec_set_option.argtypes = [vp, C.c_char_p, i64]
    _LIB = L
    return L


def device_count():
    return lib().mec_device_count()


def selftest(device=0):
    rc = lib().mec_selftest(device)
    if rc:
        raise MecError(rc, "mec_selftest")


def _u8(a):
    if isinstance(a, (bytes, bytearray, memoryview)):
        a = np.frombuffer(a, dtype=np.uint8)
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def _ptrs(arrs):
    return (C.c_void_p * len(arrs))(*[(a.ctypes.data if (a is not None and a.size) else None) for a in arrs])


class Codec:
    """NewErasure (cmd/erasure-coding.go:42) bound to one CUDA device."""

    def __init__(self, k, m, block_size=1 << 20, algo=HIGHWAYHASH256S, device=0):
        h = C.c_void_p()
        rc = lib().mec_codec_new(k, m, block_size, algo, device, C.byref(h))
        if rc:
            raise MecError(rc, "mec_codec_new")
        self.h, self.k, self.m, self.n, self.block_size, self.algo = h, k, m, k + m, block_size, algo

    def close(self):
        if getattr(self, "h", None):
            lib().mec_codec_free(self.h)
            self.h = None

    __del__ = close

    # -- helpers
    def set_option(self, name, value):
        rc = lib().mec_set_option(self.h, name.encode(), int(value))
        if rc:
            raise MecError(rc, "mec_set_option")

    @property
    def launches(self):
        return lib().mec_launch_count(self.h)

    def shard_size(self):
        return lib().mec_shard_size(self.h)

    def shard_file_size(self, total):
        return lib().mec_shard_file_size(self.h, total)

    def shard_file_offset(self, start, length, total):
        return lib().mec_shard_file_offset(self.h, start, length, total)

    def bitrot_file_size(self, total):
        return lib().mec_bitrot_shard_file_size(self.shard_file_size(total), self.shard_size(), self.algo)

    # -- block level
    def encode_blocks(self, src):
        """-> (parity [nblocks, m, S], digests [nblocks, n, 32]); the last block may use < S bytes per row."""
        src = _u8(src)
        nb = -(-src.size // self.block_size)
        S = self.shard_size()
        parity = np.zeros((nb, max(self.m, 1), S), dtype=np.uint8)
        dig = np.zeros((nb, self.n, 32), dtype=np.uint8)
        rc = lib().mec_encode_blocks(self.h, src.ctypes.data if src.size else None, src.size, parity.ctypes.data, dig.ctypes.data)
        if rc:
            raise MecError(rc, "mec_encode_blocks")
        return parity[:, :self.m], dig

    def encode_blocks_device(self, d_src, length, d_parity, parity_pitch, d_digests, stream=0):
        rc = lib().mec_encode_blocks_device(self.h, d_src, length, d_parity, parity_pitch, d_digests, stream)
        if rc:
            raise MecError(rc, "mec_encode_blocks_device")

    def reconstruct_device(self, d_frames, frame_pitch, nblocks, want, data_only, d_out, out_pitch, d_digests, d_corrupt, stream=0):
        """d_frames: list of k+m device pointers (0/None = unavailable).  data_only: bool, or the MEC_RECONSTRUCT_* flag bits
        (1 = data only, 2 = no digests for the rebuilt shards)."""
        arr = (C.c_void_p * self.n)(*[(p or None) for p in d_frames])
        want = np.asarray(want, dtype=np.uint8)
        rc = lib().mec_reconstruct_device(self.h, arr, frame_pitch, nblocks, want.ctypes.data, int(data_only),
                                          d_out, out_pitch, d_digests, d_corrupt, stream)
        if rc:
            raise MecError(rc, "mec_reconstruct_device")

    # -- whole part
    def encode(self, src, online=None, write_quorum=0):
        src = _u8(src)
        fsz = self.bitrot_file_size(src.size)
        online = [True] * self.n if online is None else online
        files = [np.zeros(fsz, dtype=np.uint8) if online[i] else None for i in range(self.n)]
        rc = lib().mec_encode(self.h, src.ctypes.data if src.size else None, src.size, _ptrs(files), write_quorum)
        if rc < 0:
            raise MecError(rc, "mec_encode")
        return files

    def encode_sg(self, src, online=None, write_quorum=0):
        """mec_encode_sg: -> (files, data_digests).  files[i] is None for data drives (their frames are the caller's own
        source slices behind data_digests[b, i]) and the complete part.N image for online parity drives."""
        src = _u8(src)
        fsz = self.bitrot_file_size(src.size)
        nb = -(-src.size // self.block_size)
        online = [True] * self.n if online is None else online
        marker = np.zeros(1, dtype=np.uint8)  # data drives: only NULL / non-NULL is looked at
        files = [(np.zeros(fsz, dtype=np.uint8) if i >= self.k else marker) if online[i] else None for i in range(self.n)]
        dd = np.zeros((max(nb, 1), self.k, 32), dtype=np.uint8)
        rc = lib().mec_encode_sg(self.h, src.ctypes.data if src.size else None, src.size, _ptrs(files), dd.ctypes.data, write_quorum)
        if rc < 0:
            raise MecError(rc, "mec_encode_sg")
        return [f if i >= self.k else None for i, f in enumerate(files)], dd[:nb]

    def checksums(self, data, which=7):
        """-> (CRC32, CRC32C, CRC64NVME) of a host byte string (mec_checksums); 0 where not asked for."""
        d = _u8(data)
        out = (C.c_uint64 * 3)()
        rc = lib().mec_checksums(self.h, d.ctypes.data if d.size else None, d.size, which, out)
        if rc:
            raise MecError(rc, "mec_checksums")
        return tuple(int(v) for v in out)

    def checksums_device(self, d_src, length, which=7, stream=0):
        out = (C.c_uint64 * 3)()
        rc = lib().mec_checksums_device(self.h, d_src, length, which, out, stream)
        if rc:
            raise MecError(rc, "mec_checksums_device")
        return tuple(int(v) for v in out)

    def last_checksums(self):
        out = (C.c_uint64 * 3)()
        n = lib().mec_last_checksums(self.h, out)
        return tuple(int(v) for v in out), n

    def stat(self, name):
        return lib().mec_get_stat(self.h, name.encode())

    def decode(self, files, offset, length, total, prefer=None):
        files = [None if f is None else _u8(f) for f in files]
        dst = np.zeros(max(length, 1), dtype=np.uint8)
        hint = C.c_int(0)
        pf = None if prefer is None else np.asarray(prefer, dtype=np.uint8)
        rc = lib().mec_decode_prefer(self.h, _ptrs(files), None if pf is None else pf.ctypes.data, offset, length, total,
                                     dst.ctypes.data, C.byref(hint))
        if rc < 0:
            raise MecError(rc, "mec_decode")
        return dst[:length], hint.value

    def heal(self, files, stale, total, prefer=None, report=False):
        """Erasure.Heal.  Like the reference, bitrot met in a source reader does not stop the rebuild but is reported:
        report=False raises MecError(errFileCorrupt) after the healed files were produced (Heal's derr); report=True
        returns (outs, rc, corrupt[n])."""
        files = [None if f is None else _u8(f) for f in files]
        fsz = self.bitrot_file_size(total)
        outs = [np.zeros(fsz, dtype=np.uint8) if stale[i] else None for i in range(self.n)]
        corrupt = np.zeros(self.n, dtype=np.uint8)
        pf = None if prefer is None else np.asarray(prefer, dtype=np.uint8)
        rc = lib().mec_heal_prefer(self.h, _ptrs(files), None if pf is None else pf.ctypes.data, total, _ptrs(outs), corrupt.ctypes.data)
        if report:
            if rc and rc != -7:
                raise MecError(rc, "mec_heal")
            return outs, rc, corrupt
        if rc:
            raise MecError(rc, "mec_heal")
        return outs

    def reconstruct_frames(self, frames, nblocks, last_shard_len, want, data_only=False):
        frames = [None if f is None else _u8(f) for f in frames]
        S = self.shard_size()
        last = last_shard_len or S
        fbytes = (nblocks - 1) * (32 + S) + 32 + last if nblocks else 0
        want = np.asarray(want, dtype=np.uint8)
        outs = [np.zeros(fbytes, dtype=np.uint8) if want[i] else None for i in range(self.n)]
        corrupt = np.zeros(self.n, dtype=np.uint8)
        rc = lib().mec_reconstruct_frames(self.h, _ptrs(frames), nblocks, last_shard_len, want.ctypes.data,
                                          1 if data_only else 0, _ptrs(outs), corrupt.ctypes.data)
        if rc:
            raise MecError(rc, "mec_reconstruct_frames")
        return outs, corrupt

    def bitrot_verify(self, file, part_len):
        f = _u8(file)
        return lib().mec_bitrot_verify(self.h, f.ctypes.data if f.size else None, f.size, part_len)

    def bitrot_verify_batch(self, files, part_lens):
        """mec_bitrot_verify_batch: -> list of 0 / -7 per shard file."""
        files = [_u8(f) for f in files]
        n = len(files)
        fl = (C.c_int64 * n)(*[f.size for f in files])
        pl = (C.c_int64 * n)(*[int(x) for x in part_lens])
        res = (C.c_int32 * n)()
        rc = lib().mec_bitrot_verify_batch(self.h, n, _ptrs(files), fl, pl, res)
        if rc:
            raise MecError(rc, "mec_bitrot_verify_batch")
        return list(res)

    # -- legacy whole-file bitrot
    def encode_whole(self, src, online=None, write_quorum=0):
        src = _u8(src)
        flen = self.shard_file_size(src.size)
        online = [True] * self.n if online is None else online
        files = [np.zeros(flen, dtype=np.uint8) if online[i] else None for i in range(self.n)]
        sums = np.zeros((self.n, 64), dtype=np.uint8)
        rc = lib().mec_encode_whole(self.h, src.ctypes.data if src.size else None, src.size, _ptrs(files), sums.ctypes.data, write_quorum)
        if rc < 0:
            raise MecError(rc, "mec_encode_whole")
        ds = lib().mec_digest_size(self.algo)
        return files, [sums[i, :ds].tobytes() for i in range(self.n)]

    def decode_whole(self, files, sums, offset, length, total):
        """Erasure.Decode over wholeBitrotReaders: files = raw shard files (None = offline), sums = their expected digests."""
        files = [None if f is None else _u8(f) for f in files]
        sm = np.zeros((self.n, 64), dtype=np.uint8)
        for i, d in enumerate(sums):
            if d is not None:
                sm[i, :len(d)] = np.frombuffer(bytes(d), dtype=np.uint8)
        dst = np.zeros(max(length, 1), dtype=np.uint8)
        hint = C.c_int(0)
        rc = lib().mec_decode_whole(self.h, _ptrs(files), sm.ctypes.data, offset, length, total, dst.ctypes.data, C.byref(hint))
        if rc < 0:
            raise MecError(rc, "mec_decode_whole")
        return dst[:length], hint.value

    def heal_whole(self, files, sums, stale, total):
        """-> (out_files, out_sums, rc, corrupt): rc is 0 or -7 (healed, bitrot met in a source)."""
        files = [None if f is None else _u8(f) for f in files]
        sm = np.zeros((self.n, 64), dtype=np.uint8)
        for i, d in enumerate(sums):
            if d is not None:
                sm[i, :len(d)] = np.frombuffer(bytes(d), dtype=np.uint8)
        flen = self.shard_file_size(total)
        outs = [np.zeros(flen, dtype=np.uint8) if stale[i] else None for i in range(self.n)]
        osum = np.zeros((self.n, 64), dtype=np.uint8)
        corrupt = np.zeros(self.n, dtype=np.uint8)
        rc = lib().mec_heal_whole(self.h, _ptrs(files), sm.ctypes.data, total, _ptrs(outs), osum.ctypes.data, corrupt.ctypes.data)
        if rc and rc != -7:
            raise MecError(rc, "mec_heal_whole")
        ds = lib().mec_digest_size(self.algo)
        return outs, [osum[i, :ds].tobytes() if stale[i] else None for i in range(self.n)], rc, corrupt

    def whole_hash(self, algo, msgs, msg_len, count):
        msgs = _u8(msgs)
        ds = lib().mec_digest_size(algo)
        out = np.zeros((count, ds), dtype=np.uint8)
        rc = lib().mec_whole_hash(self.h, algo, msgs.ctypes.data if msgs.size else None, msg_len, count, out.ctypes.data)
        if rc:
            raise MecError(rc, "mec_whole_hash")
        return out

    def whole_hash_device(self, algo, d_msgs, pitch, msg_len, count, d_digests, stream=0):
        rc = lib().mec_whole_hash_device(self.h, algo, d_msgs, pitch, msg_len, count, d_digests, stream)
        if rc:
            raise MecError(rc, "mec_whole_hash_device")

    def bitrot_verify_whole(self, algo, file, want):
        f = _u8(file); w = _u8(want)
        return lib().mec_bitrot_verify_whole(self.h, algo, f.ctypes.data if f.size else None, f.size, w.ctypes.data)

    # -- shard shaped
    def rs_encode_shards(self, shards):
        """reedsolomon.Encoder.Encode: list of k+m equally sized arrays, parity written in place."""
        per = shards[0].size
        rc = lib().mec_rs_encode_shards(self.h, _ptrs(shards), per)
        if rc:
            raise MecError(rc, "mec_rs_encode_shards")

    def rs_reconstruct_shards(self, shards, present, data_only=False):
        per = max(s.size for s in shards)
        pres = np.asarray(present, dtype=np.uint8)
        return lib().mec_rs_reconstruct_shards(self.h, _ptrs(shards), pres.ctypes.data, per, 1 if data_only else 0)

    def hh256_batch(self, msgs, msg_len, count):
        msgs = _u8(msgs)
        out = np.zeros((count, 32), dtype=np.uint8)
        rc = lib().mec_hh256_batch(self.h, msgs.ctypes.data if msgs.size else None, msg_len, count, out.ctypes.data)
        if rc:
            raise MecError(rc, "mec_hh256_batch")
        return out

    def encode_data(self, data):
        """Erasure.EncodeData (cmd/erasure-coding.go:77): Split + Encode -> k+m shard arrays."""
        data = _u8(data)
        if data.size == 0:
            return [np.zeros(0, dtype=np.uint8) for _ in range(self.n)]
        per = -(-data.size // self.k)
        store = np.zeros(self.n * per, dtype=np.uint8)
        store[:data.size] = data
        shards = [store[i * per:(i + 1) * per] for i in range(self.n)]
        if self.m:
            self.rs_encode_shards(shards)
        return shards


class Batcher:
    """mec_batcher: concurrent PutObject calls of one geometry merged into shared launches.  encode() may be called from many
    threads at once (ctypes releases the GIL for the duration of the call)."""

    def __init__(self, k, m, block_size=1 << 20, device=0, max_batch_blocks=256, max_wait_us=200):
        h = C.c_void_p()
        rc = lib().mec_batcher_new(k, m, block_size, device, max_batch_blocks, max_wait_us, C.byref(h))
        if rc:
            raise MecError(rc, "mec_batcher_new")
        self.h, self.k, self.m, self.n, self.block_size = h, k, m, k + m, block_size
        self._sizes = Codec(k, m, block_size, device=device)

    def encode(self, src, online=None, write_quorum=0, pinned=False):
        """pinned=True: the part files are allocated page-locked (and `src` must be, e.g. a pinned_array) — the batch then moves
        through the gather / scatter kernels instead of per-request copies.  The caller owns (and should free) those arrays."""
        src = _u8(src)
        fsz = self._sizes.bitrot_file_size(src.size)
        online = [True] * self.n if online is None else online
        mk = (lambda nb: pinned_array(max(nb, 1))[:nb]) if pinned else (lambda nb: np.zeros(nb, dtype=np.uint8))
        files = [mk(fsz) if online[i] else None for i in range(self.n)]
        rc = lib().mec_batcher_encode(self.h, src.ctypes.data if src.size else None, src.size, _ptrs(files), write_quorum)
        if rc < 0:
            raise MecError(rc, "mec_batcher_encode")
        return files

    def decode(self, files, offset, length, total, dst=None):
        """mec_batcher_decode: files = part-file arrays (None = offline); pass page-locked arrays (and a page-locked dst) to ride in
        merged launches.  -> (bytes, heal_hint)"""
        files = [None if f is None else _u8(f) for f in files]
        if dst is None:
            dst = np.zeros(max(length, 1), dtype=np.uint8)
        hint = C.c_int(0)
        rc = lib().mec_batcher_decode(self.h, _ptrs(files), offset, length, total, dst.ctypes.data, C.byref(hint))
        if rc < 0:
            raise MecError(rc, "mec_batcher_decode")
        return dst[:length], hint.value

    def stat(self, name):
        return lib().mec_batcher_stat(self.h, name.encode())

    def close(self):
        if getattr(self, "h", None):
            lib().mec_batcher_free(self.h)
            self.h = None
            self._sizes.close()

    __del__ = close


def pinned_array(nbytes, device=None):
    """uint8 numpy array over page-locked host memory from mec_alloc_pinned (bpool.BytePoolCap's role) — with `device`, on the
    NUMA node of that GPU (mec_alloc_pinned_on); never freed by the array — keep it for the life of the process or release
    it with lib().mec_free_pinned(arr.ctypes.data)."""
    p = lib().mec_alloc_pinned(max(int(nbytes), 1)) if device is None else lib().mec_alloc_pinned_on(int(device), max(int(nbytes), 1))
    if not p:
        raise MecError(-100, "mec_alloc_pinned")
    return np.ctypeslib.as_array((C.c_uint8 * int(nbytes)).from_address(p))


def heal_batch(pool, objects, outs=None):
    """mec_heal_batch: `pool` = list of Codec handles of one geometry, `objects` = list of (files, stale, total).
    Returns the rebuilt shard files per object (None where not stale).  `outs` = preallocated output arrays per object
    (e.g. pinned_array of bitrot_file_size(total) bytes where stale, None elsewhere); default: fresh pageable arrays."""
    n = pool[0].n
    handles = (C.c_void_p * len(pool))(*[c.h for c in pool])
    keep, fptrs, optrs, outs_all = [], [], [], []
    for o, (files, stale, total) in enumerate(objects):
        files = [None if f is None else _u8(f) for f in files]
        fsz = pool[0].bitrot_file_size(total)
        obj_outs = outs[o] if outs is not None else [np.zeros(fsz, dtype=np.uint8) if stale[i] else None for i in range(n)]
        fp, op = _ptrs(files), _ptrs(obj_outs)
        keep += [files, fp, op]
        fptrs.append(C.cast(fp, C.c_void_p)); optrs.append(C.cast(op, C.c_void_p)); outs_all.append(obj_outs)
    nobj = len(objects)
    fa = (C.c_void_p * nobj)(*fptrs); oa = (C.c_void_p * nobj)(*optrs)
    totals = (C.c_int64 * nobj)(*[t for _, _, t in objects])
    rcs = (C.c_int32 * nobj)()
    rc = lib().mec_heal_batch(handles, len(pool), nobj, fa, totals, oa, rcs)
    if rc:
        raise MecError(rc, "mec_heal_batch")
    return outs_all
