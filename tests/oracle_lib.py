"""ctypes binding of the CPU oracle (oracle/liboracle.so) — test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None

SHA256, HIGHWAYHASH256, HIGHWAYHASH256S, BLAKE2B512 = 1, 2, 3, 4
ERR = dict(INV_SHARD_NUM=-1, MAX_SHARD_NUM=-2, TOO_FEW_SHARDS=-3, SHARD_NO_DATA=-4, SHARD_SIZE=-5,
           SHORT_DATA=-6, FILE_CORRUPT=-7, LESS_DATA=-8, UNEXPECTED=-9, READ_QUORUM=-10,
           WRITE_QUORUM=-11, INVALID_ARGUMENT=-12)

u8p = C.POINTER(C.c_uint8)


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.orc_gf_mul.restype = C.c_uint8
    L.orc_gf_mul.argtypes = [C.c_uint8, C.c_uint8]
    L.orc_gf_inv.restype = C.c_uint8
    L.orc_gf_inv.argtypes = [C.c_uint8]
    L.orc_rs_matrix.argtypes = [C.c_int, C.c_int, C.c_void_p]
    L.orc_rs_encode.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64]
    L.orc_rs_reconstruct.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
    L.orc_rs_decode_rows.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_rs_split.restype = C.c_int64
    L.orc_rs_split.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
    L.orc_hh256.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.orc_hh256_fast.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.orc_sha256.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    L.orc_blake2b512.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    L.orc_xxh64.restype = C.c_uint64
    L.orc_xxh64.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
    L.orc_bitrot_hash.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    for f in ("orc_shard_size", "orc_shard_file_size", "orc_shard_file_offset", "orc_bitrot_shard_file_size", "orc_ceil_frac"):
        getattr(L, f).restype = C.c_int64
    L.orc_ceil_frac.argtypes = [C.c_int64, C.c_int64]
    L.orc_shard_size.argtypes = [C.c_int64, C.c_int]
    L.orc_shard_file_size.argtypes = [C.c_int64, C.c_int, C.c_int64]
    L.orc_shard_file_offset.argtypes = [C.c_int64, C.c_int, C.c_int64, C.c_int64, C.c_int64]
    L.orc_bitrot_shard_file_size.argtypes = [C.c_int64, C.c_int64, C.c_int]
    L.orc_erasure_encode.restype = C.c_int64
    L.orc_erasure_encode.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.orc_bitrot_verify.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
    L.orc_erasure_decode.restype = C.c_int64
    L.orc_erasure_decode.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    L.orc_erasure_heal.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.orc_simd_level.restype = C.c_char_p
    L.orc_rs_encode_fast.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64]
    L.orc_rs_encode_fast.restype = None
    L.orc_encode_hash_blocks_mt.restype = C.c_double
    L.orc_encode_hash_blocks_mt.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    for f in ("orc_crc32", "orc_crc32c", "orc_crc64nvme"):
        getattr(L, f).restype = C.c_uint64
        getattr(L, f).argtypes = [C.c_void_p, C.c_size_t]
    L.orc_pool_new.restype = C.c_void_p
    L.orc_pool_new.argtypes = [C.c_int]
    L.orc_pool_free.argtypes = [C.c_void_p]
    L.orc_pool_free.restype = None
    L.orc_pool_fill.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_uint64]
    L.orc_pool_fill.restype = None
    L.orc_pool_encode_hash.restype = C.c_double
    L.orc_pool_encode_hash.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
    _LIB = L
    return L


MAGIC_KEY = bytes.fromhex("4be734fa8e238acd263e83e6bb968552040f935da39f441497e09d1322de36a0")


def _buf(b):
    """bytes/ndarray -> (ctypes pointer-able object, keepalive)"""
    a = np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b
    a = np.ascontiguousarray(a)
    return a


def _ptr(a):
    return a.ctypes.data if a.size else None


def _ptr_array(arrs):
    arr = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    return arr


def rs_matrix(k, m):
    out = np.zeros(((k + m), k), dtype=np.uint8)
    rc = lib().orc_rs_matrix(k, m, out.ctypes.data)
    if rc:
        raise ValueError(rc)
    return out


def hh256(msg, key=MAGIC_KEY, fast=False):
    a = _buf(msg)
    k = _buf(key)
    out = np.zeros(32, dtype=np.uint8)
    (lib().orc_hh256_fast if fast else lib().orc_hh256)(k.ctypes.data, _ptr(a), a.size, out.ctypes.data)
    return out.tobytes()


def sha256(msg):
    a = _buf(msg); out = np.zeros(32, dtype=np.uint8)
    lib().orc_sha256(_ptr(a), a.size, out.ctypes.data)
    return out.tobytes()


def blake2b512(msg):
    a = _buf(msg); out = np.zeros(64, dtype=np.uint8)
    lib().orc_blake2b512(_ptr(a), a.size, out.ctypes.data)
    return out.tobytes()


def xxh64(msg, seed=0):
    a = _buf(msg)
    return lib().orc_xxh64(_ptr(a), a.size, seed)


def bitrot_hash(algo, msg):
    a = _buf(msg); out = np.zeros(64, dtype=np.uint8)
    n = lib().orc_bitrot_hash(algo, _ptr(a), a.size, out.ctypes.data)
    return out[:n].tobytes()


def shard_size(bs, k):
    return lib().orc_shard_size(bs, k)


def shard_file_size(bs, k, total):
    return lib().orc_shard_file_size(bs, k, total)


def shard_file_offset(bs, k, start, length, total):
    return lib().orc_shard_file_offset(bs, k, start, length, total)


def bitrot_shard_file_size(size, ss, algo):
    return lib().orc_bitrot_shard_file_size(size, ss, algo)


def encode_data(k, m, data, fast=False):
    """Erasure.EncodeData (erasure-coding.go:77): returns list of k+m shard arrays."""
    data = _buf(data)
    if data.size == 0:
        return [np.zeros(0, dtype=np.uint8) for _ in range(k + m)]
    per = -(-data.size // k)
    store = np.zeros((k + m) * per, dtype=np.uint8)
    got = lib().orc_rs_split(k, m, data.ctypes.data, data.size, store.ctypes.data)
    assert got == per
    shards = [store[i * per:(i + 1) * per] for i in range(k + m)]
    if m:
        pa = _ptr_array(shards)
        if fast:
            lib().orc_rs_encode_fast(k, m, pa, per)
        else:
            rc = lib().orc_rs_encode(k, m, pa, per)
            if rc:
                raise ValueError(rc)
    return shards


def reconstruct(k, m, shards, data_only=False):
    """shards: list of arrays, empty/None = missing. Fills in place; returns rc."""
    n = k + m
    if len(shards) != n:
        return ERR["TOO_FEW_SHARDS"]
    sizes = {len(s) for s in shards if s is not None and len(s)}
    if not sizes:
        return ERR["SHARD_NO_DATA"]
    if len(sizes) > 1:
        return ERR["SHARD_SIZE"]
    per = sizes.pop()
    present = np.array([1 if (s is not None and len(s)) else 0 for s in shards], dtype=np.uint8)
    bufs = [np.ascontiguousarray(s) if present[i] else np.zeros(per, dtype=np.uint8) for i, s in enumerate(shards)]
    rc = lib().orc_rs_reconstruct(k, m, _ptr_array(bufs), present.ctypes.data, per, 1 if data_only else 0)
    if rc == 0:
        for i in range(n):
            if not present[i] and (not data_only or i < k):
                shards[i] = bufs[i]
    return rc


def decode_rows(k, m, present, missing):
    present = np.asarray(present, dtype=np.uint8)
    miss = np.asarray(missing, dtype=np.int32)
    rows = np.zeros((len(miss), k), dtype=np.uint8)
    valid = np.zeros(k, dtype=np.int32)
    rc = lib().orc_rs_decode_rows(k, m, present.ctypes.data, miss.ctypes.data, len(miss), rows.ctypes.data, valid.ctypes.data)
    if rc:
        raise ValueError(rc)
    return rows, valid


def erasure_encode(k, m, bs, algo, data):
    """Erasure.Encode + bitrot writers: returns (list of shard-file arrays, list of whole-file sums or None)."""
    data = _buf(data)
    n = k + m
    ss = shard_size(bs, k)
    fsz = bitrot_shard_file_size(shard_file_size(bs, k, data.size), ss, algo)
    files = [np.zeros(fsz, dtype=np.uint8) for _ in range(n)]
    sums = np.zeros(n * 64, dtype=np.uint8)
    pa = (C.c_void_p * n)(*[f.ctypes.data if f.size else 0 for f in files])
    rc = lib().orc_erasure_encode(k, m, bs, algo, _ptr(data), data.size, pa, sums.ctypes.data)
    if rc < 0:
        raise ValueError(rc)
    ds = 64 if algo == BLAKE2B512 else 32
    ssums = None if algo == HIGHWAYHASH256S else [sums[i * 64:i * 64 + ds].tobytes() for i in range(n)]
    return files, ssums


def bitrot_verify(algo, file, part_len, ss, want=None):
    f = _buf(file)
    w = _buf(want) if want is not None else np.zeros(64, dtype=np.uint8)
    return lib().orc_bitrot_verify(algo, _ptr(f), f.size, part_len, ss, w.ctypes.data)


def erasure_decode(k, m, bs, algo, files, avail, offset, length, total):
    n = k + m
    files = [_buf(f) for f in files]
    pa = (C.c_void_p * n)(*[f.ctypes.data if f.size else 0 for f in files])
    av = np.asarray(avail, dtype=np.uint8)
    dst = np.zeros(max(length, 1), dtype=np.uint8)
    corrupt = np.zeros(n, dtype=np.uint8)
    rc = lib().orc_erasure_decode(k, m, bs, algo, pa, av.ctypes.data, offset, length, total, dst.ctypes.data, corrupt.ctypes.data)
    return rc, dst[:max(length, 0)], corrupt


def erasure_heal(k, m, bs, algo, files, avail, stale, total):
    n = k + m
    files = [_buf(f) for f in files]
    fsz = bitrot_shard_file_size(shard_file_size(bs, k, total), shard_size(bs, k), algo)
    outs = [np.zeros(fsz, dtype=np.uint8) for _ in range(n)]
    pa = (C.c_void_p * n)(*[f.ctypes.data if f.size else 0 for f in files])
    po = (C.c_void_p * n)(*[f.ctypes.data if f.size else 0 for f in outs])
    av = np.asarray(avail, dtype=np.uint8)
    st = np.asarray(stale, dtype=np.uint8)
    rc = lib().orc_erasure_heal(k, m, bs, algo, pa, av.ctypes.data, st.ctypes.data, total, po)
    return rc, outs


def crcs(data):
    """(CRC32, CRC32C, CRC64NVME) of a byte string, as Go's hash/crc32 / hash/crc64 return them."""
    a = _buf(data)
    return (lib().orc_crc32(_ptr(a), a.size), lib().orc_crc32c(_ptr(a), a.size), lib().orc_crc64nvme(_ptr(a), a.size))
