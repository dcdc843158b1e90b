"""Pins the CPU oracle against every golden the reference tree holds for this path."""
import hashlib
import json
import os

import numpy as np
import pytest

import selftest_goldens as G

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def selftest_digest(o, shards):
    buf = b"".join(bytes([i]) + s.tobytes() for i, s in enumerate(shards))
    return o.xxh64(buf)


@pytest.mark.parametrize("km", sorted(G.ERASURE_SELFTEST))
def test_erasure_selftest_golden(oracle, km):
    """cmd/erasure-coding.go:149-205 replayed: encode bytes 0..255, digest, drop shard 0, rebuild."""
    k, m = km
    shards = oracle.encode_data(k, m, bytes(range(256)))
    assert selftest_digest(oracle, shards) == G.ERASURE_SELFTEST[km]
    first = shards[0].copy()
    shards[0] = None
    assert oracle.reconstruct(k, m, shards, data_only=True) == 0
    assert np.array_equal(shards[0], first)


@pytest.mark.parametrize("algo", [1, 2, 3, 4])
def test_bitrot_selftest_golden(oracle, algo):
    """cmd/bitrot.go:224-254: msg grows by its own digest, Size()*BlockSize()/Size() rounds."""
    size, bsz = (64, 128) if algo == 4 else ((32, 64) if algo == 1 else (32, 32))
    msg, s = b"", b""
    for _ in range(0, size * bsz, size):
        s = oracle.bitrot_hash(algo, msg)
        msg += s
    assert s.hex() == G.BITROT_SELFTEST[algo]


def test_magic_key_is_hh_of_pi(oracle):
    """cmd/bitrot.go:36-37"""
    assert oracle.hh256(G.PI_100.encode(), key=bytes(32)).hex() == G.MAGIC_KEY_HEX
    assert oracle.MAGIC_KEY.hex() == G.MAGIC_KEY_HEX


def test_std_hashes_vs_hashlib(oracle):
    rng = np.random.default_rng(1)
    for n in [0, 1, 55, 56, 63, 64, 65, 127, 128, 129, 255, 256, 1000, 4096 + 17]:
        msg = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert oracle.sha256(msg) == hashlib.sha256(msg).digest()
        assert oracle.blake2b512(msg) == hashlib.blake2b(msg).digest()
    try:
        import xxhash
    except ImportError:
        return
    for n in [0, 1, 3, 4, 7, 8, 31, 32, 33, 100, 1000]:
        msg = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert oracle.xxh64(msg) == xxhash.xxh64(msg).intdigest()


def test_rs75_fixture(oracle):
    """Real shard bytes written by MinIO (cmd/testdata/undeleteable-object.tgz): RS(7,5), 1 MiB blocks."""
    z = np.load(os.path.join(GOLD, "rs75_fixture.npz"))
    k, m, S = int(z["k"]), int(z["m"]), int(z["shard_size"])
    assert S == oracle.shard_size(int(z["block_size"]), k) == 149797
    M = oracle.rs_matrix(k, m)
    for part in ("head", "tail"):
        for b in range(2):
            sl = z[part][b]  # [12][4096]
            enc = [sl[i].copy() for i in range(k)] + [np.zeros_like(sl[0]) for _ in range(m)]
            import ctypes as C
            pa = (C.c_void_p * (k + m))(*[a.ctypes.data for a in enc])
            assert oracle.lib().orc_rs_encode(k, m, pa, sl.shape[1]) == 0
            for j in range(m):
                assert np.array_equal(enc[k + j], sl[k + j]), (part, b, j)
    # Split zero padding: 7*149797 - 2^20 = 3 pad bytes at the end of data shard 7
    assert not z["tail"][0][k - 1][-3:].any() and not z["tail"][1][k - 1][-3:].any()
    for name, idx, blk in (("frame_data", 1, 0), ("frame_parity", 9, 2)):
        fr = z[name]
        assert len(fr) == 32 + S
        assert oracle.hh256(fr[32:]) == fr[:32].tobytes() == z["digests"][blk, idx].tobytes()
        assert oracle.hh256(fr[32:], fast=True) == fr[:32].tobytes()
    assert M[:k].tolist() == np.eye(k, dtype=np.uint8).tolist()


def test_inline_frames(oracle):
    """[digest||shard] frames cut from fixture xl.meta inline data: HighwayHash with many tails."""
    frames = json.load(open(os.path.join(GOLD, "inline_frames.json")))
    tails = set()
    for f in frames:
        shard = bytes.fromhex(f["shard"])
        assert oracle.hh256(shard).hex() == f["digest"], f["src"]
        assert oracle.hh256(shard, fast=True).hex() == f["digest"]
        assert oracle.bitrot_verify(oracle.HIGHWAYHASH256S, bytes.fromhex(f["digest"]) + shard, len(shard), max(len(shard), 1)) == 0
        tails.add(len(shard) % 32)
    # branches of the remainder code: &16 (17..23), mod4 in {1,2,3} without &16 (3,5,7,13,14), mod4==0 (12)
    assert {3, 5, 7, 12, 13, 14, 17, 18, 20, 21, 23} <= tails


def test_inline_notinline_end_to_end(oracle):
    """cmd/erasure-object_test.go:1131-1184: RS(2,2), shards 2 (data, inline) and 3 (parity, part.1) survive."""
    z = np.load(os.path.join(GOLD, "inline_notinline.npz"))
    from golden.make_fixtures import parse_xlmeta
    _, inline = parse_xlmeta(z["meta_disk1"].tobytes())
    data1 = inline[b"null"]
    part = z["part1"].tobytes()
    size, k, m, bs = 132096, 2, 2, 1 << 20
    S = oracle.shard_size(bs, k)
    fsz = oracle.bitrot_shard_file_size(oracle.shard_file_size(bs, k, size), S, oracle.HIGHWAYHASH256S)
    assert len(data1) == len(part) == fsz == 66080
    empty = np.zeros(fsz, dtype=np.uint8)
    files = [empty, np.frombuffer(data1, dtype=np.uint8), np.frombuffer(part, dtype=np.uint8), empty]
    rc, out, corrupt = oracle.erasure_decode(k, m, bs, oracle.HIGHWAYHASH256S, files, [0, 1, 1, 0], 0, size, size)
    assert rc == size and not corrupt.any()
    assert hashlib.md5(out.tobytes()).hexdigest() == G.INLINE_NOTINLINE_MD5


def test_derived_kats(oracle):
    """DERIVED vectors: not in the reference tree; guard against oracle drift."""
    for km, want in G.DERIVED_SELFTEST.items():
        assert selftest_digest(oracle, oracle.encode_data(*km, bytes(range(256)))) == want
    for (k, m), rows in G.DERIVED_PARITY_ROWS.items():
        M = oracle.rs_matrix(k, m)
        assert [bytes(r).hex() for r in M[k:]] == rows
    for n, want in G.DERIVED_HH.items():
        pat = ((7 * np.arange(n, dtype=np.uint64) + 3) & 0xFF).astype(np.uint8)
        assert oracle.hh256(pat).hex() == want
        assert oracle.hh256(pat, fast=True).hex() == want


def test_simd_variant_matches_scalar(oracle):
    rng = np.random.default_rng(7)
    print("oracle simd level:", oracle.lib().orc_simd_level().decode())
    for (k, m, n) in [(4, 2, 1000), (12, 4, 87382 * 12 - 8), (16, 4, 65536 * 16), (8, 8, 8 * 4097), (7, 5, 1 << 20), (5, 11, 777)]:
        data = rng.integers(0, 256, n, dtype=np.uint8)
        a = oracle.encode_data(k, m, data)
        b = oracle.encode_data(k, m, data, fast=True)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    for n in list(range(0, 70)) + [87382, 65536, 1000003]:
        msg = rng.integers(0, 256, n, dtype=np.uint8)
        assert oracle.hh256(msg) == oracle.hh256(msg, fast=True)


def test_baseline_config1_cpu_roundtrip(oracle):
    """BASELINE.json configs[0]: RS(4,2) encode of one 4 MiB object, 1 MiB blocks, HighwayHash256 bitrot, through the
    C oracle on CPU: frames verify, erase 2 shards -> reconstruct -> bytes equal, heal reproduces the erased files."""
    k, m, bs, size = 4, 2, 1 << 20, 4 << 20
    data = np.random.default_rng(0x4D494E494F00 + 1).integers(0, 256, size, dtype=np.uint8)
    files, _ = oracle.erasure_encode(k, m, bs, oracle.HIGHWAYHASH256S, data)
    S = oracle.shard_size(bs, k)
    assert S == 262144 and all(f.size == 4 * (32 + S) == 1048704 for f in files)
    for f in files:
        assert oracle.bitrot_verify(oracle.HIGHWAYHASH256S, f, oracle.shard_file_size(bs, k, size), S) == 0
    for erased in [(0, 1), (1, 4), (4, 5), (2, 3)]:
        avail = [0 if i in erased else 1 for i in range(k + m)]
        rc, out, corrupt = oracle.erasure_decode(k, m, bs, oracle.HIGHWAYHASH256S, files, avail, 0, size, size)
        assert rc == size and np.array_equal(out, data) and not corrupt.any()
        rc, healed = oracle.erasure_heal(k, m, bs, oracle.HIGHWAYHASH256S, files, avail, [1 - a for a in avail], size)
        assert rc == 0
        for i in erased:
            assert np.array_equal(healed[i], files[i])
    # a range read across block boundaries and one flipped bit (detected, still decodable)
    rc, out, _ = oracle.erasure_decode(k, m, bs, oracle.HIGHWAYHASH256S, files, [1] * 6, bs - 5, 2 * bs + 10, size)
    assert rc == 2 * bs + 10 and np.array_equal(out, data[bs - 5:3 * bs + 5])
    bad = [f.copy() for f in files]
    bad[2][100000] ^= 0x10
    rc, out, corrupt = oracle.erasure_decode(k, m, bs, oracle.HIGHWAYHASH256S, bad, [1] * 6, 0, size, size)
    assert rc == size and np.array_equal(out, data) and corrupt.tolist() == [0, 0, 1, 0, 0, 0]
