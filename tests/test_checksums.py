"""Object checksums (internal/hash/checksum.go, crc.go): the oracle's CRCs are pinned by catalogue check values and zlib; the
library's host-side merge (Checksum.AddPart) is checked against them without a GPU; the GPU kernels are checked in the gpu test."""
import zlib

import numpy as np
import pytest


def test_oracle_crcs_are_pinned(oracle):
    assert oracle.crcs(b"123456789") == (0xCBF43926, 0xE3069283, 0xAE8B14860A799888)
    assert oracle.crcs(b"") == (0, 0, 0)
    d = np.random.default_rng(3).integers(0, 256, 100003, dtype=np.uint8).tobytes()
    assert oracle.crcs(d)[0] == zlib.crc32(d)


def test_combine_is_addpart(oracle):
    """mec_checksum_combine == crc32Combine / crc64Combine of internal/hash/crc.go:98-220 (checked through the defining property
    crc(A || B), against zlib and the oracle; lengths around the powers of two the device tree uses)."""
    import minio_b200.capi as capi
    L = capi.lib()
    rng = np.random.default_rng(9)
    for la, lb in [(1, 1), (7, 1024), (1024, 1024), (1025, 262144), (262143, 5), (100000, 262145), (0, 77), (77, 0)]:
        a = rng.integers(0, 256, la, dtype=np.uint8).tobytes()
        b = rng.integers(0, 256, lb, dtype=np.uint8).tobytes()
        ca, cb, cab = oracle.crcs(a), oracle.crcs(b), oracle.crcs(a + b)
        for t, ty in enumerate((1, 2, 4)):
            assert L.mec_checksum_combine(ty, ca[t], cb[t], lb) == cab[t], (la, lb, ty)
    assert L.mec_checksum_combine(1, zlib.crc32(b"hello "), zlib.crc32(b"world"), 5) == zlib.crc32(b"hello world")


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 15, 16, 1023, 1024, 1025, 262143, 262144, 262145, 5 * (1 << 20) + 4321, 130 * (1 << 20) + 7])
def test_gpu_checksums_match_oracle(oracle, n):
    import minio_b200 as mb
    c = mb.Codec(12, 4)
    d = np.random.default_rng(n % 1000).integers(0, 256, n, dtype=np.uint8)
    want = oracle.crcs(d) if n < (8 << 20) else (zlib.crc32(d.tobytes()), None, None)
    got = c.checksums(d)
    assert got[0] == want[0]
    if want[1] is not None:
        assert got == want
    else:  # long streams: CRC32 against zlib, the others through the merge property over two halves checked by the oracle
        h = n // 2
        a, b = c.checksums(d[:h]), c.checksums(d[h:])
        L = mb.lib()
        for t, ty in enumerate((1, 2, 4)):
            assert L.mec_checksum_combine(ty, a[t], b[t], n - h) == got[t]
    assert c.checksums(d, which=2) == (0, got[1], 0)
    c.close()


@pytest.mark.gpu
def test_encode_carries_checksums(oracle):
    """mec_set_option("checksums"): the host-buffer encode computes the object's CRCs on the bytes it stages anyway."""
    import minio_b200 as mb
    k, m, bs, n = 12, 4, 1 << 20, 150 * (1 << 20) + 12345
    d = np.random.default_rng(5).integers(0, 256, n, dtype=np.uint8)
    c = mb.Codec(k, m, bs)
    c.set_option("checksums", 7)
    files, dd = c.encode_sg(d)
    got, length = c.last_checksums()
    assert length == n and got[0] == zlib.crc32(d.tobytes())
    assert got == c.checksums(d)
    want, _ = oracle.erasure_encode(k, m, bs, oracle.HIGHWAYHASH256S, d[:3 * bs])
    assert np.array_equal(files[k][:want[k].size], want[k])   # the encode itself is unchanged
    c.close()
