"""CPU-side checks of the C-ABI library: it loads, exports exactly what include/minio_ec.h declares,
validates arguments like NewErasure, mirrors the size helpers, and fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mb():
    import minio_b200
    return minio_b200


def test_exports_match_header(mb):
    hdr = open(os.path.join(ROOT, "include", "minio_ec.h")).read()
    declared = set(re.findall(r"\b(mec_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    out = subprocess.check_output(["nm", "-D", "--defined-only", mb.lib_path()], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T mec_" in ln}
    assert declared == exported, (declared - exported, exported - declared)
    L = mb.lib()
    for name in declared:
        assert getattr(L, name) is not None


def test_new_erasure_validation(mb):
    """cmd/erasure-coding.go:44-50"""
    for k, m, code in [(0, 2, -1), (-1, 2, -1), (2, -1, -1), (200, 57, -2), (256, 1, -2)]:
        with pytest.raises(mb.MecError) as ei:
            mb.Codec(k, m)
        assert ei.value.code == code
    mb.Codec(256, 0).close()
    mb.Codec(12, 4).close()


@pytest.mark.parametrize("k,m,bs", [(12, 4, 1 << 20), (4, 2, 1 << 20), (16, 4, 1 << 20), (8, 8, 256 * 1024), (7, 5, 1 << 20), (5, 3, 64), (2, 2, (1 << 20) - 1)])
def test_size_helpers_vs_oracle(mb, oracle, k, m, bs):
    """Erasure.ShardSize / ShardFileSize / ShardFileOffset / bitrotShardFileSize / ceilFrac"""
    c = mb.Codec(k, m, bs)
    assert c.shard_size() == oracle.shard_size(bs, k)
    rng = np.random.default_rng(3)
    totals = [0, -1, 1, bs - 1, bs, bs + 1, 5 * bs + 77, 10 << 30] + [int(x) for x in rng.integers(1, 50 * bs, 20)]
    for total in totals:
        assert c.shard_file_size(total) == oracle.shard_file_size(bs, k, total)
        if total > 0:
            assert c.bitrot_file_size(total) == oracle.bitrot_shard_file_size(oracle.shard_file_size(bs, k, total), oracle.shard_size(bs, k), 3)
            for _ in range(5):
                off = int(rng.integers(0, total))
                ln = int(rng.integers(0, total - off + 1))
                assert c.shard_file_offset(off, ln, total) == oracle.shard_file_offset(bs, k, off, ln, total)
    L = mb.lib()
    for a, b in [(0, 5), (1, 5), (5, 5), (6, 5), (-6, 5), (7, 0), (1 << 40, 12)]:
        assert L.mec_ceil_frac(a, b) == oracle.lib().orc_ceil_frac(a, b)
    c.close()


def test_host_side_argument_errors_without_gpu(mb):
    """Quorum / argument checks run before any device work (cmd/erasure-encode.go:59-65, erasure-decode.go:240-249)."""
    c = mb.Codec(4, 2)
    with pytest.raises(mb.MecError) as ei:
        c.encode(np.zeros(100, dtype=np.uint8), online=[True, False, False, False, False, False], write_quorum=5)
    assert ei.value.code == -11
    files = [np.zeros(c.bitrot_file_size(100), dtype=np.uint8)] * 6
    for off, ln in [(-1, 5), (0, -1), (50, 51)]:
        with pytest.raises(mb.MecError) as ei:
            c.decode(files, off, ln, 100)
        assert ei.value.code == -12
    out, hint = c.decode(files, 0, 0, 100)
    assert out.size == 0
    assert c.bitrot_verify(np.zeros(10, dtype=np.uint8), 100) == -7   # size check, cmd/bitrot.go:183
    assert all(f.size == 0 for f in c.encode(b""))
    c.close()


def test_no_cpu_fallback(mb):
    """Without a CUDA device every compute call must fail loudly (never route through the oracle)."""
    if mb.device_count() > 0:
        pytest.skip("GPU present")
    c = mb.Codec(12, 4)
    with pytest.raises(mb.MecError) as ei:
        c.encode_blocks(np.zeros(1 << 20, dtype=np.uint8))
    assert ei.value.code == -101
    with pytest.raises(mb.MecError):
        mb.selftest(0)
    # the product never links the oracle
    out = subprocess.check_output(["ldd", mb.lib_path()], text=True)
    assert "oracle" not in out
    syms = subprocess.check_output(["nm", "-D", mb.lib_path()], text=True)
    assert "orc_" not in syms
    c.close()


def test_shutdown_is_callable_without_a_device():
    """mec_shutdown only quiesces the background compiler: safe to call with no GPU, twice, before anything else ran."""
    import minio_b200.capi as capi
    L = capi.lib()
    if L.mec_device_count() > 0:
        pytest.skip("would switch background specialisation off for the GPU tests that share this process")
    L.mec_shutdown()
    L.mec_shutdown()


@pytest.mark.parametrize("case", ["decode-hashed", "decode-get", "encode-10-4"])
def test_nvrtc_specialisation_compiles_without_a_device(case, oracle):
    """The kernel headers embedded in the library must still instantiate under NVRTC for a concrete matrix: decode rows of
    RS(12,4) with four data shards lost (aligned frames, 4 blocks per CTA; with and without digests of the rebuilt shards)
    and the parity rows of a geometry that has no compiled kernel.  Compile only — loading needs a GPU."""
    import minio_b200.capi as capi
    L = capi.lib()
    if case.startswith("decode"):
        present = [0] * 4 + [1] * 12
        rows, _ = oracle.decode_rows(12, 4, present, [0, 1, 2, 3])
        k, r, align, hash_out = 12, 4, 0, 1 if case == "decode-hashed" else 0
    else:
        k, r, align, hash_out = 10, 4, (-(-(1 << 20) // 10)) % 16, 1
        rows = np.array(oracle.rs_matrix(10, 4), dtype=np.uint8).reshape(14, 10)[10:]
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    n = L.mec_jit_compile_check(k, r, rows.ctypes.data, align, 4, 0, hash_out)
    if n == -1:
        pytest.skip("libnvrtc not available")
    assert n > 0, L.mec_last_error().decode()
