"""GPU parity tests: every C-ABI entry point of libminio_ec.so against the CPU oracle, bit-exact.

Mirrors the reference's own test tables (cmd/erasure_test.go:33-43, cmd/erasure-encode_test.go:62-81,
cmd/erasure-decode_test.go:44-83, cmd/erasure-heal_test.go:42-61, cmd/bitrot_test.go) as data.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import selftest_goldens as G

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MiB = 1 << 20


@pytest.fixture(scope="module")
def mb():
    import minio_b200
    assert minio_b200.device_count() > 0, "no CUDA device"
    return minio_b200


@pytest.fixture(autouse=True, params=["latency", "throughput"])
def kernel_form(request):
    """Every test of this module runs twice: small launches through the latency kernel (ec_small.cuh, the default for launches
    that cannot fill the GPU) and with it switched off, so that the throughput kernel keeps its coverage of the same cases."""
    old = os.environ.pop("MEC_SMALL_BLOCKS", None)
    if request.param == "throughput":
        os.environ["MEC_SMALL_BLOCKS"] = "0"
    yield request.param
    os.environ.pop("MEC_SMALL_BLOCKS", None)
    if old is not None:
        os.environ["MEC_SMALL_BLOCKS"] = old


def rand(n, seed):
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8)


def test_selftest(mb):
    """erasureSelfTest + bitrotSelfTest replayed on the GPU (cmd/erasure-coding.go:149, cmd/bitrot.go:224)."""
    mb.selftest(0)


MODES = [dict(), dict(force_bytewise=1), dict(force_dynamic=1), dict(force_bytewise=1, force_dynamic=1), dict(eb=1), dict(eb=3)]


@pytest.mark.parametrize("k,m,bs,length", [
    (4, 2, MiB, 4 * MiB),            # BASELINE config 1
    (12, 4, MiB, 3 * MiB + 12345),   # S=87382 (2-byte aligned shards, HH tail 22) + short last block
    (12, 4, MiB, 7 * MiB),
    (16, 4, MiB, 5 * MiB),
    (8, 8, 256 * 1024, 2 * MiB + 1),
    (7, 5, MiB, 2 * MiB + 77),       # odd shard size 149797 (fixture geometry)
    (2, 2, MiB, 132096),
    (8, 4, MiB - 1, 3 * MiB),        # block size not a multiple of 16 -> byte-wise loader
    (5, 3, 64, 1000),                # tiny blocks
    (12, 4, MiB, 1), (12, 4, MiB, 11), (12, 4, MiB, 12), (12, 4, MiB, 13), (12, 4, MiB, 255 * 12), (12, 4, MiB, 256 * 12 + 5),
    (3, 0, MiB, MiB + 5),            # parity = 0 is legal
    (1, 1, MiB, 100000), (15, 1, MiB, MiB), (9, 7, MiB, MiB + 9),
])
@pytest.mark.parametrize("mode", MODES[:4])
def test_encode_blocks_vs_oracle(mb, oracle, k, m, bs, length, mode):
    data = rand(length, k * 1000 + m)
    c = mb.Codec(k, m, bs)
    for name, v in mode.items():
        c.set_option(name, v)
    parity, dig = c.encode_blocks(data)
    S = oracle.shard_size(bs, k)
    nb = -(-length // bs)
    for b in range(nb):
        blk = data[b * bs:(b + 1) * bs]
        sh = oracle.encode_data(k, m, blk, fast=True)
        per = sh[0].size
        for j in range(m):
            assert np.array_equal(parity[b, j, :per], sh[k + j]), (b, j)
        for i in range(k + m):
            assert dig[b, i].tobytes() == oracle.hh256(sh[i], fast=True), (b, i)
    c.close()


@pytest.mark.parametrize("mode", MODES)
def test_encode_files_vs_oracle(mb, oracle, mode):
    for (k, m, bs, length) in [(12, 4, MiB, 5 * MiB + 4321), (4, 2, MiB, 4 * MiB), (6, 2, 512 * 1024, MiB + 3)]:
        data = rand(length, 99)
        want, _ = oracle.erasure_encode(k, m, bs, oracle.HIGHWAYHASH256S, data)
        c = mb.Codec(k, m, bs)
        for name, v in mode.items():
            c.set_option(name, v)
        got = c.encode(data)
        for i in range(k + m):
            assert np.array_equal(got[i], want[i]), i
        c.close()


def test_encode_empty_and_quorum(mb):
    c = mb.Codec(4, 2)
    files = c.encode(b"")
    assert all(f.size == 0 for f in files)
    with pytest.raises(mb.MecError) as ei:
        c.encode(rand(1000, 1), online=[True, True, True, False, False, False], write_quorum=4)
    assert ei.value.code == -11
    got = c.encode(rand(1000, 1), online=[True, True, True, True, False, False], write_quorum=4)
    assert got[4] is None and got[0] is not None


# cmd/erasure_test.go:33-43 (k, m, missing data, missing parity, expect reconstruct failure)
ENCODE_DECODE_TABLE = [
    (2, 2, 0, 0, False), (3, 3, 1, 0, False), (4, 4, 2, 0, False), (5, 5, 0, 1, False), (6, 6, 0, 2, False),
    (7, 7, 1, 1, False), (8, 8, 3, 2, False), (2, 2, 2, 1, True), (4, 2, 2, 2, True), (8, 4, 2, 2, False),
]


@pytest.mark.parametrize("k,m,md,mp,fail", ENCODE_DECODE_TABLE)
def test_erasure_encode_decode_table(mb, oracle, k, m, md, mp, fail):
    """TestErasureEncodeDecode: random 256 bytes, EncodeData, drop shards, DecodeDataAndParityBlocks."""
    data = rand(256, 5)
    c = mb.Codec(k, m, 1 << 20)
    enc = c.encode_data(data)
    want = oracle.encode_data(k, m, data)
    for a, b in zip(enc, want):
        assert np.array_equal(a, b)
    present = np.ones(k + m, dtype=np.uint8)
    present[:md] = 0
    present[k:k + mp] = 0
    broken = [s.copy() if present[i] else np.full_like(s, 0xAA) for i, s in enumerate(enc)]
    rc = c.rs_reconstruct_shards(broken, present)
    if fail:
        assert rc == -3  # ErrTooFewShards
    else:
        assert rc == 0
        for a, b in zip(broken, want):
            assert np.array_equal(a, b)
    # data-only variant leaves parity alone
    broken = [s.copy() if present[i] else np.full_like(s, 0xAA) for i, s in enumerate(enc)]
    rc = c.rs_reconstruct_shards(broken, present, data_only=True)
    if not fail:
        assert rc == 0
        for i in range(k):
            assert np.array_equal(broken[i], want[i])
    c.close()


# cmd/erasure-decode_test.go:44-83 condensed: (k, n, blocksize, size, offset, length, offline data, offline parity)
DECODE_TABLE = [
    (2, 4, MiB, MiB, 0, MiB, 0, 0), (3, 6, MiB, MiB, 0, MiB, 0, 0), (4, 8, MiB, MiB, 0, MiB, 0, 0),
    (5, 10, MiB, MiB, 1, MiB - 1, 0, 0), (6, 12, MiB, MiB, 0, MiB, 0, 0), (7, 14, MiB - 1, MiB, MiB - 1, 1, 0, 0),
    (8, 16, MiB, MiB, 0, MiB, 0, 0), (7, 14, MiB, MiB, 0, MiB, 7, 0), (6, 12, MiB, MiB, 0, MiB, 0, 6),
    (5, 10, MiB, MiB, 0, MiB, 2, 3), (4, 8, MiB, MiB, 0, MiB, 2, 2), (2, 4, MiB, MiB, 0, MiB, 1, 1),
    (7, 14, MiB, 2 * MiB, MiB, MiB + MiB // 2 - MiB, 0, 0), (8, 16, MiB, 2 * MiB + 17, 17, 2 * MiB, 2, 1),
    (12, 16, MiB, 5 * MiB + 100, MiB - 5, 3 * MiB + 50, 1, 3), (12, 16, MiB, 5 * MiB + 100, 5 * MiB, 100, 4, 0),
    (12, 16, MiB, 4 * MiB, 0, 4 * MiB, 0, 4), (12, 16, MiB, 4 * MiB, 3 * MiB, MiB, 2, 0),
]


@pytest.mark.parametrize("k,n,bs,size,off,length,offd,offp", DECODE_TABLE)
def test_decode_table(mb, oracle, k, n, bs, size, off, length, offd, offp):
    m = n - k
    data = rand(size, size % 9973)
    files, _ = oracle.erasure_encode(k, m, bs, oracle.HIGHWAYHASH256S, data)
    avail = [True] * n
    for i in range(offd):
        avail[i] = False
    for i in range(offp):
        avail[k + i] = False
    c = mb.Codec(k, m, bs)
    out, hint = c.decode([f if a else None for f, a in zip(files, avail)], off, length, size)
    assert np.array_equal(out, data[off:off + length])
    assert hint == 0
    rc, ref, _ = oracle.erasure_decode(k, m, bs, oracle.HIGHWAYHASH256S, files, [int(a) for a in avail], off, length, size)
    assert rc == length and np.array_equal(ref, out)
    c.close()


def test_decode_errors(mb, oracle):
    k, m, bs, size = 4, 4, MiB, 2 * MiB
    data = rand(size, 3)
    files, _ = oracle.erasure_encode(k, m, bs, oracle.HIGHWAYHASH256S, data)
    c = mb.Codec(k, m, bs)
    for (off, ln) in [(-1, 10), (0, -1), (10, size)]:
        with pytest.raises(mb.MecError) as ei:
            c.decode(files, off, ln, size)
        assert ei.value.code == -12
    # too many offline drives -> read quorum
    with pytest.raises(mb.MecError) as ei:
        c.decode([None] * 5 + files[5:], 0, size, size)
    assert ei.value.code == -10
    c.close()


def test_decode_with_bitrot(mb, oracle):
    """Corrupt frames: the reader is dropped (errFileCorrupt), data still decodes, heal hint is raised."""
    k, m, bs, size = 12, 4, MiB, 6 * MiB + 999
    data = rand(size, 8)
    files, _ = oracle.erasure_encode(k, m, bs, oracle.HIGHWAYHASH256S, data)
    S = oracle.shard_size(bs, k)
    bad = [f.copy() for f in files]
    bad[0][3 * (32 + S) + 32 + 100] ^= 0x01      # shard 0 corrupt in block 3 (data byte)
    bad[5][1 * (32 + S) + 7] ^= 0x80             # shard 5 corrupt in block 1 (digest byte)
    c = mb.Codec(k, m, bs)
    out, hint = c.decode(bad, 0, size, size)
    assert np.array_equal(out, data) and hint == -7
    rc, ref, corrupt = oracle.erasure_decode(k, m, bs, oracle.HIGHWAYHASH256S, bad, [1] * 16, 0, size, size)
    assert rc == size and np.array_equal(ref, data) and corrupt[0] and corrupt[5]
    # 5 corrupt files > parity: unrecoverable
    for i in (1, 2, 3):
        bad[i][40] ^= 0xFF
    with pytest.raises(mb.MecError) as ei:
        c.decode(bad, 0, size, size)
    assert ei.value.code == -10
    c.close()


# cmd/erasure-heal_test.go:42-61 condensed: (k, n, size, blocksize, offline disks (stale), bad stale disks)
HEAL_TABLE = [
    (2, 4, MiB, MiB, [0]), (3, 6, MiB, MiB, [0, 5]), (4, 8, MiB, MiB, [1, 2, 6]), (5, 10, MiB, MiB, [9]),
    (6, 12, MiB, 4096, [0, 1, 2, 3]), (7, 14, MiB, MiB, [13]), (8, 16, MiB, MiB, [0, 8, 15, 3, 4, 7, 9, 12]),
    (7, 14, MiB, MiB - 1, [2, 3, 4]), (2, 4, 12345, MiB, [0, 1]), (12, 16, MiB + 1, MiB, [0, 7, 12, 15]),
    (12, 16, 8 * MiB + 300, MiB, [1, 5, 12, 15]), (16, 20, 3 * MiB, MiB, [0, 7, 16, 19]),
]


@pytest.mark.parametrize("k,n,size,bs,stale", HEAL_TABLE)
def test_heal_table(mb, oracle, k, n, size, bs, stale):
    m = n - k
    data = rand(size, size % 7919)
    files, _ = oracle.erasure_encode(k, m, bs, oracle.HIGHWAYHASH256S, data)
    c = mb.Codec(k, m, bs)
    st = [i in stale for i in range(n)]
    outs = c.heal([None if st[i] else files[i] for i in range(n)], st, size)
    for i in range(n):
        if st[i]:
            assert np.array_equal(outs[i], files[i]), i
    c.close()


def test_bitrot_verify(mb, oracle):
    """bitrotVerify / TestXLStorageVerifyFile: good file, flipped byte, wrong length."""
    k, m, bs, size = 12, 4, MiB, 4 * MiB + 5555
    files, _ = oracle.erasure_encode(k, m, bs, oracle.HIGHWAYHASH256S, rand(size, 4))
    c = mb.Codec(k, m, bs)
    part = c.shard_file_size(size)
    for f in (files[0], files[13]):
        assert c.bitrot_verify(f, part) == 0
        g = f.copy(); g[len(g) // 2] ^= 4
        assert c.bitrot_verify(g, part) == -7
        assert c.bitrot_verify(f[:-1], part) == -7
        g = f.copy(); g[3] ^= 1   # digest byte of the first frame
        assert c.bitrot_verify(g, part) == -7
    c.close()


def test_golden_fixtures_on_gpu(mb, oracle):
    """The reference's own bytes through the GPU: RS(7,5) shard slices + frames, inline frames, MD5 golden."""
    z = np.load(os.path.join(GOLD, "rs75_fixture.npz"))
    k, m, S = int(z["k"]), int(z["m"]), int(z["shard_size"])
    c = mb.Codec(k, m, 1 << 20)
    for part in ("head", "tail"):
        for b in range(2):
            sl = z[part][b]
            sh = [sl[i].copy() for i in range(k)] + [np.zeros_like(sl[0]) for _ in range(m)]
            c.rs_encode_shards(sh)
            for j in range(m):
                assert np.array_equal(sh[k + j], sl[k + j])
    for name in ("frame_data", "frame_parity"):
        fr = z[name]
        assert c.hh256_batch(fr[32:], S, 1)[0].tobytes() == fr[:32].tobytes()
    c.close()
    frames = json.load(open(os.path.join(GOLD, "inline_frames.json")))
    c = mb.Codec(2, 2)
    for f in frames:
        shard = np.frombuffer(bytes.fromhex(f["shard"]), dtype=np.uint8)
        for mode in (0, 1):
            c.set_option("force_bytewise", mode)
            assert c.hh256_batch(shard, shard.size, 1)[0].tobytes().hex() == f["digest"], (f["len"], mode)
    c.set_option("force_bytewise", 0)
    # cmd/erasure-object_test.go:1131-1184
    from golden.make_fixtures import parse_xlmeta
    zz = np.load(os.path.join(GOLD, "inline_notinline.npz"))
    _, inline = parse_xlmeta(zz["meta_disk1"].tobytes())
    files = [None, np.frombuffer(inline[b"null"], dtype=np.uint8), zz["part1"], None]
    out, hint = c.decode(files, 0, 132096, 132096)
    assert hashlib.md5(out.tobytes()).hexdigest() == G.INLINE_NOTINLINE_MD5 and hint == 0
    c.close()


def test_hh_all_tail_lengths(mb, oracle):
    c = mb.Codec(2, 2)
    for n in list(range(1, 100)) + [255, 256, 257, 511, 512, 513, 1000, 4095, 87382, 65536]:
        msgs = rand(n * 3, n)
        got = c.hh256_batch(msgs, n, 3)
        for i in range(3):
            assert got[i].tobytes() == oracle.hh256(msgs[i * n:(i + 1) * n]), n
    for d, want in G.DERIVED_HH.items():
        if d:
            pat = ((7 * np.arange(d, dtype=np.uint64) + 3) & 0xFF).astype(np.uint8)
            assert c.hh256_batch(pat, d, 1)[0].tobytes().hex() == want
    c.close()


def test_device_resident_large_roundtrip(mb, oracle):
    """BASELINE config 2/3 shape at reduced length: device-resident encode, sampled oracle check,
    digest-of-digests, then erase 4 shards -> heal -> identical frames."""
    import torch
    k, m, bs = 12, 4, MiB
    nblocks = 512
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(0x4D494E494F00 + 2)
    src = torch.randint(0, 256, (nblocks * bs,), dtype=torch.uint8, device=dev, generator=g)
    S = 87382
    pitch = (S + 15) // 16 * 16
    par = torch.zeros((nblocks * m, pitch), dtype=torch.uint8, device=dev)
    dig = torch.zeros((nblocks, k + m, 32), dtype=torch.uint8, device=dev)
    c = mb.Codec(k, m, bs)
    c.encode_blocks_device(src.data_ptr(), src.numel(), par.data_ptr(), pitch, dig.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    h_src, h_par, h_dig = src.cpu().numpy(), par.cpu().numpy(), dig.cpu().numpy()
    for b in (0, 1, 255, 256, 511):
        sh = oracle.encode_data(k, m, h_src[b * bs:(b + 1) * bs], fast=True)
        for j in range(m):
            assert np.array_equal(h_par[b * m + j, :S], sh[k + j])
        for i in range(k + m):
            assert h_dig[b, i].tobytes() == oracle.hh256(sh[i], fast=True)
    # all blocks: every digest must match the CPU hash of the bytes it covers
    for b in range(0, nblocks, 37):
        for j in range(m):
            assert h_dig[b, k + j].tobytes() == oracle.hh256(h_par[b * m + j, :S], fast=True)
    # round trip through frames: heal shards {0, 7, 12, 15} from the rest
    nb2 = 64
    files = c.encode(h_src[:nb2 * bs])
    stale = [i in (0, 7, 12, 15) for i in range(16)]
    outs = c.heal([None if stale[i] else files[i] for i in range(16)], stale, nb2 * bs)
    for i in range(16):
        if stale[i]:
            assert np.array_equal(outs[i], files[i])
    c.close()


@pytest.mark.parametrize("jit", [0, 1])
@pytest.mark.parametrize("erased", [[0, 1, 2, 3], [5], [2, 13]])
def test_reconstruct_device_without_output_digests(mb, oracle, erased, jit):
    """GetObject shape (cmd/erasure-decode.go:283-289): the shards read are hashed and checked, the rebuilt data shards are
    only written — MEC_RECONSTRUCT_NO_OUTPUT_DIGESTS leaves their digest slots untouched and the bytes identical."""
    import torch
    k, m, bs, nblocks = 12, 4, MiB, 24
    n, S = k + m, 87382
    dev = torch.device("cuda:0")
    data = rand(nblocks * bs, 4242)
    files, _ = oracle.erasure_encode(k, m, bs, oracle.HIGHWAYHASH256S, data)
    fp = (32 + S + 15) // 16 * 16
    frames = []
    for i in range(n):
        f = torch.zeros((nblocks, fp), dtype=torch.uint8, device=dev)
        f[:, :32 + S] = torch.from_numpy(files[i].reshape(nblocks, 32 + S)).to(dev)
        frames.append(f)
    frames[7][3, 40] ^= 0x10  # one flipped bit in a shard that is read
    c = mb.Codec(k, m, bs)
    c.set_option("jit", jit)
    want = [1 if i in erased else 0 for i in range(n)]
    ptrs = [0 if i in erased else frames[i].data_ptr() for i in range(n)]
    targets = [i for i in erased if i < k]
    r = len(targets)
    opitch = (S + 15) // 16 * 16
    for flags in (1, 3):
        out = torch.zeros((nblocks * r, opitch), dtype=torch.uint8, device=dev)
        dig = torch.full((nblocks, k + r, 32), 0xA5, dtype=torch.uint8, device=dev)
        cor = torch.zeros((nblocks, k), dtype=torch.uint8, device=dev)
        c.reconstruct_device(ptrs, fp, nblocks, want, flags, out.data_ptr(), opitch, dig.data_ptr(), cor.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        chosen = [i for i in range(n) if i not in erased][:k]
        h_cor = cor.cpu().numpy()
        bad = np.argwhere(h_cor)
        assert bad.tolist() == [[3, chosen.index(7)]]
        h_out = out.cpu().numpy().reshape(nblocks, r, opitch)
        h_dig = dig.cpu().numpy()
        for b in range(nblocks):
            if b == 3:
                continue  # rebuilt from a corrupt frame: the caller drops this block (fail-over is the caller's job)
            for q, i in enumerate(targets):
                ref = files[i].reshape(nblocks, 32 + S)[b]
                assert np.array_equal(h_out[b, q, :S], ref[32:]), (flags, b, i)
                if flags & 2:
                    assert (h_dig[b, k + q] == 0xA5).all()
                else:
                    assert h_dig[b, k + q].tobytes() == ref[:32].tobytes()
            for t, i in enumerate(chosen):
                if i != 7 or b != 3:
                    assert h_dig[b, t].tobytes() == files[i].reshape(nblocks, 32 + S)[b, :32].tobytes()
    c.close()


@pytest.mark.parametrize("k,m,stale", [(12, 4, [0, 1, 2, 3]), (12, 4, [1, 5, 12, 15]), (16, 4, [0, 7, 16, 19]), (8, 8, [0, 1, 2, 3, 8, 9, 10, 11]), (4, 2, [1, 4])])
def test_heal_with_runtime_specialised_kernels(mb, oracle, k, m, stale):
    """Decode matrices compiled with NVRTC at run time (option jit=1) must give the same bytes as the generic kernel."""
    bs, size = MiB, 5 * MiB + 1234
    data = rand(size, 77)
    files, _ = oracle.erasure_encode(k, m, bs, oracle.HIGHWAYHASH256S, data)
    n = k + m
    st = [i in stale for i in range(n)]
    for jit in (1, 0):
        c = mb.Codec(k, m, bs)
        c.set_option("jit", jit)
        outs = c.heal([None if st[i] else files[i] for i in range(n)], st, size)
        for i in stale:
            assert np.array_equal(outs[i], files[i]), (jit, i)
        out, hint = c.decode([None if st[i] else files[i] for i in range(n)], 17, size - 17, size)
        assert np.array_equal(out, data[17:]) and hint == 0
        c.close()


@pytest.mark.parametrize("k,m", [(12, 4), (8, 8)])
def test_warp_autonomous_variant(mb, oracle, k, m):
    """Option use_auto = 1 (k + m == 16: one warp per erasure block, no CTA barriers in the steady state; measured slower,
    off by default) must still produce the reference bytes — for the misaligned (12,4) and the 16-byte aligned (8,8) shards."""
    bs, size = MiB, 3 * MiB + 123
    data = rand(size, 31 + k)
    want, _ = oracle.erasure_encode(k, m, bs, oracle.HIGHWAYHASH256S, data)
    c = mb.Codec(k, m, bs)
    c.set_option("use_auto", 1)
    c.set_option("no_rows3d", 1)
    files = c.encode(data)
    for i in range(k + m):
        assert np.array_equal(files[i], want[i]), i
    out, _ = c.decode([None, None] + files[2:], 0, size, size)
    assert np.array_equal(out, data)
    c.close()


def test_heal_batch_over_a_codec_pool(mb, oracle):
    """mec_heal_batch (BASELINE config 4 shape, scaled down): many objects, different lengths and stale sets, healed through
    a pool of handles by concurrent host threads; every rebuilt shard file equals the oracle's."""
    k, m, bs = 16, 4, MiB
    n = k + m
    pool = [mb.Codec(k, m, bs) for _ in range(3)]
    objects, want = [], []
    for o in range(7):
        size = [3 * MiB, 5 * MiB + 17, 1, MiB, 2 * MiB + 999999, 777, 4 * MiB][o]
        data = rand(size, 500 + o)
        files, _ = oracle.erasure_encode(k, m, bs, oracle.HIGHWAYHASH256S, data)
        stale_set = [(0, 7, 16, 19), (1,), (2, 3), (19,), (0, 1, 2, 3), (5, 17), (4, 8, 12, 16)][o]
        stale = [i in stale_set for i in range(n)]
        objects.append(([None if stale[i] else files[i] for i in range(n)], stale, size))
        want.append((files, stale_set))
    outs = mb.heal_batch(pool, objects)
    for (files, stale_set), got in zip(want, outs):
        for i in range(n):
            if i in stale_set:
                assert np.array_equal(got[i], files[i])
            else:
                assert got[i] is None
    for c in pool:
        c.close()


def test_background_specialisation(mb, oracle, kernel_form):
    """Default policy (option jit = -1): the first calls of a new erasure pattern run the generic kernel while NVRTC works on
    a background thread; once the pattern is compiled later calls use the specialised kernel.  Same bytes either way."""
    import time
    if kernel_form == "latency":
        pytest.skip("needs a pattern nothing in the process has compiled yet: runs once, in the throughput form")
    k, m, bs, size = 6, 3, MiB, 48 * MiB + 999  # a geometry no other test uses: nothing in the process-wide cache yet
    data = rand(size, 99)
    files, _ = oracle.erasure_encode(k, m, bs, oracle.HIGHWAYHASH256S, data)
    n = k + m
    stale = [i in (1, 7) for i in range(n)]
    c = mb.Codec(k, m, bs)
    c.set_option("small_blocks", 0)  # the chunks of this heal are small launches: the latency kernel would serve them and never ask for a specialisation
    srcs = [None if stale[i] else files[i] for i in range(n)]
    outs = c.heal(srcs, stale, size)  # 48 MiB of input >= the 32 MiB warm-up: queued for compilation, served by the generic kernel
    assert c.stat("jit_launches") == 0
    for i in (1, 7):
        assert np.array_equal(outs[i], files[i])
    # the compile queue is process-wide (other tests' patterns may be ahead of this one): poll with real calls
    t0 = time.time()
    while c.stat("jit_launches") == 0 and time.time() - t0 < 120:
        time.sleep(0.2)
        outs = c.heal(srcs, stale, size)
    assert c.stat("jit_launches") >= 1, "the specialised kernel never took over"
    for i in (1, 7):
        assert np.array_equal(outs[i], files[i])
    c.close()


@pytest.mark.parametrize("algo", [1, 2, 4])  # SHA256, HighwayHash256 (whole-file), BLAKE2b512
def test_whole_file_bitrot(mb, oracle, algo):
    """wholeBitrotWriter (cmd/bitrot-whole.go:35-45) + BitrotAlgorithm.New (cmd/bitrot.go:47-64) + bitrotVerify's whole-file
    branch (:165-175): shard files are raw, one digest over the whole file; also the bitrotSelfTest chain per algorithm."""
    import hashlib
    c0 = mb.Codec(2, 2)
    # hash.Hash shape vs hashlib / oracle, many lengths incl. padding boundaries
    for n in [0, 1, 31, 32, 55, 56, 63, 64, 65, 111, 112, 127, 128, 129, 255, 1000, 87382, 262144 + 5]:
        msgs = rand(3 * n, n + algo)
        got = c0.whole_hash(algo, msgs, n, 3)
        for i in range(3):
            assert got[i].tobytes() == oracle.bitrot_hash(algo, msgs[i * n:(i + 1) * n]), (algo, n)
    if algo == 1:
        m = rand(5000, 1)
        assert c0.whole_hash(1, m, 5000, 1)[0].tobytes() == hashlib.sha256(m.tobytes()).digest()
    # bitrotSelfTest chain (cmd/bitrot.go:224-254)
    size, bsz = (64, 128) if algo == 4 else ((32, 64) if algo == 1 else (32, 32))
    msg, s = b"", b""
    for _ in range(0, size * bsz, size):
        s = c0.whole_hash(algo, np.frombuffer(msg, dtype=np.uint8), len(msg), 1)[0].tobytes()
        msg += s
    assert s.hex() == G.BITROT_SELFTEST[algo]
    c0.close()
    # Erasure.Encode with whole-file writers vs the oracle driver
    for (k, m, bs, length) in [(8, 8, 256 * 1024, 3 * 256 * 1024 + 777), (4, 2, MiB, 2 * MiB), (12, 4, MiB, MiB + 12345), (3, 0, 4096, 10000)]:
        data = rand(length, 5)
        want_files, want_sums = oracle.erasure_encode(k, m, bs, algo, data)
        c = mb.Codec(k, m, bs, algo=algo)
        files, sums = c.encode_whole(data)
        for i in range(k + m):
            assert np.array_equal(files[i], want_files[i]), (algo, k, m, i)
            assert sums[i] == want_sums[i], (algo, k, m, i)
            assert c.bitrot_verify_whole(algo, files[i], sums[i]) == 0
        bad = files[0].copy(); bad[len(bad) // 2] ^= 1
        assert c.bitrot_verify_whole(algo, bad, sums[0]) == -7
        with pytest.raises(mb.MecError):
            c.encode(data)   # the streaming entry point refuses whole-file algorithms
        c.close()


@pytest.mark.parametrize("k,m", [(20, 12), (32, 16), (31, 1), (17, 3)])
def test_wide_geometries(mb, oracle, k, m):
    """NewErasure allows k+m <= 256 (cmd/erasure-coding.go:48); the GPU path covers k <= 32, m <= 16."""
    bs, size = 256 * 1024, 3 * 256 * 1024 + 999
    data = rand(size, k * 7 + m)
    want, _ = oracle.erasure_encode(k, m, bs, oracle.HIGHWAYHASH256S, data)
    c = mb.Codec(k, m, bs)
    got = c.encode(data)
    for i in range(k + m):
        assert np.array_equal(got[i], want[i]), i
    stale = [i % 3 == 0 and i < 3 * min(m, 5) for i in range(k + m)]
    outs = c.heal([None if stale[i] else got[i] for i in range(k + m)], stale, size)
    for i in range(k + m):
        if stale[i]:
            assert np.array_equal(outs[i], want[i]), i
    c.close()


def test_geometry_limits_fail_loudly(mb):
    c = mb.Codec(33, 4, 1 << 20)          # legal for NewErasure, outside the GPU path
    with pytest.raises(mb.MecError) as ei:
        c.encode(rand(1 << 20, 1))
    assert ei.value.code == -102
    c.close()


def test_decode_random_corruption_matches_oracle(mb, oracle):
    """Random bit flips in random frames: same bytes out, same readers reported corrupt, same failure when too many."""
    k, m, bs, size = 6, 3, 65536, 20 * 65536 + 123
    rng = np.random.default_rng(2024)
    data = rand(size, 11)
    files, _ = oracle.erasure_encode(k, m, bs, oracle.HIGHWAYHASH256S, data)
    c = mb.Codec(k, m, bs)
    for trial in range(12):
        bad = [f.copy() for f in files]
        nbad = int(rng.integers(0, 5))
        for i in rng.choice(k + m, nbad, replace=False):
            bad[i][int(rng.integers(0, bad[i].size))] ^= 1 << int(rng.integers(0, 8))
        rc, ref, corrupt = oracle.erasure_decode(k, m, bs, oracle.HIGHWAYHASH256S, bad, [1] * (k + m), 0, size, size)
        if rc == size:
            out, hint = c.decode(bad, 0, size, size)
            assert np.array_equal(out, data)
            assert (hint == -7) == bool(corrupt.any())
        else:
            with pytest.raises(mb.MecError) as ei:
                c.decode(bad, 0, size, size)
            assert ei.value.code == rc == -10
    c.close()


def test_concurrent_callers(mb, oracle):
    """Thousands of request goroutines share codecs in MinIO (SURVEY §8b 'Threading'): one codec used from several
    threads, and several codecs on one device at once, must stay bit-exact."""
    import threading
    k, m, bs = 12, 4, MiB
    datas = [rand(3 * MiB + 1000 * i + 7, 100 + i) for i in range(6)]
    wants = [oracle.erasure_encode(k, m, bs, oracle.HIGHWAYHASH256S, d)[0] for d in datas]
    shared = mb.Codec(k, m, bs)
    errors = []

    def worker(i, codec):
        try:
            for _ in range(3):
                got = codec.encode(datas[i])
                assert all(np.array_equal(a, b) for a, b in zip(got, wants[i]))
                out, hint = codec.decode([None, None] + got[2:], 5, datas[i].size - 5, datas[i].size)
                assert np.array_equal(out, datas[i][5:]) and hint == 0
        except Exception as e:  # noqa: BLE001
            errors.append((i, repr(e)))

    own = [mb.Codec(k, m, bs) for _ in range(3)]
    threads = [threading.Thread(target=worker, args=(i, shared if i < 3 else own[i - 3])) for i in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    shared.close()
    for c in own:
        c.close()


def test_prefer_readers_and_stats(mb, oracle):
    """parallelReader.preferReaders (cmd/erasure-decode.go:92-123): preferred readers are read first; a corrupt shard on a
    non-preferred drive is then never touched, on a preferred drive it is detected.  Also the boundary counters."""
    k, m, bs, size = 4, 4, MiB, 3 * MiB + 11
    data = rand(size, 21)
    files, _ = oracle.erasure_encode(k, m, bs, oracle.HIGHWAYHASH256S, data)
    bad = [f.copy() for f in files]
    bad[1][5000] ^= 0x40                       # data shard 1 is corrupt
    c = mb.Codec(k, m, bs)
    out, hint = c.decode(bad, 0, size, size)   # default order reads shards 0..3 -> detects it
    assert np.array_equal(out, data) and hint == -7
    assert c.stat("corrupt_shards") == 1
    out, hint = c.decode(bad, 0, size, size, prefer=[0, 0, 0, 0, 1, 1, 1, 1])   # parity drives are "local": shard 1 never read
    assert np.array_equal(out, data) and hint == 0
    out, hint = c.decode(bad, 0, size, size, prefer=[0, 1, 0, 0, 0, 0, 1, 0])   # swap order [1,6,2,3,...]: reads 1,6,2,3
    assert np.array_equal(out, data) and hint == -7
    assert c.stat("blocks_read") > 0 and c.stat("shards_rebuilt") > 0 and c.stat("launches") == c.launches
    assert c.stat("no-such-counter") == -1
    h2d0, d2h0 = c.stat("bytes_h2d"), c.stat("bytes_d2h")
    assert h2d0 > 0 and d2h0 >= 3 * size   # three decodes staged survivor frames and copied the object bytes back
    c.encode_blocks(data)
    assert c.stat("blocks_encoded") == 4 and c.stat("bytes_h2d") - h2d0 == size
    c.close()
