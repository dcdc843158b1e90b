#!/usr/bin/env python
"""The reference's own micro-benchmarks as bench points: BenchmarkErasureEncode* (cmd/erasure-encode_test.go:204-243) and
BenchmarkErasureDecode* (cmd/erasure-decode_test.go:336-385) — (2,2) ... (8,8), 64 KiB ... 40 MiB objects, drives knocked out —
through the host-buffer C ABI (mec_encode / mec_decode, pinned buffers, one caller, one object per call: exactly the Go benchmark's
loop body).  MB/s as `go test -bench` prints it (b.SetBytes(size)); next to it the C oracle's SIMD encode + HighwayHash on one
host thread for the encode shapes.  One JSON object per line + a markdown table on stderr."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import minio_b200 as mb  # noqa: E402
import oracle_lib as o  # noqa: E402

MiB = 1 << 20
ENC = [("EncodeQuick 12MB", 2, 2, 12 * MiB, [(0, 0), (0, 1), (1, 0)]),
       ("Encode_4_64KB", 2, 2, 64 << 10, [(0, 0), (0, 1), (1, 0)]),
       ("Encode_8_20MB", 4, 4, 20 * MiB, [(0, 0), (0, 1), (1, 0), (0, 3), (3, 0)]),
       ("Encode_12_30MB", 6, 6, 30 * MiB, [(0, 0), (0, 1), (1, 0), (0, 5), (5, 0)]),
       ("Encode_16_40MB", 8, 8, 40 * MiB, [(0, 0), (0, 1), (1, 0), (0, 7), (7, 0)])]
DEC = [("DecodeQuick 12MB", 2, 2, 12 * MiB, [(0, 0), (0, 1), (1, 0), (1, 1)]),
       ("Decode_4_64KB", 2, 2, 64 << 10, [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)]),
       ("Decode_8_20MB", 4, 4, 20 * MiB, [(0, 0), (0, 1), (1, 0), (1, 1), (0, 4), (2, 2), (4, 0)]),
       ("Decode_12_30MB", 6, 6, 30 * MiB, [(0, 0), (0, 1), (1, 0), (1, 1), (0, 6), (3, 3), (6, 0)]),
       ("Decode_16_40MB", 8, 8, 40 * MiB, [(0, 0), (0, 1), (1, 0), (1, 1), (0, 8), (4, 4), (8, 0)])]


def pattern(k, m, dd, pd):
    return "".join("X" if i < dd else "0" for i in range(k)) + "|" + "".join("X" if i < pd else "0" for i in range(m))


def timeit(fn, min_s=0.25, min_iters=5):
    fn(); fn()
    n, t0 = 0, time.perf_counter()
    while True:
        fn(); n += 1
        dt = time.perf_counter() - t0
        if dt >= min_s and n >= min_iters:
            return dt / n


def main():
    L = mb.lib()
    L.mec_bind_thread_to_device(0)
    rows = []
    for name, k, m, size, pats in ENC:
        c = mb.Codec(k, m, MiB)
        c.set_option("jit", 1)
        src = mb.capi.pinned_array(size, device=0)
        src[:] = np.random.default_rng(size % 97).integers(0, 256, size, dtype=np.uint8)
        fsz = c.bitrot_file_size(size)
        files = [mb.capi.pinned_array(fsz, device=0) for _ in range(k + m)]
        # CPU: one thread, SIMD RS + HighwayHash of every shard (what one Go benchmark iteration computes; klauspost may add goroutines)
        nb = -(-size // MiB)
        par = np.zeros(nb * m * c.shard_size(), dtype=np.uint8); dig = np.zeros(nb * (k + m) * 32, dtype=np.uint8)
        full = size // MiB
        cpu = None
        if full:
            lo = o.lib()
            lo.orc_encode_hash_blocks_st.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
            lo.orc_encode_hash_blocks_st.restype = None
            cpu = full * MiB / timeit(lambda: lo.orc_encode_hash_blocks_st(k, m, MiB, src.ctypes.data, full, par.ctypes.data, dig.ctypes.data)) / 1e6
        for dd, pd in pats:
            fp = (C.c_void_p * (k + m))(*[None if (i < dd or k <= i < k + pd) else files[i].ctypes.data for i in range(k + m)])
            sec = timeit(lambda: L.mec_encode(c.h, src.ctypes.data, size, fp, k + 1) == size or sys.exit("encode failed"))
            rows.append({"bench": "BenchmarkErasure" + name, "k": k, "m": m, "size": size, "drives": pattern(k, m, dd, pd), "MB_per_s": size / sec / 1e6,
                         "us_per_op": sec * 1e6, "cpu_1thread_MB_per_s": cpu})
            print(json.dumps(rows[-1]), flush=True)
        c.close()
        for a in files + [src]:
            L.mec_free_pinned(a.ctypes.data)
    for name, k, m, size, pats in DEC:
        c = mb.Codec(k, m, MiB)
        c.set_option("jit", 1)
        src = mb.capi.pinned_array(size, device=0)
        src[:] = np.random.default_rng(size % 89).integers(0, 256, size, dtype=np.uint8)
        fsz = c.bitrot_file_size(size)
        files = [mb.capi.pinned_array(fsz, device=0) for _ in range(k + m)]
        fp = (C.c_void_p * (k + m))(*[f.ctypes.data for f in files])
        assert L.mec_encode(c.h, src.ctypes.data, size, fp, k + 1) == size
        dst = mb.capi.pinned_array(size, device=0)
        hint = C.c_int(0)
        for dd, pd in pats:
            rp = (C.c_void_p * (k + m))(*[None if (i < dd or k <= i < k + pd) else files[i].ctypes.data for i in range(k + m)])
            sec = timeit(lambda: L.mec_decode(c.h, rp, 0, size, size, dst.ctypes.data, C.byref(hint)) == size or sys.exit("decode failed"))
            ok = bool(np.array_equal(dst, src))
            rows.append({"bench": "BenchmarkErasure" + name, "k": k, "m": m, "size": size, "drives": pattern(k, m, dd, pd), "MB_per_s": size / sec / 1e6,
                         "us_per_op": sec * 1e6, "bit_exact": ok})
            print(json.dumps(rows[-1]), flush=True)
        c.close()
        for a in files + [src, dst]:
            L.mec_free_pinned(a.ctypes.data)
    print("| benchmark | drives | object | MB/s (GPU path, one caller) | us/op | CPU 1 thread MB/s |\n|---|---|---|---|---|---|", file=sys.stderr)
    for r in rows:
        print("| %s | `%s` | %s | %.0f | %.0f | %s |" % (r["bench"], r["drives"], ("%d KiB" % (r["size"] >> 10)) if r["size"] < MiB else ("%d MiB" % (r["size"] >> 20)),
                                                      r["MB_per_s"], r["us_per_op"], ("%.0f" % r["cpu_1thread_MB_per_s"]) if r.get("cpu_1thread_MB_per_s") else "—"), file=sys.stderr)


if __name__ == "__main__":
    main()
