#!/usr/bin/env python
"""Wall-clock latency of one small PutObject / degraded GetObject through the C ABI (pinned buffers, one caller)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minio_b200 as mb
from minio_b200 import capi


def main():
    k, m, bs = 12, 4, 1 << 20
    L = capi.lib()
    for size_mib in (1, 1.5, 4, 16):
        size = int(size_mib * (1 << 20)) + (4321 if size_mib != int(size_mib) else 0)
        for small in (-1, 0):
            c = mb.Codec(k, m, bs)
            c.set_option("small_blocks", small)
            fsz = c.bitrot_file_size(size)
            src = capi.pinned_array(size, 0)
            src[:] = np.random.default_rng(1).integers(0, 256, size, dtype=np.uint8)
            files = [capi.pinned_array(fsz, 0) for _ in range(k + m)]
            dd = capi.pinned_array((size // bs + 1) * k * 32, 0)
            ptrs = (C.c_void_p * (k + m))(*[f.ctypes.data for f in files])
            def put():
                rc = L.mec_encode_sg(c.h, src.ctypes.data, size, ptrs, dd.ctypes.data, 0)
                assert rc >= 0, rc
            def put_frames():
                rc = L.mec_encode(c.h, src.ctypes.data, size, ptrs, 0)
                assert rc >= 0, rc
            put_frames()
            dst = capi.pinned_array(size, 0)
            gptrs = (C.c_void_p * (k + m))(*[None if i < 4 else files[i].ctypes.data for i in range(k + m)])
            hint = C.c_int(0)
            def get():
                rc = L.mec_decode(c.h, gptrs, 0, size, size, dst.ctypes.data, C.byref(hint))
                assert rc == size, rc
            row = {"object_MiB": size_mib, "kernel": "latency" if small else "throughput"}
            for name, fn in (("put_sg_us", put), ("put_frames_us", put_frames), ("degraded_get_us", get)):
                for _ in range(10):
                    fn()
                ts = []
                for _ in range(100):
                    t0 = time.perf_counter()
                    fn()
                    ts.append((time.perf_counter() - t0) * 1e6)
                row[name] = round(float(np.median(ts)), 1)
            assert np.array_equal(dst, src)
            row["small_launches"] = c.stat("small_launches")
            print(json.dumps(row), flush=True)
            c.close()
            for f in files + [src, dd, dst]:
                L.mec_free_pinned(f.ctypes.data)


if __name__ == "__main__":
    main()
