#!/usr/bin/env python
"""Deep scan over many small part files: mec_bitrot_verify_batch of N files of one short frame each (objects of ~1 MiB with odd sizes)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minio_b200 as mb
from minio_b200 import capi

k, m, bs = 12, 4, 1 << 20
c = mb.Codec(k, m, bs)
rng = np.random.default_rng(3)
nobj = 256
files, plens = [], []
arena = capi.pinned_array(nobj * 16 * (32 + c.shard_size()) + 4096, 0)
off = 0
for o in range(nobj):
    size = int(rng.integers(300_000, 1_048_000))
    data = rng.integers(0, 256, size, dtype=np.uint8)
    fs = c.encode(data)
    for f in fs:
        arena[off:off + f.size] = f
        files.append(arena[off:off + f.size]); plens.append(c.shard_file_size(size))
        off += (f.size + 15) // 16 * 16
total = sum(f.size for f in files)
for small in (-1, 0):
    c.set_option("small_blocks", small)
    for _ in range(2):
        res = c.bitrot_verify_batch(files, plens)
    assert all(r == 0 for r in res)
    l0 = c.launches
    t0 = time.perf_counter()
    for _ in range(5):
        res = c.bitrot_verify_batch(files, plens)
    dt = (time.perf_counter() - t0) / 5
    print(json.dumps({"files": len(files), "bytes": total, "one_launch_per_chunk": small != 0, "GiB_per_s": round(total / dt / 2**30, 2),
                      "files_per_s": round(len(files) / dt), "launches_per_scan": (c.launches - l0) // 5}), flush=True)
bad = files[777].copy(); bad[100] ^= 1
res = c.bitrot_verify_batch(files[:777] + [bad] + files[778:], plens)
assert res[777] == -7 and sum(1 for r in res if r) == 1
print("ok")
