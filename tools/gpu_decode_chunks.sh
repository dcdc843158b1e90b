#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
cat > /tmp/dec.py <<'PY'
import sys, time, ctypes as C, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import minio_b200 as mb
L = mb.lib(); L.mec_bind_thread_to_device(0)
k, m, bs, nb = 12, 4, 1 << 20, 2048
size = nb * bs
c = mb.Codec(k, m, bs); c.set_option("jit", 1)
src = mb.capi.pinned_array(size, device=0); src[:] = np.random.default_rng(1).integers(0, 256, size, dtype=np.uint8)
fsz = c.bitrot_file_size(size)
files = [mb.capi.pinned_array(fsz, device=0) for _ in range(16)]
fp = (C.c_void_p * 16)(*[f.ctypes.data for f in files])
assert L.mec_encode(c.h, src.ctypes.data, size, fp, 13) == size
dst = mb.capi.pinned_array(size, device=0)
hint = C.c_int(0)
for erased in [(0, 1, 2, 3), (), (5,), (12, 13, 14, 15)]:
    rp = (C.c_void_p * 16)(*[None if i in erased else files[i].ctypes.data for i in range(16)])
    for chunk in (0, 8, 16, 64, 128):
        c.set_option("chunk_blocks", chunk)
        for _ in range(2): L.mec_decode(c.h, rp, 0, size, size, dst.ctypes.data, C.byref(hint))
        t0 = time.perf_counter(); n = 4
        for _ in range(n): assert L.mec_decode(c.h, rp, 0, size, size, dst.ctypes.data, C.byref(hint)) == size
        dt = (time.perf_counter() - t0) / n
        print("erased %-16s chunk %3d: %.1f GiB/s" % (erased, chunk, size / dt / 2**30), flush=True)
assert np.array_equal(dst, src)
PY
python /tmp/dec.py
