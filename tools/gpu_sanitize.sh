#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import minio_b200 as mb, oracle_lib as o
rng = np.random.default_rng(3)
for (k, m, bs, n) in [(12, 4, 1 << 20, (1 << 20) + 4321), (8, 4, 1 << 20, 2 * (1 << 20) + 999), (4, 2, 65536, 3 * 65536 + 17), (7, 5, 1 << 20, 600000)]:
    data = rng.integers(0, 256, n, dtype=np.uint8)
    c = mb.Codec(k, m, bs)
    files = c.encode(data)
    want, _ = o.erasure_encode(k, m, bs, 3, data)
    assert all(np.array_equal(a, b) for a, b in zip(files, want))
    off = [None if i in (0, k) else files[i] for i in range(k + m)]
    out, hint = c.decode(off, 0, n, n)
    assert np.array_equal(out, data)
    for jit in (0, 1):
        c.set_option("jit", jit)
        healed = c.heal(off, [i in (0, k) for i in range(k + m)], n)
        assert np.array_equal(healed[0], files[0]) and np.array_equal(healed[k], files[k])
    assert c.bitrot_verify(files[1], c.shard_file_size(n)) == 0
    c.close()
c = mb.Codec(8, 8, 65536, algo=1)
f, s = c.encode_whole(rng.integers(0, 256, 200000, dtype=np.uint8))
print("sanitizer case ok")
PY
echo "== memcheck"; timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san_case.py > gpurun_out/memcheck.txt 2>&1; echo rc=$?; grep -E "ERROR SUMMARY|sanitizer case ok|Invalid|out of bounds|misaligned" gpurun_out/memcheck.txt | head -12
echo "== racecheck"; timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/san_case.py > gpurun_out/racecheck.txt 2>&1; echo rc=$?; grep -E "RACECHECK SUMMARY|sanitizer case ok|hazard" gpurun_out/racecheck.txt | head -12
echo "== synccheck"; timeout 1200 compute-sanitizer --tool synccheck --print-limit 20 python /tmp/san_case.py > gpurun_out/synccheck.txt 2>&1; echo rc=$?; grep -E "ERROR SUMMARY|sanitizer case ok|Barrier" gpurun_out/synccheck.txt | head -8
