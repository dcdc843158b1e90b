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
d = rng.integers(0, 256, 200000, dtype=np.uint8)
f, s = c.encode_whole(d)
out, hint = c.decode_whole([None, None] + f[2:], s, 5, 150000, 200000)       # whole-file readers: verify + rebuild from raw shards
assert hint == 0 and np.array_equal(out, d[5:150005])
c.close()
# round 2: multi-chunk pipelined decode with fail-over (run-wise 3-D fetch from the arena), scatter-gather encode, CRCs, the batcher
k, m, bs, n = 12, 4, 1 << 20, 40 * (1 << 20) + 77
c = mb.Codec(k, m, bs); c.set_option("checksums", 7); c.set_option("chunk_blocks", 8)
data = rng.integers(0, 256, n, dtype=np.uint8)
files, dd = c.encode_sg(data)
crc, ln = c.last_checksums()
import zlib
assert ln == n and crc[0] == zlib.crc32(data.tobytes())
full = c.encode(data)
bad = [x.copy() for x in full]; bad[3][20 * (32 + c.shard_size()) + 40] ^= 1
out, hint = c.decode([None if i in (0, 9) else bad[i] for i in range(16)], 12345, n - 20000, n)
assert hint == -7 and np.array_equal(out, data[12345:n - 7655])
c.close()
import threading
bat = mb.Batcher(k, m, bs, max_batch_blocks=32, max_wait_us=500)
pin = [mb.pinned_array(sz)[:sz] for sz in ((1 << 20), (1 << 20) + 333, 2 * (1 << 20), 4097)]
for a in pin: a[:] = rng.integers(0, 256, a.size, dtype=np.uint8)
want = [o.erasure_encode(k, m, bs, 3, a)[0] for a in pin]
errs = []
def work(t):
    for _ in range(2):
        fs = bat.encode(pin[t % 4], pinned=(t < 4))
        if not all(np.array_equal(x, y) for x, y in zip(fs, want[t % 4])): errs.append(t)
th = [threading.Thread(target=work, args=(t,)) for t in range(8)]
[x.start() for x in th]; [x.join() for x in th]
assert not errs, errs
bat.close()
print("sanitizer case ok")
PY
echo "== memcheck"; timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san_case.py > gpurun_out/memcheck.txt 2>&1; echo rc=$?; grep -E "ERROR SUMMARY|sanitizer case ok|Invalid|out of bounds|misaligned" gpurun_out/memcheck.txt | head -12
echo "== racecheck"; timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/san_case.py > gpurun_out/racecheck.txt 2>&1; echo rc=$?; grep -E "RACECHECK SUMMARY|sanitizer case ok|hazard" gpurun_out/racecheck.txt | head -12
echo "== synccheck"; timeout 1200 compute-sanitizer --tool synccheck --print-limit 20 python /tmp/san_case.py > gpurun_out/synccheck.txt 2>&1; echo rc=$?; grep -E "ERROR SUMMARY|sanitizer case ok|Barrier" gpurun_out/synccheck.txt | head -8
