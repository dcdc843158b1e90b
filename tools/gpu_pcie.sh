#!/bin/bash
# what the PCIe link gives (pinned H2D / D2H, alone and together) next to the end-to-end encode
cd "${GRAFT_REPO_ROOT:-/root/repo}"
python - <<'PY'
import torch, time
dev = torch.device("cuda:0")
n = 2 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device=dev)
h2 = torch.empty(n // 3, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(n // 3, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
a = t(lambda: d.copy_(h, non_blocking=True)); print("H2D alone   %.1f GB/s" % (n / a / 1e9))
b = t(lambda: h2.copy_(d2, non_blocking=True)); print("D2H alone   %.1f GB/s" % (n // 3 / b / 1e9))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both); print("H2D 2 GiB + D2H 0.67 GiB together: %.1f GB/s H2D-equivalent" % (n / c / 1e9))
PY
for cb in 64 128 256 512; do echo -n "chunk_blocks=$cb: "; MEC_CHUNK_BLOCKS=$cb python bench.py --steps 3 --warmup 2 --no-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('e2e %.1f GiB/s'%d['e2e']['value'])"; done
