#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for spec in "batcher 256 1 40" "batcher 1024 1 16" "batcher 256 16 6" "batcher 64 1 80" "get 256 1 40 8" "get 256 16 6 8"; do timeout 300 ./tools/conc_bench $spec; done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-configs 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print({k:(round(d[k]['value'],1) if d.get(k) else None) for k in ('e2e','e2e_frames','e2e_decode')}, round(d['value'],1))"
