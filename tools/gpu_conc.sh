#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu -x --timeout 600 -k "batcher" 2>&1 | tail -5
: > $O/conc.jsonl
for spec in "batcher 256 1 40" "batcher 1024 1 16" "batcher 256 16 6" "batcher 64 1 80"; do
  timeout 300 ./tools/conc_bench $spec | tee -a $O/conc.jsonl
done
echo "== MAX_WAIT_US=0"; MAX_WAIT_US=0 ./tools/conc_bench batcher 256 1 40
echo "== MAX_BATCH=128"; MAX_BATCH=128 ./tools/conc_bench batcher 256 1 40
echo "== 512 threads"; ./tools/conc_bench batcher 512 1 30
