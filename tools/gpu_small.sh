#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python tools/bench_small.py 12 4 2>&1 | grep -E "\"blocks\": (1|148|592|888)," | tee gpurun_out/small_12_4_b.jsonl
python tools/bench_small.py 16 4 2>&1 | grep -E "\"blocks\": (1|148|592)," | tee gpurun_out/small_16_4_b.jsonl
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest_gpu.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.txt
