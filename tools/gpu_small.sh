#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest_gpu.txt 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.txt
python tools/lat_probe.py 2>&1 | tee gpurun_out/lat_probe.jsonl
for spec in "pool 8 1 200 8" "pool 256 1 40 8" "get 8 1 200 8" "get 256 1 40 8" "get 1 1 400 1" "pool 1 1 400 1"; do
  timeout 300 ./tools/conc_bench $spec | tee -a gpurun_out/conc3.jsonl
done
