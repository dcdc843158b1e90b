#!/bin/bash
# latency kernel: kernel time of small launches, then the parity suite (test_gpu_parity.py runs in both kernel forms)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python tools/bench_small.py 12 4 2>&1 | tee gpurun_out/small_12_4.jsonl
python tools/bench_small.py 16 4 2>&1 | tee gpurun_out/small_16_4.jsonl
python tools/bench_small.py 8 8 2>&1 | tee gpurun_out/small_8_8.jsonl
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest_gpu.txt 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.txt
for spec in "batcher 256 1 60" "batcher 64 1 120" "batcher 16 1 200" "pool 256 1 40 8" "pool 8 1 200 8" "bget 256 1 40" "bget 64 1 100" "get 256 1 40 8"; do
  timeout 300 ./tools/conc_bench $spec | tee -a gpurun_out/conc_small.jsonl
done
