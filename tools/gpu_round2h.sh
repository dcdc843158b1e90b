#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 > gpurun_out/pytest_gpu.txt 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.txt
for i in 1 2; do ./tools/conc_bench batcher 256 1 40 | cut -c1-210; MEC_BATCHER_GATHER=1 ./tools/conc_bench batcher 256 1 40 | cut -c1-210; done
MEC_BATCHER_GATHER=1 ./tools/conc_bench batcher 64 1 80 | cut -c1-210
MEC_BATCHER_GATHER=1 ./tools/conc_bench batcher 256 4 16 | cut -c1-210
MEC_BATCHER_GATHER=1 ./tools/conc_bench batcher 256 16 6 | cut -c1-210
