#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu -x --timeout 600 -k "batcher" 2>&1 | tail -4
for spec in "bget 256 1 40" "get 256 1 40 8" "bget 64 1 80" "bget 256 4 16" "bget 256 16 6" "get 256 16 6 8" "batcher 256 1 40"; do echo "## $spec"; timeout 300 ./tools/conc_bench $spec 2>&1; echo "rc=$?"; done
