#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -q -m gpu -x --timeout 900 > $O/pytest_gpu.txt 2>&1; echo "rc=$?"; tail -12 $O/pytest_gpu.txt
echo "== bench full"; nproc; numactl -H 2>/dev/null | head -5; timeout 1500 python bench.py --steps 10 --warmup 3 > $O/bench_full.txt 2> $O/bench_full.err; echo "rc=$?"; tail -c 6000 $O/bench_full.txt; tail -20 $O/bench_full.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_ref.txt 2>&1; tail -c 1500 $O/bench_ref.txt
