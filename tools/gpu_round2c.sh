#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -q -m gpu -x --timeout 900 > $O/pytest_gpu.txt 2>&1; echo "rc=$?"; tail -15 $O/pytest_gpu.txt
run() { echo -n "$1: "; env $2 timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('GiB/s %.1f frac %.3f'%(d['value'],d['roofline']['frac']))"; }
run base ""
run static "MEC_STATIC_GROUPS=1"
run semi "MEC_USE_AUTO=2"
run semi_static "MEC_USE_AUTO=2 MEC_STATIC_GROUPS=1"
