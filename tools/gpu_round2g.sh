#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest gpu (MEC_USE_AUTO=3)"; MEC_USE_AUTO=3 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 600 2>&1 | tail -2
run() { echo -n "$1: "; env $2 timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-configs 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('GiB/s %.1f frac %.3f'%(d['value'],d['roofline']['frac']))"; }
run base ""
run early "MEC_USE_AUTO=3"
run semi "MEC_USE_AUTO=2"
run early2 "MEC_USE_AUTO=3"
run base2 ""
