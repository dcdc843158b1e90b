#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest gpu (MEC_USE_AUTO=2)"; MEC_USE_AUTO=2 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 600 > $O/pytest_gpu_semi.txt 2>&1; echo "rc=$?"; tail -2 $O/pytest_gpu_semi.txt
run() { echo -n "$1: "; env $2 timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('GiB/s %.1f frac %.3f'%(d['value'],d['roofline']['frac']))"; }
run base ""
run semi "MEC_USE_AUTO=2"
run auto "MEC_USE_AUTO=1"
for v in tools/variants/libmec_*.so; do run $(basename $v) "MEC_LIB=$PWD/$v"; run "$(basename $v)+semi" "MEC_LIB=$PWD/$v MEC_USE_AUTO=2"; done
