#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { echo -n "$1: "; env $2 timeout 300 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('GiB/s %.1f frac %.3f ok=%s'%(d['value'],d['roofline']['frac'],d['verified_vs_oracle']))"; }
run base ""
for v in tools/variants/libmec_*.so; do run $(basename $v) "MEC_LIB=$PWD/$v"; done
run auto "MEC_USE_AUTO=1"
run eb2 "MEC_EB=2"
run eb1 "MEC_EB=1"
run gm5 "MEC_GRID_MULT=5"
run gm4 "MEC_GRID_MULT=4"
run gm3 "MEC_GRID_MULT=3"
