#!/bin/bash
# A/B of build variants (tools/build_variants.sh -> tools/variants/libmec_<name>.so) and run-time options on the device-resident headline shape
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { echo -n "$1: "; env $2 timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-configs 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('GiB/s %.1f frac %.3f'%(d['value'],d['roofline']['frac']))"; }
run base ""
for v in tools/variants/libmec_*.so; do [ -e "$v" ] && run $(basename $v) "MEC_LIB=$PWD/$v"; done
for e in $EXTRA_ENVS; do run "$e" "$e"; done
run base-again ""
