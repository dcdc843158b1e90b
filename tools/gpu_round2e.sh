#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 > $O/pytest_gpu.txt 2>&1; echo "rc=$?"; tail -6 $O/pytest_gpu.txt
for c in 3a 3b 4k; do python tools/bench_configs.py $c 8192 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['config'][:40], 'GiB/s %.0f frac %.3f generic %.0f ok=%s'%(d['GiB_per_s_object'],d['frac_of_hbm_peak'],d['generic_GiB_per_s'],d['bit_exact_vs_encode']))"; done
echo "== no runs"; for c in 3a 4k; do MEC_NO_ROWS3D=1 python tools/bench_configs.py $c 8192 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['config'][:40], 'GiB/s %.0f frac %.3f generic %.0f ok=%s'%(d['GiB_per_s_object'],d['frac_of_hbm_peak'],d['generic_GiB_per_s'],d['bit_exact_vs_encode']))"; done
