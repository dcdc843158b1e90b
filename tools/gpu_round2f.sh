#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 > $O/pytest_gpu.txt 2>&1; echo "rc=$?"; tail -6 $O/pytest_gpu.txt
python tools/bench_configs.py gen 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['config'][:50], 'generic %.0f GiB/s'%d.get('generic_GiB_per_s',0), 'jit %.0f'%d.get('GiB_per_s_object',0), d.get('bit_exact_vs_encode'))"
