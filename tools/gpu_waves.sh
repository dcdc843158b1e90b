#!/bin/bash
# does wave quantisation of the persistent grid show up?  bench at block counts around whole waves
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for nb in 4440 8880 10240 13320 14000 17760; do
  timeout 300 python bench.py --blocks $nb --steps 5 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('blocks $nb GiB/s %.1f frac %.3f ms %.3f'%(d['value'],d['roofline']['frac'],d['ms_per_step']))"
done
