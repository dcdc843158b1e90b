#!/bin/bash
# persistent-grid shape: occupancy grid vs balanced passes vs fewer CTAs per SM, at a few stream lengths
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for nb in 10240 4096 16384; do
for v in "" MEC_BALANCE_GRID=1 MEC_GRID_MULT=6; do
  echo -n "blocks $nb [$v]: "; env $v timeout 300 python bench.py --blocks $nb --steps 5 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('GiB/s %.1f frac %.3f'%(d['value'],d['roofline']['frac']))"
done; done
