#!/bin/bash
# ncu evidence: launch list + one full capture of the fused kernel (1 GPU; never under a multi-rank launch)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
CMD="python bench.py --blocks ${BLOCKS:-2048} --steps 2 --warmup 3 --no-e2e --no-cpu"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches.csv $CMD > $O/launches_run.log 2>&1; echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fused_rs_hh -s 3 -c 1 -o $O/prof $CMD > $O/prof_run.log 2>&1; echo "full rc=$?"
ls -la $O/
