#!/bin/bash
# ncu capture of one bench_configs case (CASE=4k|3b|3a): the NVRTC-specialised reconstruct kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
CASE=${CASE:-4k}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fused_rs_hh -s ${SKIP:-24} -c 1 -o $O/prof_$CASE python tools/bench_configs.py $CASE > $O/prof_${CASE}_run.log 2>&1; echo "rc=$?"
tail -3 $O/prof_${CASE}_run.log
ls -la $O/prof_$CASE.ncu-rep
