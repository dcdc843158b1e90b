#!/bin/bash
# ncu counters of the reconstruct kernel (config 3a)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,smsp__warps_active.avg.per_cycle_active,gpu__time_duration.sum,launch__registers_per_thread,launch__block_size,launch__grid_size,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__shared_mem_per_block_dynamic,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:fused_rs_hh --csv --log-file gpurun_out/decode_ncu.csv python tools/bench_configs.py 3a 3552 > gpurun_out/decode_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/decode_ncu.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
seen={}
for r in rows[1:]:
    seen.setdefault(r[ii],{'k':r[ki][:60]})[r[mi]]=r[vi]
ids=sorted(seen,key=int)
for i in (ids[0], ids[len(ids)//2], ids[-1]):
    print(i, seen[i])
PY
