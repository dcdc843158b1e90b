#!/bin/bash
# quick iteration on the GPU box: parity tests, device-resident bench, a handful of ncu counters
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 > $O/pytest_gpu.txt 2>&1; echo "rc=$?"; tail -4 $O/pytest_gpu.txt
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $O/bench_q.txt 2>&1; echo "rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_q.txt').read().strip().split('\n')[-1])
    print("GiB/s %.1f  ms %.3f  frac %.3f  verified %s clocks %s"%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['verified_vs_oracle'],d['clocks']))
except Exception as e: print("bench parse failed", e); print(open('gpurun_out/bench_q.txt').read()[-2000:])
PY
for v in $EXTRA_ENVS; do echo "== bench $v"; env $v timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('GiB/s %.1f frac %.3f'%(d['value'],d['roofline']['frac']))"; done
M=smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,smsp__warps_active.avg.per_cycle_active,smsp__warps_eligible.avg.per_cycle_active,gpu__time_duration.sum,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none -k regex:fused_rs_hh -s 3 -c 1 --csv --log-file $O/quick_ncu.csv python bench.py --blocks 4144 --steps 1 --warmup 3 --no-e2e --no-cpu > $O/quick_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
try:
    rows=[r for r in csv.reader(open('gpurun_out/quick_ncu.csv')) if len(r)>10]
    for r in rows[1:]: print("%-90s %s %s"%(r[-3],r[-1],r[-2]))
except Exception as e: print("ncu parse failed",e)
PY
