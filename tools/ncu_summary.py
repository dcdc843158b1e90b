#!/usr/bin/env python
"""Turns an `ncu --set full` report into the markdown table kept under profiles/ (run where ncu is installed; no GPU needed).
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep "<title>" "<command>" > profiles/<name>.md"""
import csv
import io
import subprocess
import sys

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__block_size",
        "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]


def main():
    rep, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    print(f"# {title}\n\nCommand: `{cmd}`.\nNumbers under a profiler are never bench values; this file is evidence for DRAM traffic and pipe utilisation.\n")
    print(f"Kernel: `{m.get('Kernel Name', ('', '?'))[1]}`\n\n| metric | unit | value |\n|---|---|---|")
    for k in KEEP:
        if k in m:
            print(f"| {k} | {m[k][0]} | {m[k][1]} |")
    for k in sorted(m):
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
            print(f"| {k} | {m[k][0]} | {m[k][1]} |")


if __name__ == "__main__":
    main()
