#!/bin/bash
# Runs on the GPU box under gpurun: staged so that one hanging stage cannot eat the whole call.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread" > $O/host_cpu.txt 2>&1
echo "== ubench"; timeout 300 tools/ubench > $O/ubench.txt 2>&1; echo "rc=$?"; tail -15 $O/ubench.txt
echo "== smoke bytewise"; MEC_FORCE_BYTEWISE=1 timeout 300 python __graft_entry__.py smoke > $O/smoke_bytewise.txt 2>&1; echo "rc=$?"; tail -5 $O/smoke_bytewise.txt
echo "== smoke tma"; timeout 300 python __graft_entry__.py smoke > $O/smoke_tma.txt 2>&1; echo "rc=$?"; tail -5 $O/smoke_tma.txt
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu -x --timeout 600 > $O/pytest_gpu.txt 2>&1; echo "rc=$?"; tail -15 $O/pytest_gpu.txt
echo "== bench small"; timeout 600 python bench.py --blocks 2048 --steps 3 --warmup 3 > $O/bench_small.txt 2>&1; echo "rc=$?"; tail -3 $O/bench_small.txt
echo "== bench full"; timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_full.txt 2>&1; echo "rc=$?"; tail -3 $O/bench_full.txt
for eb in 1 2 4; do for gm in 0; do
echo "== bench eb=$eb"; MEC_EB=$eb timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > $O/bench_eb$eb.txt 2>&1; echo "rc=$?"; tail -1 $O/bench_eb$eb.txt | cut -c1-400
done; done
