#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${N:-8}
nvidia-smi topo -m 2>/dev/null | head -14
lscpu | grep -i "numa\|socket\|model name" | head -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.txt 2> gpurun_out/bench_n$N.err; echo "rc=$?"
tail -c 7000 gpurun_out/bench_n$N.txt; tail -5 gpurun_out/bench_n$N.err
timeout 600 python bench.py --impl reference --gpus $N --steps 5 --warmup 2 2>/dev/null | tail -1 | cut -c1-300
