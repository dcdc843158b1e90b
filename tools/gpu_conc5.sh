#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for spec in "bget 16 1 200" "bget 64 1 100" "bget 256 1 40" "bget 64 256k 200" "batcher 16 1 200" "batcher 16 256k 300"; do
  timeout 300 ./tools/conc_bench $spec | tee -a gpurun_out/conc5.jsonl
done
