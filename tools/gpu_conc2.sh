#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/conc2.jsonl
for spec in "pool 256 1 40 8" "pool 8 1 200 8" "pool 32 1 100 32" "pool 256 16 6 8" "get 256 1 40 8" "get 8 1 200 8" "get 256 16 6 8" "batcher 256 1 60" "batcher 64 1 120" "batcher 16 1 200" "batcher 1024 1 16" "batcher 256 4 16" "batcher 256 16 6" "bget 256 1 40" "bget 64 1 100" "bget 16 1 200" "bget 256 4 16" "bget 256 16 6"; do
  timeout 300 ./tools/conc_bench $spec | tee -a gpurun_out/conc2.jsonl
done
