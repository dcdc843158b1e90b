#!/bin/bash
# the concurrency table of profiles/r2_concurrency.md in one run (raw lines: profiles/r2_conc_final.jsonl)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/conc_final.jsonl
for spec in "batcher 256 1 60" "batcher 64 1 120" "batcher 16 1 200" "batcher 1024 1 16" "batcher 256 4 16" "batcher 256 16 6" "batcher 256 256k 100" "batcher 256 64k 200" "batcher 64 256k 200" \
            "pool 8 1 200 8" "pool 256 1 40 8" "pool 8 256k 300 8" "pool 256 16 6 8" "cpu 128 1 40" "cpu 128 256k 200" \
            "bget 256 1 40" "bget 64 1 100" "bget 16 1 200" "bget 256 4 16" "bget 256 16 6" "bget 256 256k 100" "get 8 1 200 8" "get 256 1 40 8" "get 8 256k 300 8" "get 256 16 6 8"; do
  timeout 300 ./tools/conc_bench $spec | tee -a gpurun_out/conc_final.jsonl
done
