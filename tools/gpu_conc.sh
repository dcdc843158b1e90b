#!/bin/bash
# the many-small-concurrent-requests regime on one GPU (profiles/r2_concurrency.md): coalesced vs pooled PUTs and degraded GETs, CPU per-request path
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/conc.jsonl
for spec in "batcher 256 1 40" "pool 256 1 40 8" "cpu 128 1 40" "batcher 64 1 80" "batcher 1024 1 16" "batcher 256 4 16" "batcher 256 16 6" "pool 256 16 6 8" "cpu 128 16 6" \
            "bget 256 1 40" "get 256 1 40 8" "bget 256 4 16" "bget 256 16 6" "get 256 16 6 8"; do
  timeout 300 ./tools/conc_bench $spec | tee -a gpurun_out/conc.jsonl
done
