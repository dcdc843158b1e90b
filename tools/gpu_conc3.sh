#!/bin/bash
# objects smaller than one erasure block through the coalescer: one launch per batch (per-block geometry table) vs one per request
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/conc4.jsonl
for spec in "batcher 256 256k 100" "batcher 256 64k 200" "batcher 64 256k 200" "bget 256 256k 100" "batcher 256 1 60" "bget 256 1 40" "pool 8 256k 300 8" "get 8 256k 300 8" "cpu 128 256k 200"; do
  timeout 300 ./tools/conc_bench $spec | tee -a gpurun_out/conc4.jsonl
done
echo "--- one launch per request (MEC_SMALL_BLOCKS=0)"
for spec in "batcher 256 256k 100" "bget 256 256k 100"; do
  MEC_SMALL_BLOCKS=0 timeout 300 ./tools/conc_bench $spec | tee -a gpurun_out/conc4.jsonl
done
python -m pytest tests -q -m gpu -x -k "batcher" 2>&1 | tail -2
