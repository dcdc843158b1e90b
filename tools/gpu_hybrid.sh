#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/hybrid.jsonl
export MEC_BATCHER_TRACE=1
run() { echo "{\"env\": \"$1\"}" | tee -a gpurun_out/hybrid.jsonl; shift; env "$@" | tee -a gpurun_out/hybrid.jsonl; }
for spec in "bget 256 1 40" "bget 256 4 16" "bget 256 16 6" "bget 64 1 100"; do
  for mode in 0 1 2 0 1 2; do
    run "out$mode" MEC_BATCHER_GET_OUT=$mode timeout 300 ./tools/conc_bench $spec
  done
done
for spec in "batcher 256 1 60" "batcher 1024 1 16"; do
  run "put" timeout 300 ./tools/conc_bench $spec
  run "put" timeout 300 ./tools/conc_bench $spec
done
for mode in 0 1 2; do MEC_BATCHER_GET_OUT=$mode python -m pytest tests -m gpu -x -q -k "batcher" 2>&1 | tail -1; done
