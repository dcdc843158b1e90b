#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for spec in "batcher 256 1 40" "batcher 64 1 80" "batcher 1024 1 16" "batcher 256 4 16"; do timeout 300 ./tools/conc_bench $spec; done
echo "== one copy stream"; for spec in "batcher 256 1 40" "batcher 64 1 80"; do MEC_BATCHER_ONE_COPY_STREAM=1 timeout 300 ./tools/conc_bench $spec; done
