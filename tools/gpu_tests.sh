#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 ${PYTEST_ARGS} > gpurun_out/pytest_gpu.txt 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu.txt
