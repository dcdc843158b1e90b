#!/bin/bash
# round-2 kernel A/B: parity (default + warp-autonomous pipeline), then device-resident bench of every variant
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest gpu (default)"; timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 > $O/pytest_gpu.txt 2>&1; echo "rc=$?"; tail -2 $O/pytest_gpu.txt
echo "== pytest gpu (MEC_USE_AUTO=1)"; MEC_USE_AUTO=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 600 > $O/pytest_gpu_auto.txt 2>&1; echo "rc=$?"; tail -2 $O/pytest_gpu_auto.txt
bash tools/gpu_variants.sh 2>&1 | tee $O/variants.txt
