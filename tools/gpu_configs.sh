#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 2>&1 | tail -3
timeout 900 python tools/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
python - <<'PY'
import json
for l in open('gpurun_out/configs.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print("%-60s bs=%8d  %.0f GiB/s  frac=%.3f ok=%s generic=%s first=%s" % (d["config"][:60], d["block_size"], d["GiB_per_s_object"], d["frac_of_hbm_peak"], d.get("bit_exact_vs_encode"), d.get("generic_GiB_per_s"), d.get("first_call_seconds_incl_nvrtc")))
PY
tail -5 gpurun_out/configs.err
