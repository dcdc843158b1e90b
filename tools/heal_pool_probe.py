#!/usr/bin/env python
"""Config 4 host leg (mec_heal_batch, 128 x 64 MiB RS(16,4) objects) for several pool sizes / kernel choices."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import minio_b200 as mb
mb.capi.lib().mec_bind_thread_to_device(0)
for pool, jit in ((3, 1), (6, 1)):
    os.environ["MEC_HEAL_POOL"] = str(pool); os.environ["MEC_HEAL_JIT"] = str(jit)
    r = bench.heal_batch_leg(mb, 0, 128, lambda: None, lambda x: x, 1)
    print(json.dumps({"pool": pool, "jit": jit, "GiB_per_s": round(r["value"], 2), "ok": r["bit_exact_vs_encode"]}), flush=True)
