"""The compile-time GF(2^8) plans the kernels instantiate (bit-plane Horner + four-Russians groups + subset-sum change of
basis, ec_device.cuh) run on the HOST against a plain table multiply: every compile-time (k, m), several decode matrices,
every transform level and group size.  No GPU needed — the binary is built by __graft_entry__.build()."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_gf_plans_match_table_multiply():
    exe = os.path.join(HERE, "cpp", "test_gfplan")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/test_gfplan not built (run __graft_entry__.build())")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "all GF plans match the table multiply" in out.stdout
    # the headline geometry must have picked the subset-sum transform
    line = [l for l in out.stdout.splitlines() if l.startswith("encode(12,4)")][0]
    assert "level 2" in line, line
