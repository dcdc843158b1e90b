"""Runs the C++ port of the reference's Go test tables (tests/cpp/test_erasure.cc) on the GPU."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_mirror_tables():
    exe = os.path.join(ROOT, "tests", "cpp", "test_erasure")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    print(r.stdout)
    print(r.stderr[-4000:])
    assert r.returncode == 0
    assert r.stdout.count("ok  ") == 6
