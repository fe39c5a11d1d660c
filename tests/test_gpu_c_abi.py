"""The drop-in boundary exercised from plain C (tests/c/abi_demo.c): golden tables of the reference through host
Arrow-layout buffers, no Python between the caller and libbowgpu.so."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_consumer_runs(tmp_path):
    exe = tmp_path / "abi_demo"
    subprocess.check_call(["gcc", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_demo.c"),
                           "-o", str(exe), "-L", os.path.join(ROOT, "bow_b200"), "-lbowgpu",
                           "-Wl,-rpath," + os.path.join(ROOT, "bow_b200"), "-lm"])
    fixture = os.path.join(ROOT, "tests", "golden", "parquet", "refwriter_bow1_10_rows.parquet")
    p = subprocess.run([str(exe), fixture], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "aggregate: ok" in p.stdout and "interpolate: ok" in p.stdout and "sort: ok" in p.stdout
    assert "parquet: ok" in p.stdout
