"""The C-ABI library loads in the CPU container and exports every symbol include/bowgpu.h declares
(no compute calls here: there is no GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    from bow_b200 import native as N
    if not os.path.exists(N.LIB_PATH):
        g.build()
    return N.lib()


def header_functions():
    src = open(os.path.join(ROOT, "include", "bowgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bowgpu_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from bow_b200 import native as N
    names = header_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(N.SYMBOLS) == names, set(N.SYMBOLS) ^ set(names)


def test_abi_version_and_status_strings(lib):
    assert lib.bowgpu_abi_version() == 1
    assert lib.bowgpu_status_string(0) == b"ok"
    assert b"sorted" in lib.bowgpu_status_string(7)
    assert lib.bowgpu_agg_needs_inclusive(9) == 1 and lib.bowgpu_agg_needs_inclusive(8) == 0
    assert lib.bowgpu_agg_return_type(6, 2) == 2 and lib.bowgpu_agg_return_type(6, 1) == 1  # First: InputDependent
    assert lib.bowgpu_agg_return_type(1, 1) == 2 and lib.bowgpu_agg_return_type(2, 2) == 1  # Count int64, Sum float64


def test_struct_layouts_match_header():
    from bow_b200 import native as N
    assert C.sizeof(N.Col) == 48 and C.sizeof(N.AggSpec) == 48 and C.sizeof(N.OutCol) == 24
    assert C.sizeof(N.Timing) == 16 and C.sizeof(N.GenSpec) == 64


def test_no_cpu_fallback_without_a_gpu(lib):
    """On a box without a GPU the product path must fail loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from bow_b200 import native as N
    with pytest.raises(N.BowGpuError):
        N.Ctx(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no CPU fallback", ""), os.path.join(dirpath, f)


def test_c_consumer_compiles_and_links(tmp_path):
    """the boundary is usable from plain C: tests/c/abi_demo.c builds against include/bowgpu.h and links the library
    (it is RUN by the GPU suite, tests/test_gpu_c_abi.py)"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "abi_demo"
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "c", "abi_demo.c"), "-o", str(exe), "-L", os.path.join(root, "bow_b200"),
                           "-lbowgpu", "-Wl,-rpath," + os.path.join(root, "bow_b200"), "-lm"])
    assert exe.exists()
