"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every declared symbol."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "recgraph_b200", "csrc"), "-j8"],
                          stdout=subprocess.DEVNULL)
    from recgraph_b200 import _lib
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "recgraph_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", hdr))
    from recgraph_b200 import _lib
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for sym in declared:
        assert getattr(lib, sym) is not None


def test_struct_sizes_match_the_header(lib):
    from recgraph_b200 import _lib
    import ctypes
    assert ctypes.sizeof(_lib.Run) == 8
    assert ctypes.sizeof(_lib.ReadResult) == 80
    assert ctypes.sizeof(_lib.Scoring) == 36 * 4 + 8 * 4


def test_no_device_fails_loudly(lib):
    """Without a GPU the product must refuse to run (no CPU fallback)."""
    import ctypes
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = ctypes.c_void_p()
    assert lib.rg_init(0, ctypes.byref(ctx)) == -3  # RG_ERR_NO_DEVICE
    from recgraph_b200 import run_cli
    ex = os.path.join(ROOT, "tests", "golden", "example")
    rc, out, err = run_cli(["-m", "2", os.path.join(ex, "reads.fa"), os.path.join(ex, "graph.gfa")])
    assert rc != 0 and out == ""


def test_product_does_not_touch_the_oracle():
    """Nothing under recgraph_b200/ may import, include or link oracle/."""
    pkg = os.path.join(ROOT, "recgraph_b200")
    for d, _dirs, files in os.walk(pkg):
        if "_build" in d or "__pycache__" in d:
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", "Makefile")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f == "synth.py", os.path.join(d, f)
