"""C-ABI shape checks that need no GPU: the library loads, exports every symbol include/rtb200.h
declares, and refuses to run without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import rtb200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "rtb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = C.CDLL(rtb200.device.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"librtb200.so does not export {n}"
    assert sorted(rtb200.device.ABI_SYMBOLS) == names


def test_record_sizes_match_header():
    assert rtb200.RAY_DTYPE.itemsize == 32 and rtb200.HIT_DTYPE.itemsize == 16


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rtb200.RtError, match="no CPU fallback"):
        rtb200.Context(0)


def test_product_does_not_touch_the_oracle():
    """nothing under the package may import, link or execute oracle/"""
    pkg = os.path.join(ROOT, "real-time-opencl-raytracer_b200")
    for base, _dirs, files in os.walk(pkg):
        if os.sep + "build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                src = open(os.path.join(base, f), errors="ignore").read()
                assert "liboracle" not in src and "oracle_py" not in src and "orc_" not in src, f"{f} references the oracle"
