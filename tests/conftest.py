import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture(scope="session")
def emu():
    """Host emulation of the kernel bodies (same sources compiled with g++ -DXL_HOST_EMU).  Test tooling: lets the index
    math, butterflies and fused factors of every kernel be checked against the oracle without a GPU.  Never loaded by the
    xlumina_b200 package."""
    from xlumina_b200 import _lib
    out = os.path.join(ROOT, "tests", "emu", "libxlprop_emu.so")
    src = os.path.join(ROOT, "xlumina_b200", "csrc", "xl_api.cu")
    import glob
    deps = glob.glob(os.path.join(ROOT, "xlumina_b200", "csrc", "*"))
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-x", "c++", "-std=c++17", "-O2", "-DXL_HOST_EMU", "-shared", "-fPIC", "-w", src, "-o", out])
    return _lib.declare(ctypes.CDLL(out))


def emu_variant_path(macros):
    """Host emulation of an experiment variant (-DXL_EXP_...), built once per source state under tests/emu/."""
    tag = "_".join(m.replace("XL_EXP_", "") for m in macros)
    out = os.path.join(ROOT, "tests", "emu", f"libxlprop_emu_{tag}.so")
    src = os.path.join(ROOT, "xlumina_b200", "csrc", "xl_api.cu")
    import glob
    deps = glob.glob(os.path.join(ROOT, "xlumina_b200", "csrc", "*"))
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-x", "c++", "-std=c++17", "-O1", "-DXL_HOST_EMU"] + ["-D" + m for m in macros] +
                              ["-shared", "-fPIC", "-w", src, "-o", out])
    return out
