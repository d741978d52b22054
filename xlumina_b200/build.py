"""Build libxlprop.so in-tree with nvcc for sm_100a (the only target).  Used by __graft_entry__.build()."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "xl_api.cu")
DEPS = [os.path.join(HERE, "csrc", f) for f in ("xl_api.cu", "xl_kernels.cuh", "xl_long.cuh", "xl_fft.cuh", "xl_platform.h")] + \
       [os.path.join(os.path.dirname(HERE), "include", "xlprop.h")]
OUT = os.path.join(HERE, "libxlprop.so")

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-split-compile", "0"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    fast = ["-DXL_DEV_FAST"] if os.environ.get("XL_FAST") else []   # development only: L in {2048, 4096}
    cmd = [_nvcc()] + NVCC_FLAGS + fast + (["-Xptxas", "-v"] if verbose else []) + [SRC, "-o", OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libxlprop.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
