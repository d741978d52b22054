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

# -split-compile > 1 partitions the translation unit for optimisation; the partitioning (and with it the register allocation
# of every kernel) is not reproducible from run to run -- two binaries have been seen from identical sources and flags
# (build note).  Compare per-kernel SASS hashes before attributing a timing change to a source change.
SPLIT_COMPILE = "8"
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


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


def build(force=False, verbose=False, defines=(), out=None, split=None):
    """Default: the product library, in-tree.  `defines` (e.g. ["XL_EXP_TREE_REDUCE"]) + `out` build an experiment variant
    of the same sources somewhere else (build/...) for A/B timing with XLPROP_LIB; the product library is never a variant."""
    if (defines or split) and not out:
        raise ValueError("an experiment variant needs its own output path")
    out = out or OUT
    if out == OUT and not force and up_to_date():
        return OUT
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    fast = ["-DXL_DEV_FAST"] if os.environ.get("XL_FAST") else []   # development only: L in {2048, 4096}
    cmd = [_nvcc()] + NVCC_FLAGS + ["-split-compile", str(split or SPLIT_COMPILE)] + fast + ["-D" + d for d in defines if d] + (["-Xptxas", "-v"] if verbose else []) + [SRC, "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libxlprop.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return out


if __name__ == "__main__":
    # python -m xlumina_b200.build [--force] [-v] [--exp MACRO[,MACRO...]] [--split N] [--out build/libxlprop_<name>.so]
    argv = sys.argv[1:]
    exp = argv[argv.index("--exp") + 1].split(",") if "--exp" in argv else ()
    dst = argv[argv.index("--out") + 1] if "--out" in argv else None
    spl = argv[argv.index("--split") + 1] if "--split" in argv else None
    print(build(force="--force" in argv, verbose="-v" in argv, defines=exp, out=dst, split=spl))
