"""Build libxlprop.so in-tree with nvcc for sm_100a (the only target).  Used by __graft_entry__.build().

One translation unit per kernel family (csrc/xl_core.cu, xl_rs.cu, xl_slab.cu, xl_czt.cu, xl_elements.cu), each compiled by its own nvcc
process WITHOUT -split-compile, so the binary is reproducible: identical sources and flags give identical per-kernel SASS
(round 1 used one unit with -split-compile 8, whose partitioning -- and with it every kernel's register allocation --
changed from run to run).  The units are compiled in parallel and linked into one shared object."""
import concurrent.futures
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
UNITS = ["xl_core.cu", "xl_rs.cu", "xl_slab.cu", "xl_czt.cu", "xl_elements.cu"]
OUT = os.path.join(HERE, "libxlprop.so")
OBJ_DIR = os.path.join(os.path.dirname(HERE), "build", "obj")

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC"]


def deps():
    return sorted(glob.glob(os.path.join(CSRC, "*"))) + [os.path.join(os.path.dirname(HERE), "include", "xlprop.h")]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in deps())


def _compile(unit, obj, defines, verbose):
    cmd = [_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines if d] + (["-Xptxas", "-v"] if verbose else []) + \
          ["-c", os.path.join(CSRC, unit), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return unit, r


def build(force=False, verbose=False, defines=(), out=None, units=None):
    """Default: the product library, in-tree.  `defines` + `out` build a variant of the same sources somewhere else
    (build/...) for A/B timing with XLPROP_LIB; the product library is never a variant."""
    if defines and not out:
        raise ValueError("a variant needs its own output path")
    out = out or OUT
    if out == OUT and not force and up_to_date():
        return OUT
    defines = list(defines) + (["XL_DEV_FAST"] if os.environ.get("XL_FAST") else [])   # development only: L in {2048, 4096}
    tag = "_".join(sorted(defines)) or "product"
    odir = os.path.join(OBJ_DIR, tag)
    os.makedirs(odir, exist_ok=True)
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    objs = {u: os.path.join(odir, u.replace(".cu", ".o")) for u in UNITS}
    todo = [u for u in UNITS if units is None or u in units or not os.path.exists(objs[u])]
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        for unit, r in ex.map(lambda u: _compile(u, objs[u], defines, verbose), todo):
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"nvcc failed compiling {unit}")
            if verbose:
                sys.stderr.write(r.stderr)
    r = subprocess.run([_nvcc(), "-shared", "-o", out] + [objs[u] for u in UNITS], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libxlprop.so")
    return out


if __name__ == "__main__":
    # python -m xlumina_b200.build [--force] [-v] [--exp MACRO[,MACRO...]] [--out build/libxlprop_<name>.so] [--units xl_rs.cu,...]
    argv = sys.argv[1:]
    exp = argv[argv.index("--exp") + 1].split(",") if "--exp" in argv else ()
    dst = argv[argv.index("--out") + 1] if "--out" in argv else None
    un = argv[argv.index("--units") + 1].split(",") if "--units" in argv else None
    print(build(force="--force" in argv, verbose="-v" in argv, defines=exp, out=dst, units=un))
