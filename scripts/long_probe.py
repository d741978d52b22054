"""Per-kernel CUDA-event times (xl_prof_*) of the split-line / slab chain on ONE GPU (csrc/xl_long.cuh), fresh z and with the
transfer function reused.   usage: python scripts/long_probe.py [N ...]      (default 16384)"""
import sys, os, ctypes, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xlumina_b200 as xb
from xlumina_b200 import slab, _lib

dev = torch.device("cuda:0")
L = _lib.lib()
lam, z = 0.6328, 5.0e4
k = 2 * math.pi / lam
buf = ctypes.create_string_buffer(1 << 16)
for N in [int(a) for a in sys.argv[1:]] or [16384]:
    x, _ = xb.space(15000.0, N)
    dx = float(x[1] - x[0])
    g = torch.Generator(device="cpu").manual_seed(1)
    u = torch.view_as_complex(torch.randn(N, N, 2, generator=g)).to(dev)
    out, H = slab.rs_propagation_slab(u, z, dx, dx, k, return_transfer=True, group=slab._LOCAL)
    del out
    torch.cuda.synchronize()
    iters = 3
    for what in ("fresh", "reused"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            slab.rs_propagation_slab(u, z, dx, dx, k, transfer=H if what == "reused" else None, group=slab._LOCAL)
        e1.record(); torch.cuda.synchronize()
        total = e0.elapsed_time(e1) / iters
        L.xl_prof_enable(1)
        for _ in range(iters):
            slab.rs_propagation_slab(u, z, dx, dx, k, transfer=H if what == "reused" else None, group=slab._LOCAL)
        torch.cuda.synchronize()
        L.xl_prof_report(buf, len(buf)); L.xl_prof_enable(0)
        print(f"--- N={N} {what}: {total:9.3f} ms per call (unprofiled)")
        ksum = 0.0
        for ln in buf.value.decode().strip().splitlines():
            nm, cnt, tot = ln.split()
            per_call = float(tot) / iters
            ksum += per_call
            print(f"    {nm:18s} {int(cnt) // iters:3d} launches  {per_call:9.3f} ms/call")
        print(f"    kernels sum {ksum:9.3f} ms", flush=True)
    del u, H
    torch.cuda.empty_cache()
