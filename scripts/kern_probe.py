"""Per-kernel CUDA-event times (xl_prof_*) of each operator at 2048^2: development aid for A/B work.
usage: python scripts/kern_probe.py [iters]      (XLPROP_LIB selects the library)"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import xlumina_b200 as xb
from xlumina_b200 import ops, _lib

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
N = 2048
dev = torch.device("cuda:0")
L = _lib.lib()
x, y = xb.space(15000.0, N); lam = 0.6328; k = 2 * np.pi / lam; dx = x[1] - x[0]
us = [torch.randn(N, N, dtype=torch.complex64, device=dev) for _ in range(3)]
es = [torch.randn(2, N, N, dtype=torch.complex64, device=dev) for _ in range(3)]
ct3 = torch.randn(3, N, N, dtype=torch.complex64, device=dev)
z = torch.tensor([50000.0], dtype=torch.float64, device=dev)
xo, yo = xb.space(10.0, 400); x2, y2 = xb.space(2500.0, N)

def rsgrad(i):
    uu = us[i % 3].detach().requires_grad_(True); zz = (z + 0.1 * i).requires_grad_(True)
    o = ops.rs_propagation(uu, zz, dx, dx, k); o.backward(us[(i + 1) % 3])
def vrsgrad(i):
    e = es[i % 3].detach().requires_grad_(True); zz = (z + 0.1 * i).requires_grad_(True)
    o = ops.vrs_propagation(e, None, zz, x[0], y[0], dx, dx, k); o.backward(ct3)
def cztgrad(i):
    uu = us[i % 3].detach().requires_grad_(True)
    o = ops.czt(uu, 5000.0 + 0.01 * i, lam, x, y, x, y); o.backward(us[(i + 1) % 3])
def vcztgrad(i):
    e = es[i % 3].detach().requires_grad_(True)
    o = ops.vczt(e, None, 5000.0 + 0.01 * i, lam, x, y, x, y); o.backward(ct3)
def highna(i):
    e = es[i % 3].detach().requires_grad_(True)
    o = ops.highna_focus(e, None, 1800.0, 2000.0, 0.635, x2, y2, xo, yo); o.backward(o.detach())

buf = ctypes.create_string_buffer(1 << 16)
for name, fn in (("RS fwd+grad", rsgrad), ("VRS fwd+grad", vrsgrad), ("CZT fwd+grad", cztgrad), ("VCZT fwd+grad", vcztgrad), ("highNA fwd+grad", highna)):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / iters * 1e3
    L.xl_prof_enable(1)
    for i in range(iters): fn(i)
    torch.cuda.synchronize()
    L.xl_prof_report(buf, len(buf)); L.xl_prof_enable(0)
    print(f"--- {name}: {total:9.1f} us per call (unprofiled)")
    ksum = 0.0
    for ln in buf.value.decode().strip().splitlines():
        nm, cnt, tot = ln.split()
        per_call = float(tot) * 1e3 / iters
        ksum += per_call
        print(f"    {nm:14s} {int(cnt) // iters:3d} launches  {per_call:9.1f} us/call  {float(tot) * 1e3 / int(cnt):8.1f} us/launch")
    print(f"    kernels sum {ksum:9.1f} us")
