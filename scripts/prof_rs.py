"""Profile target: a few calls of one operator at N (default 2048) -- run under ncu.
modes: fwd (RS forward), grad (RS forward + backward with d/dz), czt, cztgrad, vrsgrad, vczt, vcztgrad, highna"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import xlumina_b200 as xb
from xlumina_b200 import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
mode = sys.argv[2] if len(sys.argv) > 2 else "fwd"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
x, y = xb.space(15000.0, N); lam = 0.6328; k = 2*np.pi/lam; dx = x[1]-x[0]
u = torch.randn(N, N, dtype=torch.complex64, device=dev)
exy = torch.randn(2, N, N, dtype=torch.complex64, device=dev)
z = torch.tensor([50000.0], dtype=torch.float64, device=dev)
for it in range(iters):
    if mode == "fwd":
        ops.rs_propagation(u, z, dx, dx, k)
    elif mode == "grad":
        uu = u.detach().requires_grad_(True); zz = z.detach().requires_grad_(True)
        o = ops.rs_propagation(uu, zz, dx, dx, k); o.backward(o)
    elif mode == "vrsgrad":
        e = exy.detach().requires_grad_(True); zz = z.detach().requires_grad_(True)
        o = ops.vrs_propagation(e, None, zz, x[0], y[0], dx, dx, k); o.backward(o)
    elif mode == "czt":
        ops.czt(u, 5000.0, lam, x, y, x, y)
    elif mode == "cztgrad":
        uu = u.detach().requires_grad_(True)
        o = ops.czt(uu, 5000.0, lam, x, y, x, y); o.backward(o)
    elif mode == "vczt":
        ops.vczt(exy, None, 5000.0, lam, x, y, x, y)
    elif mode == "vcztgrad":
        e = exy.detach().requires_grad_(True)
        o = ops.vczt(e, None, 5000.0, lam, x, y, x, y); o.backward(o)
    elif mode == "highna":
        xo, yo = xb.space(10.0, 400); x2, y2 = xb.space(2500.0, N)
        e = exy.detach().requires_grad_(True)
        o = ops.highna_focus(e, None, 1800.0, 2000.0, 0.635, x2, y2, xo, yo); o.backward(o)
torch.cuda.synchronize()
