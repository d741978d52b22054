"""Quick GPU probe: smoke parity + CUDA-event timings of every operator (not the bench; development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import xlumina_b200 as xb
from xlumina_b200 import ops

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us

def main():
    import __graft_entry__ as g
    if "--nosmoke" not in sys.argv:
        g.smoke()
    dev = torch.device("cuda:0")
    sizes = (2048,) if "--only2048" in sys.argv else (1024, 2048)
    for N in sizes:
        x, y = xb.space(15000.0, N); lam = 0.6328; k = 2*np.pi/lam
        dx = x[1]-x[0]
        u = torch.randn(N, N, dtype=torch.complex64, device=dev)
        exy = torch.randn(2, N, N, dtype=torch.complex64, device=dev)
        z = torch.tensor([50000.0], dtype=torch.float64, device=dev)
        print(f"--- N={N}")
        t = timeit(lambda: ops.rs_propagation(u, z, dx, dx, k)); print(f"RS fwd           {t:9.1f} us")
        def fb():
            uu = u.detach().requires_grad_(True); zz = z.detach().requires_grad_(True)
            o = ops.rs_propagation(uu, zz, dx, dx, k); o.backward(o)
        t = timeit(fb); print(f"RS fwd+grad      {t:9.1f} us")
        def fb2():
            uu = u.detach().requires_grad_(True)
            o = ops.rs_propagation(uu, z, dx, dx, k); o.backward(o)
        t = timeit(fb2); print(f"RS fwd+grad(noz) {t:9.1f} us")
        t = timeit(lambda: ops.vrs_propagation(exy[0], exy[1], z, x[0], y[0], dx, dx, k)); print(f"VRS fwd          {t:9.1f} us")
        def fbv():
            e = exy.detach().requires_grad_(True); zz = z.detach().requires_grad_(True)
            o = ops.vrs_propagation(e[0], e[1], zz, x[0], y[0], dx, dx, k); o.backward(o)
        t = timeit(fbv); print(f"VRS fwd+grad     {t:9.1f} us")
        zc = 5000.0
        t = timeit(lambda: ops.czt(u, zc, lam, x, y, x, y)); print(f"CZT fwd          {t:9.1f} us")
        def fbc():
            uu = u.detach().requires_grad_(True)
            o = ops.czt(uu, zc, lam, x, y, x, y); o.backward(o)
        t = timeit(fbc); print(f"CZT fwd+grad     {t:9.1f} us")
        t = timeit(lambda: ops.vczt(exy[0], exy[1], zc, lam, x, y, x, y)); print(f"VCZT fwd         {t:9.1f} us")
        xo, yo = xb.space(10.0, 400)
        x2, y2 = xb.space(2500.0, N)
        t = timeit(lambda: ops.highna_focus(exy[0], exy[1], 1800.0, 2000.0, 0.635, x2, y2, xo, yo)); print(f"highNA fwd ->400 {t:9.1f} us")
        def fbh():
            e = exy.detach().requires_grad_(True)
            o = ops.highna_focus(e[0], e[1], 1800.0, 2000.0, 0.635, x2, y2, xo, yo); o.backward(o)
        t = timeit(fbh); print(f"highNA fwd+grad  {t:9.1f} us")
        # raw ABI pieces
        L = xb._lib.lib()
        H = torch.empty(L.xl_rs_transfer_bytes(N), dtype=torch.uint8, device=dev)
        st = ops._stream(H)
        t = timeit(lambda: L.xl_rs_transfer(ops._ptr(H), ops._ptr(z), N, dx, dx, k, 0, st)); print(f"  transfer H     {t:9.1f} us")
        ws = ops._workspace(u, L.xl_rs_workspace_bytes(N, 1, 1)); out = torch.empty_like(u)
        t = timeit(lambda: L.xl_rs_fwd(ops._ptr(u), ops._ptr(out), ops._ptr(H), ops._ptr(z), N, 1, dx, dx, k, 16, ops._ptr(ws), ws.numel(), st)); print(f"  apply (H reuse){t:9.1f} us")

if __name__ == "__main__":
    main()
