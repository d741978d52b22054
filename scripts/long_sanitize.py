"""Small run of the split-line chain for compute-sanitizer (racecheck / memcheck): sub-line length forced to 32 so that 2, 4 and
8 sub-lines -- cluster kernels (R <= 4) and the two-launch form (R = 8) -- all execute, forward and d/dz.
    compute-sanitizer --tool racecheck python scripts/long_sanitize.py"""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xlumina_b200 as xb
from xlumina_b200 import slab, _lib

L = _lib.lib()
L.xl_debug_set_max_line(32)
dev = torch.device("cuda:0")
k = 2 * math.pi / 0.6328
for N in (24, 48, 96):                                   # padded 64 / 128 / 256 = 2 / 4 / 8 sub-lines of 32
    x, _ = xb.space(600.0, N)
    dx = float(x[1] - x[0])
    g = torch.Generator(device="cpu").manual_seed(N)
    u = torch.view_as_complex(torch.randn(N, N, 2, generator=g)).to(dev)
    ct = torch.view_as_complex(torch.randn(N, N, 2, generator=g)).to(dev)
    out, H = slab.rs_propagation_slab(u, 7000.0, dx, dx, k, return_transfer=True, group=slab._LOCAL)
    v = slab.rs_slab_vjp(ct, H, group=slab._LOCAL)
    gz = slab.rs_slab_grad_z(u, ct, out, 7000.0, dx, dx, k, group=slab._LOCAL)
    L.xl_debug_set_long_cluster(1 if N == 96 else 0)     # the other form of the inverse step
    out2 = slab.rs_propagation_slab(u, 7000.0, dx, dx, k, group=slab._LOCAL)
    L.xl_debug_set_long_cluster(-1)
    torch.cuda.synchronize()
    print(N, float(torch.linalg.norm(out)), float(torch.linalg.norm(out - out2) / torch.linalg.norm(out)), float(gz))
L.xl_debug_set_max_line(4096)
