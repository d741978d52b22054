#!/usr/bin/env python
"""Large-grid scalar RS through the slab chain on ONE GPU (split-line kernels, csrc/xl_long.cuh): a point source must
reproduce the sampled impulse response, out[p,q] = dx dy h((q-j0) dx, (p-i0) dy; z) (wave_optics.py:291-297), which pins every
stage at sizes no CPU oracle run reaches; then a random field is timed.   python scripts/long_check.py [N ...]"""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xlumina_b200 as xb
from xlumina_b200 import slab

dev = torch.device("cuda:0")
lam, z = 0.6328, 5.0e4
k = 2 * math.pi / lam
for N in [int(a) for a in sys.argv[1:]] or [4096]:
    x, _ = xb.space(15000.0, N)
    dx = float(x[1] - x[0])
    i0, j0 = N // 3, (2 * N) // 5
    f = torch.zeros(N, N, dtype=torch.complex64, device=dev)
    f[i0, j0] = 1.0
    out, H = slab.rs_propagation_slab(f, z, dx, dx, k, return_transfer=True)
    q = (torch.arange(N, device=dev, dtype=torch.float64) - j0) * dx
    p = (torch.arange(N, device=dev, dtype=torch.float64) - i0) * dx
    r = torch.sqrt(p[:, None] ** 2 + q[None, :] ** 2 + z * z)
    ref = (1 / (2 * math.pi)) * z / r ** 2 * (1 / r - 1j * k) * torch.exp(1j * k * r) * dx * dx
    err = float(torch.linalg.norm(out.to(torch.complex128) - ref) / torch.linalg.norm(ref))
    del ref, r
    g = torch.Generator(device="cpu").manual_seed(1)
    u = torch.view_as_complex(torch.randn(N, N, 2, generator=g)).to(dev)
    for _ in range(2):
        slab.rs_propagation_slab(u, z, dx, dx, k, transfer=H)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for _ in range(3):
        slab.rs_propagation_slab(u, z, dx, dx, k, transfer=H)
    e1.record()
    for _ in range(3):
        slab.rs_propagation_slab(u, z, dx, dx, k)
    e2.record()
    torch.cuda.synchronize()
    ub = 8.0 * N * N
    t_reuse, t_fresh = e0.elapsed_time(e1) / 3, e1.elapsed_time(e2) / 3
    print(json.dumps({"long_rs": {"N": N, "padded": 2 * N, "n_gpus": 1, "point_source_rel_l2_vs_analytic_h": err,
                                  "ms_transfer_reused": t_reuse, "ms_fresh_z": t_fresh,
                                  "alg_GBps_fresh_z(14u)": 14 * ub / t_fresh / 1e6,
                                  "max_mem_GB": torch.cuda.max_memory_allocated() / 1e9}}), flush=True)
    del out, H, u, f
    torch.cuda.empty_cache()
