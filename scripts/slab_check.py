#!/usr/bin/env python
"""Multi-GPU check + timing of the slab-decomposed RS path: every rank owns N/G rows; result compared with the single-GPU
library on the same field.   python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 scripts/slab_check.py [N]"""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"
import numpy as np, torch, torch.distributed as dist
import xlumina_b200 as xb
from xlumina_b200 import ops, slab

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
x, _ = xb.space(15000.0, N); lam = 0.6328; k = 2 * math.pi / lam; dx = float(x[1] - x[0]); z = 50000.0
g = torch.Generator(device="cpu").manual_seed(3)
rows = N // world
if N <= 2048:      # reference: the fused single-GPU path on the same random field
    full = torch.view_as_complex(torch.randn(N, N, 2, generator=g)).to(dev)
    ref = ops.rs_propagation(full, z, dx, dx, k)[rank * rows:(rank + 1) * rows]
else:              # beyond the fused path: a point source must reproduce the sampled impulse response (wave_optics.py:291-297)
    i0, j0 = N // 3, (2 * N) // 5
    full = torch.zeros(N, N, dtype=torch.complex64, device=dev)
    full[i0, j0] = 1.0
    q = (torch.arange(N, device=dev, dtype=torch.float64) - j0) * dx
    p = (torch.arange(rank * rows, (rank + 1) * rows, device=dev, dtype=torch.float64) - i0) * dx
    r = torch.sqrt(p[:, None] ** 2 + q[None, :] ** 2 + z * z)
    ref = ((1 / (2 * math.pi)) * z / r ** 2 * (1 / r - 1j * k) * torch.exp(1j * k * r) * dx * dx).to(torch.complex64)
mine = full[rank * rows:(rank + 1) * rows].contiguous()
out, H = slab.rs_propagation_slab(mine, z, dx, dx, k, return_transfer=True)
err = float(torch.linalg.norm(out - ref) / torch.linalg.norm(ref))
def timed(fn, it=10):
    for _ in range(3): fn()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / it], device=dev, dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]) * 1e3
t_fresh = timed(lambda: slab.rs_propagation_slab(mine, z, dx, dx, k))
t_reuse = timed(lambda: slab.rs_propagation_slab(mine, z, dx, dx, k, transfer=H))
t_single = timed(lambda: ops.rs_propagation(full, z, dx, dx, k)) if N <= 2048 else None
e = torch.tensor([err], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(e, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"slab_rs": {"N": N, "n_gpus": world, "max_rel_l2_vs_reference": float(e[0]), "reference": "fused single-GPU path" if N <= 2048 else "analytic impulse response (point source)",
                                  "us_fresh_z": t_fresh, "us_transfer_reused": t_reuse, "us_single_gpu_fresh_z": t_single,
                                  "exchanged_bytes_per_rank_per_all_to_all": (N * 2 * N * 8 // world) * (world - 1) // world}}), flush=True)
if world > 1: dist.destroy_process_group()
