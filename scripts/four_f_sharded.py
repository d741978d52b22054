#!/usr/bin/env python
"""
cfg 4 of BASELINE.json: the 4f-system phase-mask optimizer with a batch of candidate input masks sharded over the GPUs of
one box (reference: experiments/four_f_optical_table.py:36-141, experiments/four_f_optimizer.py:35-112,
experiments/generate_synthetic_data.py:38-60).

Per sample:  mask -> RS(z0) -> SLM(phase1) -> RS(z1) -> SLM(phase2) -> RS(z2) -> |.|^2,  loss = mean MSE-intensity against
the 2x-magnified mask; parameters (3 distances, 2 phase masks) are shared by the whole batch, so every propagation uses ONE
transfer function for all the samples of a rank (a batch of fields per library call), and the only collective is one
flattened all-reduce of the parameter gradients per step (xlumina_b200.sharding.allreduce_grads).

    python scripts/four_f_sharded.py [--batch 64] [--n 1024] [--steps 10] [--warmup 3]
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 scripts/four_f_sharded.py ...
Prints one JSON line (optimizer steps/s; "strong" scaling: the global batch is fixed).
"""
import argparse
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

import numpy as np
import torch
import torch.distributed as dist

import xlumina_b200 as xb
from xlumina_b200 import four_f
from xlumina_b200.sharding import allreduce_flat, flat_grad_buffers, shard_range


def synthetic_circles(n_samples, x, rng):
    """Binary elliptical masks and their 2x-magnified target intensities |0.5 * mask(2r)|^2 (generate_synthetic_data.py:38-60)."""
    X, Y = np.meshgrid(x, x)
    masks, targets = [], []
    for _ in range(n_samples):
        r1, r2 = rng.uniform(100, 1000, size=2)
        masks.append(((X / r1) ** 2 + (Y / r2) ** 2 < 1).astype(np.float32))
        targets.append((0.25 * ((X / (2 * r1)) ** 2 + (Y / (2 * r2)) ** 2 < 1)).astype(np.float32))
    return np.stack(masks), np.stack(targets)


class _Beam:
    """The minimal light-source view four_f.vector_dualSLM_4f_system needs (axis, wavenumber, field)."""

    def __init__(self, x, k, field):
        self.x, self.k, self.field = x, k, field


def forward_loss(params, masks, targets, beam, dx, k, fused=True):
    """loss_dualSLM of the reference (four_f_optical_table.py:103-141) on this rank's slice; returns the SUM of per-sample
    MSEs (the caller divides by the GLOBAL batch).  The table itself lives in xlumina_b200/four_f.py; `fused` folds the
    beam, both SLMs and the intensity-MSE detector into the propagation kernels (four_f.loss_dualSLM_fused)."""
    x = dx * (np.arange(beam.shape[-1]) - (beam.shape[-1] - 1) / 2)
    p = [params[0], params[1], params[2], params[3], params[4]]
    if fused:
        return four_f.loss_dualSLM_fused(p, masks, targets, _Beam(x, k, beam)) * masks.shape[0]
    inten, _, _ = four_f.vector_dualSLM_4f_system(masks, _Beam(x, k, beam), p)
    return four_f.MSE_Intensity(inten, targets).sum()


def setup(batch, n, dev, rank, world, fused=True, graph=False, fused_adam=True):
    """Build the sharded optimizer problem on `dev`; returns (step, params, samples_on_this_rank).  step() runs one optimizer
    step (forward, backward, one flattened gradient all-reduce when world > 1, AdamW) and returns this rank's loss share."""
    N, lam = n, 0.6328
    x, _ = xb.space(1500.0, N)
    dx, k = float(x[1] - x[0]), 2 * math.pi / lam
    rng = np.random.default_rng(0)                                    # same data on every rank; each takes its slice
    masks_all, targets_all = synthetic_circles(batch, x, rng)
    a, b = shard_range(batch, rank, world)
    masks = torch.as_tensor(masks_all[a:b], device=dev).to(torch.complex64)
    targets = torch.as_tensor(targets_all[a:b], device=dev)
    X, Y = np.meshgrid(x, x)
    beam = torch.as_tensor(np.exp(-(X ** 2 + Y ** 2) / 1200.0 ** 2).astype(np.complex64), device=dev)   # gaussian_beam, z_w0 = 0
    prng = np.random.default_rng(1)                                   # four_f_optimizer.py:104-109
    params = [torch.tensor([prng.uniform(0.027, 1)], dtype=torch.float64, device=dev, requires_grad=True) for _ in range(3)]
    params += [torch.tensor(prng.uniform(0, 1, (N, N)).astype(np.float32), device=dev, requires_grad=True) for _ in range(2)]
    # optax.adamw(0.01, weight_decay=1e-4), four_f_optimizer.py:97-98,112; fused_adam: torch's single-kernel multi-tensor AdamW
    opt = torch.optim.AdamW(params, lr=0.01, weight_decay=1e-4, capturable=True, fused=True if fused_adam else None)
    flats = flat_grad_buffers(params)       # .grad of every parameter = a view of one flat buffer per dtype

    def compute():
        opt.zero_grad(set_to_none=False)
        loss = forward_loss(params, masks, targets, beam, dx, k, fused) / batch    # mean over the GLOBAL batch
        loss.backward()
        return loss.detach()

    def step():
        loss = compute()
        allreduce_flat(flats)
        opt.step()
        return loss

    if graph:
        # Forward + backward of this rank's slice as ONE CUDA graph and AdamW as a second one (the library neither allocates nor
        # synchronises in steady state, so it captures as is); the gradient all-reduce stays an eager NCCL call between them.
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                compute()
                opt.step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if graph == "nccl":
            # the whole step INCLUDING the NCCL all-reduce as one graph (thread-local capture mode: the NCCL watchdog thread
            # may query events while this thread captures)
            g_all = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_all, capture_error_mode="thread_local"):
                static_loss = compute()
                allreduce_flat(flats)
                opt.step()

            def step():   # noqa: F811
                g_all.replay()
                return static_loss
        else:
            g_compute, g_opt = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_compute):
                static_loss = compute()
            with torch.cuda.graph(g_opt):
                opt.step()

            def step():   # noqa: F811
                g_compute.replay()
                allreduce_flat(flats)
                g_opt.replay()
                return static_loss

    return step, params, b - a


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--unfused", action="store_true", help="pointwise elements and the loss as separate torch operations")
    ap.add_argument("--graph", action="store_true", help="replay forward+backward and AdamW as CUDA graphs (the all-reduce stays eager)")
    ap.add_argument("--unfused-adam", action="store_true", help="torch.optim.AdamW without fused=True (the multi-tensor foreach kernels; round 2: +0.12 ms per step)")
    ap.add_argument("--graph-nccl", action="store_true", help="one CUDA graph for the whole step, the NCCL all-reduce captured too")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N = args.n
    step, params, mine = setup(args.batch, N, dev, rank, world, fused=not args.unfused, graph=("nccl" if args.graph_nccl else args.graph), fused_adam=not args.unfused_adam)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    lsum = loss.double().clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lsum, op=dist.ReduceOp.SUM)
    if rank == 0:
        ms = float(t[0]) / args.steps
        print(json.dumps({"metric": "4f optimizer steps/s (batch %d, %d^2, 3 RS fwd+grad per sample, shared parameters)" % (args.batch, N),
                          "value": 1e3 / ms, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "samples_per_rank": mine,
                          "propagations_per_s": 3 * args.batch * 1e3 / ms, "loss": float(lsum), "fused_elements": not args.unfused, "cuda_graphs": "with nccl" if args.graph_nccl else bool(args.graph),
                          "collective": "one all-reduce of 2*N^2 fp32 + 3 fp64 gradients per step"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
