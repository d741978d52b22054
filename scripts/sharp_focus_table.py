#!/usr/bin/env python
"""
cfg 3 of BASELINE.json: one value+gradient of the hybrid sharp-focus loss (reference:
experiments/hybrid_sharp_optical_table.py:26-56, optical_elements.hybrid_setup_sharp_focus :1503-1649) -- 16 vectorial RS
propagations and 6 high-NA focusings between beam splitters, sSLMs and wave plates, 29 parameters (6 phase masks N x N,
6 wave-plate angles, 8 distances, 9 splitter ratios), loss = softmin over the six detectors of small_area_hybrid(|Ez|^2).

    python scripts/sharp_focus_table.py [--n 1024] [--m 400] [--steps 10] [--warmup 3] [--cache 8]
Prints one JSON line (loss evaluations with gradient per second on ONE GPU; candidates of a discovery run are independent,
so more GPUs are replicas -- see bench.py / xlumina_b200.sharding for the batch-sharded pattern).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import xlumina_b200 as xb
from xlumina_b200 import loss_functions, ops
from xlumina_b200.toolbox import softmin


def build_problem(n, m, device, seed=0):
    """The reference's set-up: 635 nm, +-2500 um window, Gaussian beam w0 = 1200 um polarised (1, 1), objective NA 0.9 with
    radius 1.8 mm, detector window +-10 um sampled m x m; parameters ~ U(0, 1) (hybrid_sharp_optical_table.py:26-46)."""
    wavelength = 635 * xb.nm
    x, y = xb.space(2500 * xb.um, n)
    ls = xb.PolarizedLightSource(x, y, wavelength, device=device)
    ls.gaussian_beam(w0=(1200 * xb.um, 1200 * xb.um), jones_vector=(1, 1))
    x_out, y_out = xb.space(10 * xb.um, m)
    radius = 3.6 * xb.mm / 2
    fixed = [radius, radius / 0.9, x_out, y_out]
    rng = np.random.default_rng(seed)
    masks = (0, 1, 6, 7, 12, 13)
    params = [torch.tensor(rng.uniform(0, 1, (n, n)).astype(np.float32) if i in masks else rng.uniform(0, 1, (1,)),
                           device=device, requires_grad=True) for i in range(29)]
    return ls, params, fixed


def loss_hybrid_sharp_focus(ls, params, fixed):
    """hybrid_sharp_optical_table.py:49-56."""
    intensities, _ = xb.hybrid_setup_sharp_focus(ls, ls, ls, ls, ls, ls, params, fixed)
    return softmin(loss_functions.vectorized_loss_hybrid(intensities))


def make_step(ls, params, fixed, graph=False):
    """step() -> loss: one value + gradient of the table (parameter gradients land in p.grad).  graph=True replays the whole
    evaluation -- 22 propagations, every element, the loss and the backward pass -- as ONE CUDA graph: the library neither
    allocates nor synchronises in steady state, and the table's ~1500 small launches are otherwise bound by the host."""
    def compute():
        for p in params:
            p.grad = None
        loss = loss_hybrid_sharp_focus(ls, params, fixed)
        loss.backward()
        return loss.detach()

    if not graph:
        return compute
    dev = params[0].device
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            compute()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    ops.set_transfer_cache(ops._transfer_cache_size)   # same size; the entries of the warm-up are dropped below
    ops._transfer_cache.clear()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        static_loss = compute()
    ops._transfer_cache.clear()                        # entries made during capture live in the graph's memory pool

    def step():
        g.replay()
        return static_loss
    return step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--m", type=int, default=400)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cache", type=int, default=8, help="transfer functions kept for distances repeated inside the table (0 = off)")
    ap.add_argument("--graph", action="store_true", help="replay the whole value+gradient evaluation as one CUDA graph")
    args = ap.parse_args()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    ls, params, fixed = build_problem(args.n, args.m, dev)
    ops.set_transfer_cache(args.cache)

    step = make_step(ls, params, fixed, graph=args.graph)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches0 = xb.ops._lib.lib().xl_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (xb.ops._lib.lib().xl_launch_count() - launches0) // args.steps
    finite = all(p.grad is None or bool(torch.isfinite(p.grad).all()) for p in params)
    print(json.dumps({"metric": "sharp-focus table value+grad per second (%d^2 -> %d^2, 16 VRS + 6 high-NA focus, 29 parameters)" % (args.n, args.m),
                      "value": 1e3 / ms, "unit": "loss-grads/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": ms, "higher_is_better": True, "loss": float(loss), "gradients_finite": finite,
                      "library_launches_per_step": int(launches), "transfer_cache": args.cache, "cuda_graph": bool(args.graph),
                      "propagations_per_s": 22 * 1e3 / ms}), flush=True)


if __name__ == "__main__":
    main()
