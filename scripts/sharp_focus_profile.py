#!/usr/bin/env python
"""Kernel-level breakdown of one cfg-3 loss value+gradient on one GPU (torch.profiler; development aid): library kernels
against the torch pointwise kernels of the elements and losses around them.   python scripts/sharp_focus_profile.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import sharp_focus_table as sf
from xlumina_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
ls, params, fixed = sf.build_problem(n, 400, dev)
ops.set_transfer_cache(8)
def step():
    for p in params:
        p.grad = None
    loss = sf.loss_hybrid_sharp_focus(ls, params, fixed)
    loss.backward()
for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
evs.sort(key=lambda e: -e.device_time_total)
lib = sum(e.device_time_total for e in evs if "xl_kernel" in e.key) / 5
tot = sum(e.device_time_total for e in evs) / 5
print(f"n {n}: CUDA time per loss+grad {tot:.1f} us, library kernels {lib:.1f} us ({100 * lib / tot:.1f} %), torch kernels {tot - lib:.1f} us")
for e in evs[:40]:
    print(f"  {e.device_time_total / 5:9.1f} us  x{e.count / 5:5.1f}  {e.key[:110]}")
