#!/usr/bin/env python
"""Kernel-level breakdown of one cfg-4 optimizer step on one GPU (torch.profiler; development aid).
   python scripts/four_f_profile.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import four_f_sharded as ff
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
step, params, mine = ff.setup(batch, 1024, dev, 0, 1, fused=True, graph=False)
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
print(f"batch {batch}: CUDA kernels per step (5 steps averaged), total {sum(e.device_time_total for e in evs) / 5:.1f} us")
for e in evs[:45]:
    print(f"  {e.device_time_total / 5:9.1f} us  x{e.count / 5:5.1f}  {e.key[:120]}")
