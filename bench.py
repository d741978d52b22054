#!/usr/bin/env python
"""
bench.py -- propagations/s of XLuminA's propagation hot path (forward + gradient) at 2048^2 on B200.

One "step" = one pass of the hot path over one batch of synthetic fields (BASELINE.json metric / configs[0..2]):
    scalar RS fwd+grad (d/dfield, d/dz)  +  VRS fwd+grad (d/dEx, d/dEy, d/dz)  +  CZT fwd+grad  +  VCZT fwd+grad
= 4 propagations at 2048 x 2048, lambda = 632.8 nm, window +-15 mm (examples/scalar_xlumina.py:22-33,
examples/vectorial_xlumina.py:21-32 of the reference), RS/VRS at z = 5 cm (BASELINE.json) and CZT/VCZT at z = 5 mm (the
files' value), a FRESH z every step (so the transfer function is regenerated, as in an optimizer step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

N > 1: the path shards by independent candidate set-ups: every rank runs the same step on its own batch,
no data-path collective, value = all ranks' propagations / max-over-ranks time ("scaling": "weak").

--impl reference: the CPU restatement of the reference (oracle/, torch-CPU complex128 with autograd, all host threads) on
a bounded sample of the same workload; real JAX is not installable in this image.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION on the GPU boxes) off it
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

N_GRID = 2048
LAMBDA = 0.6328
HALF_WINDOW = 15000.0
Z_RS = 50000.0      # 5 cm  (BASELINE.json configs[0])
Z_CZT = 5000.0      # 5 mm  (examples/*_xlumina.py)
PROPS_PER_STEP = 4
METRIC = "RS/VRS/CZT propagations/s at 2048^2 (fwd+grad)"
UNIT = "propagations/s"
WORKLOAD = ("scalar RS + VRS (z=5cm) + CZT + VCZT (z=5mm, 2048->2048), each forward+gradient, 2048x2048 complex64, "
            "lambda=632.8nm, window +-15mm, fresh z per step; working set per step (~2.6 GB of fields, spectra and transfer "
            "functions) exceeds the 126 MB L2, and 4 input sets rotate between steps")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region (pynvml; nvidia-smi fields of the profiling recipe)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_step_times(threads, reps, mix):
    """Forward+gradient at 2048^2 with the torch-CPU complex128 restatement of the reference (oracle/oracle_torch.py).
    mix=True : step i runs propagation i % 4 of the bench's own mix (scalar RS, VRS, CZT, VCZT; same shapes and z as the GPU
               arm), i.e. every step is ONE propagation and 4 consecutive steps are one GPU-arm step.
    mix=False: every step = scalar RS (the bounded sample used for the in-line cpu_baseline).  Returns seconds per step."""
    import numpy as np
    import torch
    from oracle import oracle_torch as ot
    torch.set_num_threads(threads)
    rng = np.random.default_rng(0)
    x = np.linspace(-HALF_WINDOW, HALF_WINDOW, N_GRID)

    def crand(*shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    u0, e0 = crand(N_GRID, N_GRID), crand(2, N_GRID, N_GRID)
    ct1, ct3 = torch.tensor(crand(N_GRID, N_GRID)), torch.tensor(crand(3, N_GRID, N_GRID))
    times = []
    for r in range(reps):
        kind = r % 4 if mix else 0
        t0 = time.perf_counter()
        if kind == 0:
            u = torch.tensor(u0, requires_grad=True)
            z = torch.tensor(Z_RS + 0.37 * r, dtype=torch.float64, requires_grad=True)
            torch.real(torch.sum(ct1 * ot.RS_propagation(u, x, x, LAMBDA, z))).backward()
        elif kind == 1:
            e = torch.tensor(e0, requires_grad=True)
            zv = torch.tensor(Z_RS + 0.53 * r, dtype=torch.float64, requires_grad=True)
            torch.real(torch.sum(ct3 * ot.VRS_propagation(e[0], e[1], x, x, LAMBDA, zv))).backward()
        elif kind == 2:
            c = torch.tensor(u0, requires_grad=True)
            torch.real(torch.sum(ct1 * ot.CZT(c, x, x, LAMBDA, Z_CZT + 0.01 * r, x, x))).backward()
        else:
            v = torch.tensor(e0, requires_grad=True)
            torch.real(torch.sum(ct3 * ot.VCZT(v[0], v[1], x, x, LAMBDA, Z_CZT + 0.02 * r, x, x))).backward()
        times.append(time.perf_counter() - t0)
    return times


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path on this box's host cores.  JAX is not installable in this image
    (DESIGN.md), so this is the line-by-line torch-CPU complex128 restatement, with autograd, on all host threads.  Each
    step is a bounded sample of the GPU arm's step: ONE of its four propagations, cycling RS, VRS, CZT, VCZT."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    w4 = 4 * ((args.warmup + 3) // 4)          # keep the cycle aligned: timed step 0 is scalar RS
    times = cpu_step_times(threads, w4 + args.steps, mix=True)[w4:]
    total = sum(times)
    value = len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": {"workload": WORKLOAD, "l2": "n/a (CPU)", "sharding": "rank 0 only"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "each step = ONE forward+gradient propagation at 2048^2, cycling scalar RS, VRS, CZT, VCZT "
                                   "(4 steps = 1 step of the GPU arm); torch-CPU complex128 restatement of the reference "
                                   "(oracle/oracle_torch.py) on all host threads; the reference itself needs JAX, not installable here"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    import xlumina_b200 as xb
    from xlumina_b200 import ops, _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    x, y = xb.space(HALF_WINDOW, N_GRID)
    dx = float(x[1] - x[0])
    k = 2 * np.pi / LAMBDA
    NSETS = 4
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)

    def host_c64(*shape):
        t = torch.randn(*shape, 2, generator=g, dtype=torch.float32)
        return torch.view_as_complex(t).contiguous().pin_memory()

    host_sets = [{"u": host_c64(N_GRID, N_GRID), "exy": host_c64(2, N_GRID, N_GRID),
                  "c": host_c64(N_GRID, N_GRID), "vxy": host_c64(2, N_GRID, N_GRID)} for _ in range(NSETS)]
    dev_sets = [{kk: v.to(dev) for kk, v in s.items()} for s in host_sets]
    cts = {"u": torch.randn(N_GRID, N_GRID, dtype=torch.complex64, device=dev),
           "v3": torch.randn(3, N_GRID, N_GRID, dtype=torch.complex64, device=dev)}
    h2d_bytes = sum(v.numel() * 8 for v in host_sets[0].values())
    z_base = torch.full((1,), Z_RS, dtype=torch.float64, device=dev)

    def step(i, s):
        """Public-API forward + gradient of the four propagators on input set `s`; returns device scalars."""
        # fresh z every step, derived on the device (no pageable host-to-device copy that would make the host wait)
        zr = (z_base + 0.37 * i).requires_grad_(True)
        zv = (z_base + 0.53 * i).requires_grad_(True)
        u = s["u"].detach().requires_grad_(True)
        o1 = ops.rs_propagation(u, zr, dx, dx, k)
        o1.backward(cts["u"])
        exy = s["exy"].detach().requires_grad_(True)
        o2 = ops.vrs_propagation(exy, None, zv, float(x[0]), float(y[0]), dx, dx, k)
        o2.backward(cts["v3"])
        c = s["c"].detach().requires_grad_(True)
        o3 = ops.czt(c, Z_CZT + 0.01 * i, LAMBDA, x, y, x, y)
        o3.backward(cts["u"])
        vxy = s["vxy"].detach().requires_grad_(True)
        o4 = ops.vczt(vxy, None, Z_CZT + 0.02 * i, LAMBDA, x, y, x, y)
        o4.backward(cts["v3"])
        return zr.grad, zv.grad, u.grad, exy.grad, c.grad, vxy.grad

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) -----------------------------------------------------------------------
    for i in range(args.warmup):
        step(i, dev_sets[i % NSETS])
    barrier()
    launches0 = L.xl_launch_count()
    L.xl_prof_enable(1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i, dev_sets[i % NSETS])
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = L.xl_launch_count() - launches0
    buf = __import__("ctypes").create_string_buffer(1 << 16)
    L.xl_prof_report(buf, len(buf))
    L.xl_prof_enable(0)
    kern = {}
    for ln in buf.value.decode().strip().splitlines():
        nm, cnt, tot = ln.split()
        kern[nm] = (int(cnt), float(tot))

    # ---- roofline of the dominant kernel: CUDA events around each launch of the scalar-RS column kernels (1 field per
    # launch, the shape the committed ncu capture has), inputs rotating through sets larger than L2 -------------------
    L.xl_prof_enable(1)
    for i in range(max(args.steps, 5)):
        sset = dev_sets[i % NSETS]
        zr = (z_base + 0.11 * i).requires_grad_(True)
        u = sset["u"].detach().requires_grad_(True)
        ops.rs_propagation(u, zr, dx, dx, k).backward(cts["u"])
    torch.cuda.synchronize()
    L.xl_prof_report(buf, len(buf))
    L.xl_prof_enable(0)
    kern1 = {}
    for ln in buf.value.decode().strip().splitlines():
        nm, cnt, tot = ln.split()
        kern1[nm] = (int(cnt), float(tot))

    # ---- end-to-end timing through the public API with HOST buffers (e2e) ------------------------------------------
    copy_stream = torch.cuda.Stream(device=dev)
    staged = [dict((kk, torch.empty_like(v, device=dev)) for kk, v in host_sets[0].items()) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def stage(i):
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])
            for kk, v in host_sets[i % NSETS].items():
                staged[b][kk].copy_(v, non_blocking=True)
            ready[b].record(copy_stream)

    def e2e_loop(n, base):
        res = torch.empty(2, dtype=torch.float64).pin_memory()
        for b in range(2):
            consumed[b].record(torch.cuda.current_stream(dev))
        stage(0)
        out = None
        for i in range(n):
            b = i % 2
            if i + 1 < n:
                stage(i + 1)
            torch.cuda.current_stream(dev).wait_event(ready[b])
            gz1, gz2, *_ = step(base + i, staged[b])
            consumed[b].record(torch.cuda.current_stream(dev))
            res.copy_(torch.cat([gz1, gz2]), non_blocking=True)     # device -> host read of the step's result
            out = res
        return out

    e2e_loop(max(2, args.warmup), 10_000)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_loop(args.steps, 20_000)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s fallback (of fallback)"
        # Dominant kernel family: the column kernels of the RS path (rs_cols: column FFT x transfer function x inverse
        # column FFT; rs_cols_gz: the same on the cotangent plus the column FFT of conj(U) and the Parseval sum for d/dz).
        # Algorithmic bytes per launch (u = 8*N^2 bytes = one N x N complex64 plane):
        #   rs_cols     per field: read 2u + write 2u of the N x L row spectra; per launch the transfer function, y-even: 2u
        #   rs_cols_gz  per field: read 2u (cotangent spectra) + 2u (conj-field spectra) + write 2u; per launch H and Hz: 4u
        u_bytes = 8.0 * N_GRID * N_GRID
        roof = None
        ncu_traffic = {}
        try:   # the latest committed ncu --set full capture (scripts/make_profiles.py writes one file per round)
            import glob
            files = sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_traffic_r*.json")))
            ncu_traffic = json.load(open(files[-1])) if files else {}
        except Exception:
            pass
        kname = max((kk for kk in ("rs_cols", "rs_cols_gz") if kk in kern), key=lambda kk: kern[kk][1], default=None)
        if kname and kname in kern1:
            cnt, tot = kern1[kname]                      # scalar-RS launches only: 1 field per launch
            cls = "XlRsCols" if kname == "rs_cols" else "XlRsColsGz"
            alg = (6.0 if kname == "rs_cols" else 10.0) * u_bytes
            avg_s = tot / cnt * 1e-3
            ach = alg / avg_s / 1e9
            roof = {"kernel": f"{kname} (xl_kernel<{cls}<4096>>, 1 field per launch)", "bound": "hbm",
                    "achieved": ach, "peak": peak_gbs, "unit": "GB/s", "frac": ach / peak_gbs,
                    "traffic": ncu_traffic.get(f"{cls}<4096>"),
                    "peak_source": peak_src, "alg_bytes_per_launch": alg, "avg_launch_us": avg_s * 1e6,
                    "share_of_step": kern[kname][1] / ms,
                    "note": "algorithmic bytes: 2u+2u (+2u conj-field spectra for _gz) of row spectra per field + 2u per "
                            "y-even transfer function (SURVEY 8d); the kernel is co-bound by fp32 FFT arithmetic "
                            "(fp32 pipe 40-50% busy, profiles/summary_r01.txt); traffic = ncu dram read+write of the same launch shape"}
        line = {
            "metric": METRIC, "value": world * PROPS_PER_STEP * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "c64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": "inputs larger than L2 (no flush needed)", "sharding": "independent batches per rank, no collective"},
            "clocks": clocks,
            "e2e": {"value": world * PROPS_PER_STEP * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 16,
                    "note": "pinned host inputs -> H2D (double-buffered on a copy stream) -> fwd+grad via the public API -> D2H of the z-gradients"},
            "gpu_launches": int(launches),
            "roofline": roof,
            "kernels_ms_total": {kk: round(v[1], 4) for kk, v in kern.items()},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            tt = cpu_step_times(threads, 5, mix=True)[1:]     # warm-up (RS), then VRS, CZT, VCZT, RS: one of each
            line["cpu_baseline"] = {"value": len(tt) / sum(tt), "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "one step of the same mix (scalar RS, VRS, CZT, VCZT forward+gradient at 2048^2) after "
                                              "1 warm-up propagation, torch-CPU complex128 restatement of the reference "
                                              "(oracle/oracle_torch.py) on all host threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
