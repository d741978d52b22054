"""Slab-decomposed RS (xlumina_b200/slab.py) on CPU: 2 and 4 gloo ranks drive the host-emulated kernel
bodies through the same Python code and the same C-ABI stage entry points as the GPU path; the reassembled result must
equal the complex128 oracle (and the single-rank library result)."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, emu_path, N, z, out_dir, max_line=4096, cluster=-1):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from xlumina_b200 import _lib, slab
        emu = _lib.declare(ctypes.CDLL(emu_path))
        emu.xl_debug_set_max_line(max_line)
        emu.xl_debug_set_long_cluster(cluster)
        rng = np.random.default_rng(7)
        field = (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))).astype(np.complex64)
        ct = (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))).astype(np.complex64)
        x = np.linspace(-300.0, 300.0, N)
        dx, k = float(x[1] - x[0]), 2 * np.pi / 0.6328
        rows = N // world
        mine = torch.as_tensor(field[rank * rows:(rank + 1) * rows].copy())
        out, H = slab.rs_propagation_slab(mine, z, dx, dx, k, lib=emu, return_transfer=True)
        ct_mine = torch.as_tensor(ct[rank * rows:(rank + 1) * rows].copy())
        vjp = slab.rs_slab_vjp(ct_mine, H, lib=emu)
        gz = slab.rs_slab_grad_z(mine, ct_mine, out, z, dx, dx, k, lib=emu)     # all-reduced: every rank holds the total
        np.save(os.path.join(out_dir, f"out{rank}.npy"), out.numpy())
        np.save(os.path.join(out_dir, f"vjp{rank}.npy"), vjp.numpy())
        np.save(os.path.join(out_dir, f"gz{rank}.npy"), gz.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,N,z", [(2, 32, 2500.0), (2, 48, -4000.0), (4, 64, 6000.0)])
def test_slab_rs_matches_oracle(emu, tmp_path, world, N, z):
    from conftest import rel_l2
    from oracle import oracle_np as o
    emu_path = os.path.join(ROOT, "tests", "emu", "libxlprop_emu.so")
    mp.spawn(_worker, args=(world, _free_port(), emu_path, N, z, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(7)
    field = (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))).astype(np.complex64)
    ct = (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))).astype(np.complex64)
    x = np.linspace(-300.0, 300.0, N)
    ref, _ = o.RS_propagation(field.astype(np.complex128), x, x, 0.6328, z)
    got = np.concatenate([np.load(tmp_path / f"out{r}.npy") for r in range(world)])
    assert rel_l2(got, ref) < 5e-6
    vref, _ = o.RS_propagation(ct.astype(np.complex128), x, x, 0.6328, z)     # A is complex-symmetric: A^T ct = A ct
    vgot = np.concatenate([np.load(tmp_path / f"vjp{r}.npy") for r in range(world)])
    assert rel_l2(vgot, vref) < 5e-6


@pytest.mark.parametrize("world,N,z", [(1, 32, 2500.0), (1, 128, 9000.0), (2, 64, -4000.0), (4, 128, 6000.0), (2, 24, 1500.0)])
def test_slab_rs_split_lines_match_oracle(emu, tmp_path, world, N, z):
    """The split-line kernels (csrc/xl_long.cuh: padded length = R x sub-line) with the sub-line length forced to 32, so that
    R = 2, 4 and 8 are reached at N = 24..128: the same code path as 4096 < padded length <= 32768 in production."""
    from conftest import rel_l2
    from oracle import oracle_np as o
    emu_path = os.path.join(ROOT, "tests", "emu", "libxlprop_emu.so")
    mp.spawn(_worker, args=(world, _free_port(), emu_path, N, z, str(tmp_path), 32), nprocs=world, join=True)
    rng = np.random.default_rng(7)
    field = (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))).astype(np.complex64)
    ct = (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))).astype(np.complex64)
    x = np.linspace(-300.0, 300.0, N)
    ref, _ = o.RS_propagation(field.astype(np.complex128), x, x, 0.6328, z)
    got = np.concatenate([np.load(tmp_path / f"out{r}.npy") for r in range(world)])
    assert rel_l2(got, ref) < 5e-6
    vref, _ = o.RS_propagation(ct.astype(np.complex128), x, x, 0.6328, z)
    vgot = np.concatenate([np.load(tmp_path / f"vjp{r}.npy") for r in range(world)])
    assert rel_l2(vgot, vref) < 5e-6


@pytest.mark.parametrize("world,N,z,cluster", [(1, 128, 9000.0, 1), (1, 64, 5000.0, 0), (2, 128, -4000.0, 1), (2, 24, 1500.0, 0)])
def test_slab_rs_split_lines_both_inverse_forms(emu, tmp_path, world, N, z, cluster):
    """The inverse radix step of the split kernels as a cluster kernel (forced on: also for 8 sub-lines) and as the two-launch
    form through the scratch buffer (forced off: also for 2 and 4 sub-lines); the default picks by the split factor."""
    from conftest import rel_l2
    from oracle import oracle_np as o
    emu_path = os.path.join(ROOT, "tests", "emu", "libxlprop_emu.so")
    mp.spawn(_worker, args=(world, _free_port(), emu_path, N, z, str(tmp_path), 32, cluster), nprocs=world, join=True)
    rng = np.random.default_rng(7)
    field = (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))).astype(np.complex64)
    x = np.linspace(-300.0, 300.0, N)
    ref, _ = o.RS_propagation(field.astype(np.complex128), x, x, 0.6328, z)
    got = np.concatenate([np.load(tmp_path / f"out{r}.npy") for r in range(world)])
    assert rel_l2(got, ref) < 5e-6


def test_split_lines_at_the_production_sub_line_length(emu):
    """The split-line kernels with the sub-line length of production (4096): N = 2304 -> padded length 8192 = 2 sub-lines (the
    cluster form of the inverse step), one rank, against the complex128 oracle.  The other tests force sub-lines of 32."""
    from conftest import rel_l2
    from oracle import oracle_np as o
    from xlumina_b200 import slab
    o.set_workers(8)
    emu.xl_debug_set_max_line(4096)
    emu.xl_debug_set_long_cluster(-1)
    N, lam, z = 2304, 0.6328, 50000.0
    x = np.linspace(-15000.0, 15000.0, N)
    rng = np.random.default_rng(2304)
    X, Y = np.meshgrid(x, x)
    f = (np.exp(-(X ** 2 + Y ** 2) / 4000.0 ** 2) * (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N)))).astype(np.complex64)
    assert slab.SlabPlan(N, 1, emu).L == 8192
    out = slab.rs_propagation_slab(torch.as_tensor(f), z, float(x[1] - x[0]), float(x[1] - x[0]), 2 * np.pi / lam,
                                   group=slab._LOCAL, lib=emu)
    ref, _ = o.RS_propagation(f.astype(np.complex128), x, x, lam, z)
    assert rel_l2(out.numpy(), ref) < 5e-6


def _grad_z_oracle(field, ct, x, z):
    """d/dz Re sum(ct * RS(field, z)) by autograd through the complex128 torch oracle."""
    from oracle import oracle_torch as ot
    zt = torch.tensor(float(z), dtype=torch.float64, requires_grad=True)
    out = ot.RS_propagation(torch.as_tensor(field.astype(np.complex128)), x, x, 0.6328, zt)
    (torch.as_tensor(ct.astype(np.complex128)) * out).real.sum().backward()
    return float(zt.grad)


@pytest.mark.parametrize("world,N,z,max_line", [(2, 32, 2500.0, 4096), (4, 64, 6000.0, 4096), (2, 64, -4000.0, 32), (4, 128, 6000.0, 32)])
def test_slab_grad_z_matches_oracle(emu, tmp_path, world, N, z, max_line):
    """d/dz on the slab path (slab.rs_slab_grad_z: per-rank partial sums + one all-reduce), single-pass and split-line
    kernels, against autograd through the complex128 oracle."""
    emu_path = os.path.join(ROOT, "tests", "emu", "libxlprop_emu.so")
    mp.spawn(_worker, args=(world, _free_port(), emu_path, N, z, str(tmp_path), max_line), nprocs=world, join=True)
    rng = np.random.default_rng(7)
    field = (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))).astype(np.complex64)
    ct = (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))).astype(np.complex64)
    ref = _grad_z_oracle(field, ct, np.linspace(-300.0, 300.0, N), z)
    got = [float(np.load(tmp_path / f"gz{r}.npy")) for r in range(world)]
    assert max(got) == min(got)                      # the all-reduce left the same total on every rank
    assert abs(got[0] - ref) < 1e-4 * abs(ref)


def test_slab_plan_rejects_bad_partitions(emu):
    from xlumina_b200 import slab, _lib
    with pytest.raises(_lib.XlpropError):
        slab.SlabPlan(30, 4, emu)          # 30 rows cannot be split into 4 slabs of whole row pairs
    with pytest.raises(_lib.XlpropError):
        slab.SlabPlan(20000, 2, emu)       # padded length 65536 > 32768
    p = slab.SlabPlan(2048, 8, emu)
    assert (p.L, p.rows, p.pairs) == (4096, 256, 256) and p.hrows * 8 >= 2049 and p.hrows % 2 == 0
    p = slab.SlabPlan(16384, 8, emu)       # BASELINE.json cfg 5
    assert (p.L, p.rows, p.pairs) == (32768, 2048, 2048) and p.scratch_bytes == 2048 * 32768 * 2 * 8
