"""Batch sharding across ranks: host logic on CPU with world_size-2 gloo.  Every rank owns a contiguous
slice of the candidate batch, no collective on the data path, one flattened all-reduce for shared-parameter gradients."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xlumina_b200.sharding import allreduce_grads, shard, shard_range


@pytest.mark.parametrize("n,world", [(64, 8), (10, 4), (3, 8), (0, 2), (7, 1)])
def test_shard_range_partitions_exactly(n, world):
    spans = [shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a0, b0), (a1, b1) in zip(spans, spans[1:]):
        assert b0 == a1 and b0 >= a0
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n


def test_shard_range_rejects_bad_rank():
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # a batch of 5 "candidate set-ups" with a parameter shared by all of them (the 4f optimizer pattern)
        batch = torch.arange(5 * 4, dtype=torch.float64).reshape(5, 4)
        mine = shard(batch)                                  # rank/world from the process group
        p = torch.tensor([1.5, -2.0], dtype=torch.float64, requires_grad=True)
        pc = torch.tensor([0.5 + 1.0j], dtype=torch.complex128, requires_grad=True)
        loss = (mine.sum(dim=1) * p[0]).sum() + p[1] * mine.shape[0] + (pc * pc.conj()).real.sum() * mine.shape[0]
        loss.backward()
        g = allreduce_grads([p.grad, pc.grad])
        np.save(os.path.join(out_dir, f"g{rank}.npy"), g[0].numpy())
        np.save(os.path.join(out_dir, f"gc{rank}.npy"), g[1].numpy())
        np.save(os.path.join(out_dir, f"n{rank}.npy"), np.array(mine.shape[0]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shared_parameter_gradient(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    batch = np.arange(20, dtype=np.float64).reshape(5, 4)
    expect = np.array([batch.sum(), 5.0])                    # single-process gradient of the same loss over the WHOLE batch
    g0, g1 = np.load(tmp_path / "g0.npy"), np.load(tmp_path / "g1.npy")
    assert np.allclose(g0, expect) and np.allclose(g1, expect)
    gc0 = np.load(tmp_path / "gc0.npy")
    assert np.allclose(gc0, 5 * 2 * (0.5 + 1.0j))            # torch convention for |pc|^2: 2*pc per unit
    assert int(np.load(tmp_path / "n0.npy")) + int(np.load(tmp_path / "n1.npy")) == 5


def _four_f_worker(rank, world, port, emu_path, out_dir):
    """BASELINE config 4 pattern on the real operators (host-emulated kernels): each rank evaluates the 4f table on its
    slice of the masks with the SHARED parameters, gradients meet in one flattened all-reduce."""
    import ctypes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from conftest import golden
        from test_elements import four_f_problem
        from xlumina_b200 import _lib, four_f, ops
        _lib._lib = _lib.declare(ctypes.CDLL(emu_path))          # test tooling: CPU tensors through the emulated kernels
        ops._require_device = lambda t: None
        ops._stream = lambda t: ctypes.c_void_p(0)
        ops._stream_key = lambda t: ("cpu", 0)
        g = golden("four_f_n32")
        src, params, masks, targets = four_f_problem(g, "cpu", torch.complex64)
        B = masks.shape[0]
        a, b = shard_range(B, rank, world)
        inten, _, _ = four_f.vector_dualSLM_4f_system(masks[a:b], src, params)
        loss = four_f.MSE_Intensity(inten, targets[a:b]).sum() / B        # mean over the GLOBAL batch
        loss.backward()
        grads = allreduce_grads([p.grad for p in params])
        total = loss.detach().clone()
        dist.all_reduce(total)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), loss=total.numpy(), **{f"g{i}": gr.numpy() for i, gr in enumerate(grads)})
    finally:
        dist.destroy_process_group()


def test_two_rank_four_f_table_equals_single_process(emu, tmp_path):
    import ctypes
    from conftest import golden, ROOT
    emu_path = os.path.join(ROOT, "tests", "emu", "libxlprop_emu.so")
    for world in (1, 2):
        d = tmp_path / f"w{world}"
        d.mkdir()
        mp.spawn(_four_f_worker, args=(world, _free_port(), emu_path, str(d)), nprocs=world, join=True)
    one = np.load(tmp_path / "w1" / "r0.npz")
    g = golden("four_f_n32")
    assert abs(float(one["loss"]) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))      # and equals the reference's loss
    for rank in (0, 1):
        two = np.load(tmp_path / "w2" / f"r{rank}.npz")
        assert abs(float(two["loss"]) - float(one["loss"])) < 1e-6 * abs(float(one["loss"]))
        for i in range(5):
            # same kernels, different batch grouping: distances see fp32 summation-order noise of the cancelling terms only
            tol = 2e-3 if i < 3 else 1e-5
            assert np.linalg.norm(two[f"g{i}"] - one[f"g{i}"]) <= tol * np.linalg.norm(one[f"g{i}"]), i


def test_flat_grad_buffers_are_accumulated_in_place():
    """sharding.flat_grad_buffers: every .grad is a view of one flat buffer per dtype, autograd accumulates into it in place
    and zero_grad(set_to_none=False) keeps the views -- so the step's all-reduce runs on the flat buffers directly."""
    import torch
    from xlumina_b200.sharding import allreduce_flat, flat_grad_buffers
    a = torch.randn(4, 5, requires_grad=True)
    b = torch.randn(3, dtype=torch.float64, requires_grad=True)
    c = torch.randn(7, requires_grad=True)
    flats = flat_grad_buffers([a, b, c])
    assert sorted(f.numel() for f in flats) == [3, 27]
    ptrs = [p.grad.data_ptr() for p in (a, b, c)]
    opt = torch.optim.AdamW([a, b, c], lr=0.01)
    for _ in range(2):
        opt.zero_grad(set_to_none=False)
        ((a ** 2).sum() + (b ** 3).sum() + c.sum()).backward()
        assert [p.grad.data_ptr() for p in (a, b, c)] == ptrs
        f32 = next(f for f in flats if f.dtype == torch.float32)
        assert torch.allclose(f32[:20].view(4, 5), 2 * a.detach()) and torch.allclose(f32[20:], torch.ones(7))
        allreduce_flat(flats)          # no process group: a no-op
        opt.step()
