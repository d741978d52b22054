"""
Batch sharding of independent propagation units across the GPUs of one box.

The reference batches candidate set-ups / input masks / noisy tables with `jax.vmap` on a single device
(`experiments/four_f_optical_table.py:98`, `examples/noisy_optimization.ipynb` cell 7).  Here the same axis is split
across ranks (one process per GPU, `torchrun`): every rank propagates its own contiguous slice and the data path needs
NO collective.  Only when the candidates share parameters (the 4f optimizer's two phase masks and three distances) does
the caller sum the parameter gradients across ranks once per optimizer step: `allreduce_grads`.

Host logic only (torch.distributed is plumbing); works with the gloo backend on CPU for tests.
"""
import torch
import torch.distributed as dist

__all__ = ["shard_range", "shard", "allreduce_grads", "flat_grad_buffers", "allreduce_flat"]


def shard_range(n_units, rank, world):
    """[start, stop) of the units owned by `rank`: contiguous blocks whose sizes differ by at most one
    (the first n_units % world ranks own one more)."""
    if n_units < 0 or world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad shard request: n_units={n_units} rank={rank} world={world}")
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard(batch, rank=None, world=None, dim=0):
    """The slice of `batch` (tensor or sequence) owned by this rank along `dim`."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    n = batch.shape[dim] if isinstance(batch, torch.Tensor) else len(batch)
    a, b = shard_range(n, rank, world)
    if isinstance(batch, torch.Tensor):
        return batch.narrow(dim, a, b - a)
    return batch[a:b]


def allreduce_grads(grads, average=False):
    """Sum (or average) shared-parameter gradients over all ranks, flattened into ONE collective per dtype so that the
    cost is one launch latency, not one per parameter.  No-op without an initialised process group."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return grads
    by_dtype = {}
    for g in grads:
        by_dtype.setdefault((g.dtype, g.device), []).append(g)
    for (_, _), gs in by_dtype.items():
        reals = [torch.view_as_real(g) if g.is_complex() else g for g in gs]
        flat = torch.cat([r.reshape(-1) for r in reals])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat /= dist.get_world_size()
        off = 0
        for r in reals:
            r.copy_(flat[off:off + r.numel()].view_as(r))
            off += r.numel()
    return grads


def flat_grad_buffers(params):
    """Give every parameter a persistent `.grad` that is a VIEW of one flat buffer per dtype (autograd accumulates into an
    existing .grad in place, optimizers read it in place): the per-step gradient all-reduce is then one collective per dtype
    on the flat buffers themselves -- no concatenation before and no scatter after (`allreduce_flat`).  Returns the buffers."""
    groups = {}
    for p in params:
        groups.setdefault((p.dtype, p.device), []).append(p)
    flats = []
    for (dtype, device), ps in groups.items():
        flat = torch.zeros(sum(p.numel() for p in ps), dtype=dtype, device=device)
        off = 0
        for p in ps:
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        flats.append(flat)
    return flats


def allreduce_flat(flats, average=False):
    """Sum (or average) the flat gradient buffers of `flat_grad_buffers` over all ranks; no-op without a process group."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return flats
    for flat in flats:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat /= dist.get_world_size()
    return flats
