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

__all__ = ["shard_range", "shard", "allreduce_grads"]


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
