"""Batch sharding over ranks (one process per GPU, torch.distributed).

Syndromes never interact (the PyG batch is block-diagonal), so
  * inference shards the batch contiguously across ranks with NO data-path collective;
  * training is data-parallel: one SUM all-reduce of the flat gradient vector per step
    (decoder_v2_4: 1 283 floats) -- NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous [start, end) of `n` items for `rank`; sizes differ by at most one and, when
    n >= 8 * world, every start is a multiple of 8 (output alignment of the kernels)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world: %r/%r" % (rank, world))
    if n >= 8 * world:
        units, rem = divmod(n, 8)
        base, extra = divmod(units, world)
        start = 8 * (rank * base + min(rank, extra))
        end = start + 8 * (base + (1 if rank < extra else 0))
        if rank == world - 1:
            end += rem
        return start, end
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def allreduce_flat_grads(params, group=None, average=True):
    """One collective for all gradients: flatten -> all_reduce(SUM) -> scatter back."""
    params = [p for p in params if p.grad is not None]
    if not params:
        return None
    _, world = world_info()
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= world
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat


def allreduce_counts(counts, group=None):
    """Sum failure counters (evaluate.count_failures) over ranks."""
    _, world = world_info()
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts
