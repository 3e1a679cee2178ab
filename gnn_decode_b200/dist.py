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


class P2PAllReduce(object):
    """One-shot gradient all-reduce over NVLink / NVSwitch peer memory (csrc/gd_p2p.cu): one kernel per step that
    publishes the local flat gradient, waits for the peers' epoch flags and sums every rank's vector in rank order
    straight from peer memory.  The symmetric buffers and their peer mappings come from
    torch.distributed._symmetric_memory; everything on the data path is the kernel.  Results are bit-identical on all
    ranks.  Falls back to nothing: construction raises if symmetric memory cannot be set up (callers then use
    allreduce_flat_grads / NCCL)."""

    def __init__(self, n_floats, device, group=None):
        import ctypes as ct
        import torch.distributed._symmetric_memory as symm_mem
        from . import _cabi
        self._ct, self._cabi = ct, _cabi
        self.n = int(n_floats)
        self.rank, self.world = world_info()
        if self.world < 2:
            raise ValueError("P2PAllReduce needs an initialised process group with world_size >= 2")
        group = group if group is not None else dist.group.WORLD
        n_buf = int(_cabi.lib().gd_p2p_buffer_floats(self.n))
        self.buf = symm_mem.empty(n_buf, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group.group_name if hasattr(group, "group_name") else group)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)                 # every rank's flags are zero before the first epoch
        self.ptrs = (ct.c_uint64 * self.world)(*[int(p) for p in self.handle.buffer_ptrs])
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.epoch = 0
        self.device = device

    def __call__(self, flat_in, flat_out=None, average=True):
        """flat_in: contiguous fp32 CUDA vector of n floats -> flat_out (default: in place) = sum (or mean) over ranks."""
        ct = self._ct
        if flat_in.numel() != self.n or flat_in.dtype != torch.float32 or not flat_in.is_contiguous():
            raise ValueError("flat_in must be a contiguous fp32 vector of %d elements" % self.n)
        out = flat_in if flat_out is None else flat_out
        self.epoch += 1
        st = ct.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            self._cabi.check(self._cabi.lib().gd_p2p_allreduce(
                self.ptrs, self.world, self.rank, ct.c_void_p(flat_in.data_ptr()), ct.c_void_p(out.data_ptr()), self.n,
                ct.c_uint32(self.epoch), ct.c_float(1.0 / self.world if average else 1.0), ct.c_void_p(self.err.data_ptr()), st),
                "gd_p2p_allreduce")
        return out

    def allreduce_adam(self, flat_grad, adam, weights, exp_avg, exp_avg_sq, flat_out=None, average=True):
        """The exchange step and the optimizer step as ONE kernel (gd_p2p_allreduce_adam): the flat gradient is summed
        (or averaged) over the ranks from peer memory and torch.optim.Adam's update is applied to the flat fp32 master
        `weights` / moments in place.  adam: a _cabi.GdAdam with .step already advanced.  flat_out (optional) receives the
        reduced gradient."""
        ct = self._ct
        for t in (flat_grad, weights, exp_avg, exp_avg_sq):
            if t.numel() != self.n or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("flat gradient / weights / moments must be contiguous fp32 vectors of %d elements" % self.n)
        self.epoch += 1
        st = ct.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            self._cabi.check(self._cabi.lib().gd_p2p_allreduce_adam(
                self.ptrs, self.world, self.rank, ct.c_void_p(flat_grad.data_ptr()),
                ct.c_void_p(flat_out.data_ptr()) if flat_out is not None else None, self.n,
                ct.c_uint32(self.epoch), ct.c_float(1.0 / self.world if average else 1.0), ct.c_void_p(self.err.data_ptr()),
                ct.byref(adam), ct.c_void_p(weights.data_ptr()), ct.c_void_p(exp_avg.data_ptr()),
                ct.c_void_p(exp_avg_sq.data_ptr()), st), "gd_p2p_allreduce_adam")
        return weights

    def check(self):
        """Synchronises and raises if a peer failed to arrive in any previous call."""
        if int(self.err.item()) != 0:
            raise self._cabi.GdError("gd_p2p_allreduce: a peer rank never raised its epoch flag")


def allreduce_flat_grads_p2p(params, p2p, average=True):
    """allreduce_flat_grads through the peer-memory kernel: flatten -> one kernel -> scatter back."""
    params = [p for p in params if p.grad is not None]
    if not params:
        return None
    flat = torch.cat([p.grad.reshape(-1).to(torch.float32) for p in params]).contiguous()
    p2p(flat, average=average)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat
