"""Fused training step for the decoder_v2_4 program (BASELINE config 4): forward-with-stash kernel ->
fused sparse loss + gradient kernel (LossFunc of quantum/decoder_v2_4.py:304-317) -> hand-written backward
kernel -> parameter .grad, with no autograd graph and no dense H matmul.  `train_step_grads` stops at the gradients
(any torch optimizer can follow); `FusedTrainer` also runs the reference's optimizer (Adam, decoder_v2_4.py:323) as
a kernel on the flat fp32 master weights -- fused with the peer-memory gradient all-reduce when data-parallel."""
import ctypes as ct

import numpy as np
import torch

from . import _cabi


def _ptr(t):
    return ct.c_void_p(t.data_ptr()) if t is not None else None


def _logical_dev(graph, logical):
    if logical is None:
        return None, 0
    key = np.ascontiguousarray(np.asarray(logical.detach().cpu() if torch.is_tensor(logical) else logical, dtype=np.uint8))
    if key.shape[1] != graph.V:
        raise ValueError("logical must be [K, V=%d]" % graph.V)
    ldev = graph._logical_dev.get(key.tobytes())
    if ldev is None:
        ldev = torch.from_numpy(key).to(graph.device)
        graph._logical_dev[key.tobytes()] = ldev
    return ldev, key.shape[0]


def sin_loss(graph, prob, y, logical=None, want_grad_prob=False, want_grad_logit=True):
    """LossFunc.forward(pred, datas) of decoder_v2_4.py:304-317 on prob [B, V] fp32 CUDA and y [B, V] uint8
    (0/1 errors), fused with its gradient.  Returns (loss [scalar tensor], grad_prob or None, grad_logit or None)."""
    if not prob.is_cuda:
        raise _cabi.GdError("prob is on %s: gnn_decode_b200 has no CPU fallback" % prob.device)
    B, V = prob.shape
    if V != graph.V:
        raise ValueError("prob must be [B, V=%d]" % graph.V)
    dev = prob.device
    prob = prob.detach().to(torch.float32).contiguous()
    y8 = y.detach().reshape(B, V).to(device=dev, dtype=torch.uint8).contiguous()
    ldev, K = _logical_dev(graph, logical)
    per = torch.empty(B, dtype=torch.float32, device=dev)
    gp = torch.empty((B, V), dtype=torch.float32, device=dev) if want_grad_prob else None
    gl = torch.empty((B, V), dtype=torch.float32, device=dev) if want_grad_logit else None
    st = ct.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().gd_loss_v2_4(graph.handle, _ptr(ldev), K, _ptr(prob), _ptr(y8), _ptr(per), _ptr(gp), _ptr(gl),
                                             B, st), "gd_loss_v2_4")
    return per.double().sum(), gp, gl


class _SinLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prob, graph, y, logical):
        loss, gp, _ = sin_loss(graph, prob, y, logical, want_grad_prob=True, want_grad_logit=False)
        ctx.save_for_backward(gp)
        ctx.dtype = prob.dtype
        return loss.to(prob.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        (gp,) = ctx.saved_tensors
        return (gp * grad_out).to(ctx.dtype), None, None, None


class LossFunc(torch.nn.Module):
    """Drop-in for `LossFunc(H, H_prep)` of decoder_v2_4.py:297-317.  The reference reads the dense `H` and the
    module global `logical`; here the sparse tables of `graph` and the `logical` argument are used.
    forward(pred, datas): pred [B*V, 1] as returned by GNNI.forward, datas.y [B*V, 1] -> scalar loss (autograd ok)."""

    def __init__(self, H=None, H_prep=None, *, graph, logical=None):
        super(LossFunc, self).__init__()
        self.graph, self.logical = graph, logical

    def forward(self, pred, datas):
        V = self.graph.V
        return _SinLossFn.apply(pred.reshape(-1, V), self.graph, datas.y.reshape(-1, V), self.logical)


def train_step_grads(decoder, graph, x, y, logical=None, accumulate=False, p2p=None):
    """One forward + loss + backward of decoder_v2_4.GNNI on a batch: x [B, V+C] CUDA, y [B, V] 0/1.
    Sets / accumulates p.grad of the decoder's MLP parameters and returns (loss, prob).
    p2p: a dist.P2PAllReduce -- data-parallel training: the flat gradient is averaged over the ranks by the
    peer-memory kernel before it is scattered into p.grad (no NCCL call, no flatten / unflatten round trip)."""
    if decoder._gd_program != _cabi.PROG_V2_4:
        raise _cabi.GdError("the training kernels exist for the decoder_v2_4 program only")
    lib = _cabi.lib()
    dev = x.device
    B = x.size(0)
    params = decoder._gd_params()
    model = decoder.gd_model()
    x32 = x.detach().to(torch.float32).contiguous()
    if x32.data_ptr() % 16:
        x32 = x32.clone()
    w = decoder.packed_weights(dev)
    n_stash = lib.gd_stash_floats(graph.handle, ct.byref(model), B)
    n_ws = lib.gd_bwd_workspace_floats(graph.handle, ct.byref(model), B)
    if n_stash < 0 or n_ws < 0:
        _cabi.check(_cabi.GD_ERR_INVALID, "gd_stash_floats / gd_bwd_workspace_floats")
    stash = torch.empty(max(n_stash, 1), dtype=torch.float32, device=dev)
    ws = torch.empty(max(n_ws, 1), dtype=torch.float32, device=dev)
    prob = torch.empty((B, graph.V), dtype=torch.float32, device=dev)
    logit = torch.empty((B, graph.V), dtype=torch.float32, device=dev)
    gw = torch.empty(w.numel(), dtype=torch.float32, device=dev)
    st = ct.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        _cabi.check(lib.gd_decode_fwd_train(graph.handle, ct.byref(model), _ptr(w), _ptr(x32), _ptr(prob), _ptr(logit),
                                            _ptr(stash), B, st), "gd_decode_fwd_train")
    loss, _, grad_logit = sin_loss(graph, prob, y, logical)
    with torch.cuda.device(dev):
        _cabi.check(lib.gd_decode_bwd(graph.handle, ct.byref(model), _ptr(w), _ptr(x32), _ptr(stash), _ptr(grad_logit),
                                      _ptr(gw), _ptr(ws), 0, B, st), "gd_decode_bwd")
    if p2p is not None:
        p2p(gw, average=True)
    off = 0
    for p in params:
        n = p.numel()
        g = gw[off:off + n].view(p.shape).to(p.dtype)
        if accumulate and p.grad is not None:
            p.grad += g
        else:
            p.grad = g.clone()
        off += n
    return loss, prob


class FusedTrainer(object):
    """The body of the reference's training loop (quantum/decoder_v2_4.py:325-341: `pred = decoder(datas)`,
    `loss = criterion(pred, datas)`, `loss.backward()`, `optimizer.step()` with `Adam(lr=3e-4, weight_decay=1e-9)`, :323)
    with every stage a kernel of this library and nothing in between: forward-with-stash -> sparse loss + gradient ->
    backward -> Adam on the flat fp32 master weights (gd_adam_step), or, data-parallel, the peer-memory all-reduce with
    Adam fused into it (gd_p2p_allreduce_adam).  Five launches per step; all buffers are allocated once per batch size.

    The master weights live in `self.w` (the packed layout the decode kernels read); `sync_module()` writes them back
    into the decoder's parameters (fp64 in the reference's state_dict) -- call it before `state_dict()` / evaluation
    through the module.  Data-parallel (an initialised process group with world_size > 1): the flat gradient is summed over
    the ranks by p2p (a dist.P2PAllReduce, one kernel with Adam fused in) or, without it, by ONE NCCL all-reduce; rank 0's
    weights and moments are broadcast at construction so every replica starts, and therefore stays, bit-identical.
    average: divide the summed gradient by world_size.  The reference's loss is a batch SUM (decoder_v2_4.py:314-315), so
    average=False reproduces `W ranks x batch B == one GPU x batch W B` exactly; the default True keeps the per-step
    gradient scale independent of the number of GPUs (Adam is nearly scale-invariant: only eps = 1e-8 and
    weight_decay = 1e-9 see the difference).  check_every: synchronise and raise if a peer missed an all-reduce, every that
    many steps (and in sync_module())."""

    def __init__(self, decoder, graph, logical=None, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-9, p2p=None,
                 average=True, check_every=256):
        if decoder._gd_program != _cabi.PROG_V2_4:
            raise _cabi.GdError("the training kernels exist for the decoder_v2_4 program only")
        self.decoder, self.graph, self.logical, self.p2p = decoder, graph, logical, p2p
        dev = graph.device
        with torch.no_grad():
            self.w = torch.cat([p.detach().reshape(-1).to(device=dev, dtype=torch.float32) for p in decoder._gd_params()]).contiguous()
        self.exp_avg = torch.zeros_like(self.w)
        self.exp_avg_sq = torch.zeros_like(self.w)
        self.grad = torch.zeros_like(self.w)
        self.adam = _cabi.GdAdam(lr, betas[0], betas[1], eps, weight_decay, 0)
        self._bufs = {}
        self.average, self.check_every = bool(average), int(check_every)
        import torch.distributed as dist
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        if self.world > 1:
            for t in (self.w, self.exp_avg, self.exp_avg_sq):
                dist.broadcast(t, src=0)

    @property
    def step_count(self):
        return int(self.adam.step)

    def _buffers(self, B):
        b = self._bufs.get(B)
        if b is None:
            lib, g, dev = _cabi.lib(), self.graph, self.graph.device
            model = self.decoder.gd_model()
            n_stash = lib.gd_stash_floats(g.handle, ct.byref(model), B)
            n_ws = lib.gd_bwd_workspace_floats(g.handle, ct.byref(model), B)
            if n_stash < 0 or n_ws < 0:
                _cabi.check(_cabi.GD_ERR_INVALID, "gd_stash_floats / gd_bwd_workspace_floats")
            f32 = dict(dtype=torch.float32, device=dev)
            b = dict(stash=torch.empty(max(n_stash, 1), **f32), ws=torch.empty(max(n_ws, 1), **f32),
                     prob=torch.empty((B, g.V), **f32), logit=torch.empty((B, g.V), **f32),
                     per=torch.empty(B, **f32), gl=torch.empty((B, g.V), **f32))
            self._bufs = {B: b}                      # one batch size at a time: the stash is the big allocation
        return b

    def step(self, x, y):
        """x [B, V+C] fp32 CUDA, y [B, V] uint8 CUDA (the sampled error).  One optimizer step; returns the batch loss
        (0-dim fp64 CUDA tensor, this rank's shard) -- `prob` of the step stays in `self.prob`."""
        lib, g, dev = _cabi.lib(), self.graph, self.graph.device
        if not x.is_cuda or x.dtype != torch.float32 or not x.is_contiguous() or x.data_ptr() % 16:
            x = x.detach().to(device=dev, dtype=torch.float32).contiguous().clone()
        B = x.size(0)
        y8 = y if (y.dtype == torch.uint8 and y.is_contiguous() and y.is_cuda) else y.detach().reshape(B, g.V).to(device=dev, dtype=torch.uint8).contiguous()
        b = self._buffers(B)
        model = self.decoder.gd_model()
        ldev, K = _logical_dev(g, self.logical)
        st = ct.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            _cabi.check(lib.gd_decode_fwd_train(g.handle, ct.byref(model), _ptr(self.w), _ptr(x), _ptr(b["prob"]), _ptr(b["logit"]),
                                                _ptr(b["stash"]), B, st), "gd_decode_fwd_train")
            _cabi.check(lib.gd_loss_v2_4(g.handle, _ptr(ldev), K, _ptr(b["prob"]), _ptr(y8), _ptr(b["per"]), None, _ptr(b["gl"]),
                                         B, st), "gd_loss_v2_4")
            _cabi.check(lib.gd_decode_bwd(g.handle, ct.byref(model), _ptr(self.w), _ptr(x), _ptr(b["stash"]), _ptr(b["gl"]),
                                          _ptr(self.grad), _ptr(b["ws"]), 0, B, st), "gd_decode_bwd")
            self.adam.step += 1
            scale = 1.0 / self.world if self.average else 1.0
            if self.p2p is not None:
                self.p2p.allreduce_adam(self.grad, self.adam, self.w, self.exp_avg, self.exp_avg_sq, average=self.average)
                if self.check_every > 0 and self.adam.step % self.check_every == 0:
                    self.p2p.check()
            else:
                if self.world > 1:                 # the path's only exchange: one NCCL all-reduce of the flat gradient
                    import torch.distributed as dist
                    dist.all_reduce(self.grad, op=dist.ReduceOp.SUM)
                _cabi.check(lib.gd_adam_step(ct.byref(self.adam), _ptr(self.w), _ptr(self.grad), _ptr(self.exp_avg),
                                             _ptr(self.exp_avg_sq), self.w.numel(), scale if self.world > 1 else 1.0, st), "gd_adam_step")
        self.prob = b["prob"]
        return b["per"].double().sum()

    def sync_module(self):
        """Write the master weights back into the decoder's parameters (state_dict keys / dtypes unchanged)."""
        if self.p2p is not None:
            self.p2p.check()
        off = 0
        with torch.no_grad():
            for p in self.decoder._gd_params():
                n = p.numel()
                p.copy_(self.w[off:off + n].view(p.shape).to(p.dtype))
                off += n
        return self.decoder
