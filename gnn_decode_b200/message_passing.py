"""Host-side mirror of the reference's operator interface for the message-passing hot path.

The reference copies a hand-edited fork of PyG's MessagePassing into every script
(quantum/decoder_v2_4.py:66-158, classical/CGNNI.py:33-122, ...): `propagate()` hard-codes a
flow-specific phase program
    pre    : tanh(m/2) for flow='target_to_source' (or the sum-product `log|tanh|` for BP)
    reduce : scatter_add(m, edge_index[j])[edge_index[j]] - m       ("sum over siblings minus self")
    post   : + extra[edge_index[j]]   or   cat[., extra[edge_index[j]]]
    update : self.update(out)         (per-edge MLP / multiply by the syndrome sign / identity)
over EDGE-RESIDENT messages m [B*E, 1] of a PyG-batched block-diagonal Tanner graph.

Here the same names, positional signatures, argument meaning and error behaviour are kept, but
pre/reduce/post (and the built-in update) run as ONE hand-written CUDA kernel through the C ABI
(gd_propagate_fwd), and GNNI.forward runs the whole Nc-iteration loop as ONE persistent fused
kernel (gd_decode_fwd).  There is no CPU fallback: CPU tensors raise.
"""
import ctypes as C
import inspect

import torch

from . import _cabi
from .graph import TannerGraph, graph_from_batched

special_args = ['edge_index', 'edge_index_i', 'edge_index_j', 'size', 'size_i', 'size_j']
__size_error_msg__ = ('All tensors which should get mapped to the same source'
                      'or target nodes must be of same size in dimension 0.')


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t, what):
    if not t.is_cuda:
        raise _cabi.GdError("%s is on %s: gnn_decode_b200 has no CPU fallback, move it to a CUDA device"
                            % (what, t.device))


def pack_mlp(seq):
    """Flatten Sequential(Linear(k,h), act, Linear(h,1)) in state_dict order."""
    return [seq[0].weight, seq[0].bias, seq[2].weight, seq[2].bias]


class _PackedWeights(object):
    """fp32 packed copy of a list of parameters, rebuilt only when a parameter changed."""

    def __init__(self):
        self._key = None
        self._buf = None

    def get(self, params, device):
        key = tuple((p.data_ptr(), p._version, str(p.device), p.dtype) for p in params) + (str(device),)
        if key != self._key:
            with torch.no_grad():
                flat = [p.detach().reshape(-1).to(device=device, dtype=torch.float32) for p in params]
                self._buf = torch.cat(flat).contiguous() if flat else torch.zeros(1, device=device)
            self._key = key
        return self._buf


class MessagePassingBase(torch.nn.Module):
    """Drop-in for the reference's forked `MessagePassing(aggr='add', flow='source_to_target')`.

    Subclasses set `_gd_program` (which script's phase program `propagate` hard-codes) and may
    override `message(x_j)` / `update(aggr_out)` exactly as with the reference.
    """
    _gd_program = None
    _gd_builtin_update = None     # the class's own fused update, if any

    def __init__(self, aggr='add', flow='source_to_target'):
        super(MessagePassingBase, self).__init__()
        self.aggr = aggr
        assert self.aggr in ['add', 'mean', 'max']
        self.flow = flow
        assert self.flow in ['source_to_target', 'target_to_source']
        msg_args = inspect.getfullargspec(self.message)[0][1:]
        self.__special_args__ = [(i, a) for i, a in enumerate(msg_args) if a in special_args]
        self.__message_args__ = [a for a in msg_args if a not in special_args]
        self.__update_args__ = inspect.getfullargspec(self.update)[0][2:]
        self._gd_rows = None
        self._gd_cols = None
        self._gd_packed = _PackedWeights()

    # -- code binding (replaces the reference's module globals rows / cols / BATCH_SIZE) --------
    def bind_code(self, rows, cols):
        self._gd_rows, self._gd_cols = (None if rows is None else int(rows)), (None if cols is None else int(cols))
        return self

    # -- overridable hooks, as in the reference ----------------------------------------------------
    def message(self, x_j):  # pragma: no cover
        return x_j

    def update(self, aggr_out):  # pragma: no cover
        return aggr_out

    # -- pieces the per-script subclasses configure ---------------------------------------------------
    def _gd_hidden(self):
        return 0

    def _gd_update_params(self):
        """Parameters of the fused built-in update of THIS flow (packed MLP), or []."""
        return []

    def _gd_update_is_builtin(self):
        cls = type(self)
        return cls._gd_builtin_update is not None and cls.update is cls._gd_builtin_update

    # -- the operator -------------------------------------------------------------------------------
    def _propagate(self, edge_index, node_in, size, kwargs):
        if self.aggr != 'add':
            # the reference only ever instantiates aggr='add' (SURVEY 2.2); mean/max are dead code there
            raise _cabi.GdError("aggr=%r has no CUDA kernel: only 'add' is on the reference's hot path" % self.aggr)
        size_given = size is not None and size[0] is not None
        size = [None, None] if size is None else list(size)
        assert len(size) == 2
        i, j = (0, 1) if self.flow == 'target_to_source' else (1, 0)
        ij = {"_i": i, "_j": j}

        # resolve message() arguments exactly as the reference does (no gather: messages are edge-resident)
        message_args = []
        for arg in self.__message_args__:
            if arg[-2:] in ij:
                tmp = kwargs[arg[:-2]]
                if tmp is not None:
                    idx = ij[arg[-2:]]
                    if isinstance(tmp, (tuple, list)):
                        assert len(tmp) == 2
                        if size[1 - idx] is None:
                            size[1 - idx] = tmp[1 - idx].size(0)
                        if size[1 - idx] != tmp[1 - idx].size(0):
                            raise ValueError(__size_error_msg__)
                        tmp = tmp[idx]
                    if size[idx] is None:
                        size[idx] = tmp.size(0)
                message_args.append(tmp)
            else:
                message_args.append(kwargs[arg])
        size[0] = size[1] if size[0] is None else size[0]
        size[1] = size[0] if size[1] is None else size[1]
        kwargs['edge_index'] = edge_index
        kwargs['size'] = size
        for (pos, arg) in self.__special_args__:
            if arg[-2:] in ij:
                message_args.insert(pos, kwargs[arg[:-2]][ij[arg[-2:]]])
            else:
                message_args.insert(pos, kwargs[arg])
        update_args = [kwargs[arg] for arg in self.__update_args__]

        m = self.message(*message_args)
        _require_cuda(m, "message tensor")
        _require_cuda(edge_index, "edge_index")

        n_edges = edge_index.size(1)
        if m.dim() != 2 or m.size(0) != n_edges or m.size(1) != 1:
            raise ValueError("messages must be edge-resident [B*E, 1]=%s, got %s" % ((n_edges, 1), tuple(m.shape)))
        if node_in is not None:
            n_rows = node_in.size(0)
        elif size_given and self._gd_rows is None:
            n_rows = int(size[0])
        else:
            n_rows = None
        if n_rows is None:
            if self._gd_rows is None:
                raise ValueError("cannot size the graph: pass extra/post or bind_code(rows, cols)")
            g, B = self._graph_no_rows(edge_index, m.device)
        else:
            g, B = graph_from_batched(edge_index, n_rows, self._gd_rows, self._gd_cols, True, m.device)
        return self._launch(g, B, m, node_in, update_args)

    def _graph_no_rows(self, edge_index, device):
        # CGNNI's check phase passes post=None and a stale `size`: recover B from a bound code.
        V, Cn = self._gd_rows, self._gd_cols
        head = edge_index[:, :1 << 16].detach().to("cpu", torch.int64)
        N = V + Cn
        later = (head[0] >= N).nonzero()
        E = int(later[0]) if later.numel() else edge_index.size(1)
        B = edge_index.size(1) // E
        return graph_from_batched(edge_index, B * N, V, Cn, True, device)

    def _launch(self, g, B, m, node_in, update_args):
        phase = _cabi.PHASE_CHK if self.flow == 'target_to_source' else _cabi.PHASE_VAR
        fuse = self._gd_update_is_builtin()
        model = _cabi.GdModel(self._gd_program, self._gd_hidden(), 0, 0)
        lib = _cabi.lib()
        F = 1 if fuse else lib.gd_propagate_features(self._gd_program, phase)
        dev = m.device
        m32 = m.detach().reshape(B, g.E).to(torch.float32).contiguous()
        x32 = None
        if node_in is not None:
            _require_cuda(node_in, "extra/post tensor")
            if node_in.numel() != B * g.N:
                raise ValueError("extra/post must have B*(rows+cols)=%d rows, got %d" % (B * g.N, node_in.numel()))
            x32 = node_in.detach().reshape(B, g.N).to(torch.float32).contiguous()
        w = self._gd_packed.get(self._gd_update_params(), dev) if fuse and self._gd_update_params() else None
        out = torch.empty((B * g.E, F), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _cabi.check(lib.gd_propagate_fwd(g.handle, C.byref(model), phase, 1 if fuse else 0, _ptr(m32), _ptr(x32),
                                             _ptr(w), _ptr(out), B, _stream(dev)), "gd_propagate_fwd")
        out = out.to(m.dtype)
        if fuse:
            return out
        return self.update(out, *update_args)


class DecoderBase(torch.nn.Module):
    """Shared machinery of the drop-in `GNNI(Nc)` decoders: the fused persistent kernel."""
    _gd_program = None

    def __init__(self, Nc, rows=None, cols=None):
        super(DecoderBase, self).__init__()
        self.Nc = Nc
        self._gd_rows = None if rows is None else int(rows)
        self._gd_cols = None if cols is None else int(cols)
        self._gd_packed = _PackedWeights()
        self._gd_graph = None

    # -- configuration ---------------------------------------------------------------------------
    def bind_code(self, rows, cols):
        """Give the Tanner-graph sizes the reference reads from module globals (rows, cols)."""
        self._gd_rows, self._gd_cols = int(rows), int(cols)
        for mod in self.children():
            if isinstance(mod, MessagePassingBase):
                mod.bind_code(rows, cols)
        return self

    def bind_graph(self, graph):
        """Attach a TannerGraph for the tensor-level API (`decode`, `decode_host`)."""
        assert isinstance(graph, TannerGraph)
        self._gd_graph = graph
        return self.bind_code(graph.V, graph.C)

    def _gd_hidden(self):
        return 0

    def _gd_params(self):
        return []

    def gd_model(self):
        return _cabi.GdModel(self._gd_program, self._gd_hidden(), int(self.Nc), 0)

    def packed_weights(self, device):
        params = self._gd_params()
        return self._gd_packed.get(params, device) if params else None

    # -- the reference's entry point ---------------------------------------------------------------
    def forward(self, data):
        """`GNNI.forward(data)`: data.x [B*(V+C), 1], data.edge_index [2, B*E] (PyG-batched, check
        ids NOT yet offset by rows) -> P(flip) [B*V, 1] in data.x's dtype."""
        x, edge_index = data.x, data.edge_index
        _require_cuda(x, "data.x")
        _require_cuda(edge_index, "data.edge_index")
        g, B = graph_from_batched(edge_index, x.size(0), self._gd_rows, self._gd_cols, False, x.device)
        prob = self.decode(x.reshape(B, g.N), graph=g)
        return prob.reshape(B * g.V, 1).to(x.dtype)

    # -- tensor-level API -----------------------------------------------------------------------------
    def decode(self, x, graph=None, return_logits=False, return_hard=False):
        """x [B, V+C] CUDA tensor -> prob [B, V] fp32 (optionally also logits and uint8 hard bits)."""
        g = graph or self._gd_graph
        if g is None:
            raise ValueError("no Tanner graph bound: call bind_graph(graph) or pass graph=")
        _require_cuda(x, "x")
        if x.dim() != 2 or x.size(1) != g.N:
            raise ValueError("x must be [B, V+C=%d], got %s" % (g.N, tuple(x.shape)))
        B = x.size(0)
        dev = x.device
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self._gd_params())
        if needs_grad and self.training:
            from .autograd import decode_with_grad
            return decode_with_grad(self, g, x, return_logits, return_hard)
        x32 = x.detach().to(torch.float32).contiguous()
        if x32.data_ptr() % 16:                # e.g. a row slice of a larger batch: the bulk copies need 16-byte alignment
            x32 = x32.clone()
        prob = torch.empty((B, g.V), dtype=torch.float32, device=dev)
        logit = torch.empty((B, g.V), dtype=torch.float32, device=dev) if return_logits else None
        hard = torch.empty((B, g.V), dtype=torch.uint8, device=dev) if return_hard else None
        model = self.gd_model()
        w = self.packed_weights(dev)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().gd_decode_fwd(g.handle, C.byref(model), _ptr(w), _ptr(x32), _ptr(prob),
                                                  _ptr(logit), _ptr(hard), B, _stream(dev)), "gd_decode_fwd")
        if return_logits or return_hard:
            return tuple(t for t in (prob, logit, hard) if t is not None)
        return prob

    def decode_packed(self, prior, synd_bits, graph=None, return_prob=False):
        """Packed form of decode() (gd_decode_packed_fwd): prior [B] fp32 CUDA = the one prior LLR every variable of the
        syndrome carries (gen_syn, quantum/error_generate.py:258), synd_bits [B, ceil(C/32)] int32 CUDA with bit c set where
        check c fired (input -1).  Returns hard_bits [B, ceil(V/32)] int32 (bit v = prob > 0.5), and prob [B, V] if asked."""
        g = graph or self._gd_graph
        if g is None:
            raise ValueError("no Tanner graph bound: call bind_graph(graph) or pass graph=")
        _require_cuda(prior, "prior")
        _require_cuda(synd_bits, "synd_bits")
        B, nw, vw = prior.numel(), (g.C + 31) // 32, (g.V + 31) // 32
        if prior.dtype != torch.float32 or synd_bits.dtype != torch.int32 or tuple(synd_bits.shape) != (B, nw):
            raise ValueError("prior must be fp32 [B], synd_bits int32 [B, %d]" % nw)
        prior, synd_bits = prior.contiguous(), synd_bits.contiguous()
        dev = prior.device
        bits = torch.empty((B, vw), dtype=torch.int32, device=dev)
        prob = torch.empty((B, g.V), dtype=torch.float32, device=dev) if return_prob else None
        model = self.gd_model()
        w = self.packed_weights(dev)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().gd_decode_packed_fwd(g.handle, C.byref(model), _ptr(w), _ptr(prior), _ptr(synd_bits), _ptr(prob),
                                                         _ptr(bits), B, _stream(dev)), "gd_decode_packed_fwd")
        return (bits, prob) if return_prob else bits

    def autotune(self, x, graph=None, max_candidates=0):
        """One-time geometry autotuning for this (decoder, graph, batch size): times the planner's best candidates on x and
        remembers the fastest for later decode() / decode_host() calls with the same batch size.  Returns the launch info."""
        g = graph or self._gd_graph
        if g is None:
            raise ValueError("no Tanner graph bound: call bind_graph(graph) or pass graph=")
        _require_cuda(x, "x")
        x32 = x.detach().to(torch.float32).contiguous()
        if x32.data_ptr() % 16:
            x32 = x32.clone()
        info = _cabi.GdLaunchInfo()
        model = self.gd_model()
        w = self.packed_weights(x.device)
        B = x32.size(0)
        n_chunks = min(4, max(1, B // 8192))                     # the chunking of gd_decode_host (csrc/gd_host.cu)
        per = ((B + n_chunks - 1) // n_chunks + 7) // 8 * 8
        with torch.cuda.device(x.device):
            for nb in sorted({B, min(per, B), B - (n_chunks - 1) * per if n_chunks > 1 else B}, reverse=True):
                if nb <= 0:
                    continue
                out = info if nb == B else _cabi.GdLaunchInfo()
                _cabi.check(_cabi.lib().gd_decode_autotune(g.handle, C.byref(model), _ptr(w), _ptr(x32), nb, int(max_candidates),
                                                           _stream(x.device), C.byref(out)), "gd_decode_autotune")
        return {k: getattr(info, k) for k, _ in _cabi.GdLaunchInfo._fields_}

    def decode_host(self, x_host, prob_out=None, hard_out=None, graph=None):
        """End-to-end call with HOST tensors (pinned for full copy bandwidth): x_host [B, V+C] fp32
        -> prob [B, V] fp32 and/or hard [B, V] uint8 on the host.  Copies are inside the call."""
        g = graph or self._gd_graph
        if g is None:
            raise ValueError("no Tanner graph bound: call bind_graph(graph) or pass graph=")
        if x_host.is_cuda or x_host.dtype != torch.float32 or not x_host.is_contiguous():
            raise ValueError("x_host must be a contiguous fp32 CPU tensor")
        B = x_host.size(0)
        if prob_out is None and hard_out is None:
            prob_out = torch.empty((B, g.V), dtype=torch.float32, pin_memory=True)
        params = self._gd_params()
        w_host = None
        if params:
            key = tuple((p.data_ptr(), p._version) for p in params)
            if getattr(self, "_gd_w_host_key", None) != key:
                with torch.no_grad():
                    self._gd_w_host = torch.cat([p.detach().reshape(-1).to("cpu", torch.float32) for p in params])
                self._gd_w_host_key = key
            w_host = self._gd_w_host
        model = self.gd_model()
        with torch.cuda.device(g.device):
            _cabi.check(_cabi.lib().gd_decode_host(g.handle, C.byref(model), _ptr(w_host), _ptr(x_host),
                                                   _ptr(prob_out), _ptr(hard_out), B), "gd_decode_host")
        if prob_out is not None and hard_out is not None:
            return prob_out, hard_out
        return prob_out if prob_out is not None else hard_out
