"""Drop-in for the model classes of the reference's quantum/decoder_v1_2_2.py: a neural sum-product decoder with DEEP SUPERVISION.
Variable phase: 2 -> 256 -> 1 Tanh MLP of [sum over the variable's other edges, prior] (MessagePassing :120-124, GraphConv :216-233);
check phase: the sum-product rule with the syndrome sign (:105-119, eps 1e-20 / 1e-12, no input clamp) and identity update; residual
+ m_p (GNNI :266); every iteration's messages are read out through `mlp` (2 -> 256 -> 1 Tanh of [sum of the messages at the
variable, prior], :267-278), so `forward` returns a LIST of Nc predictions.  Same class names, signatures and state_dict keys
(ggc1.mlp.*, mlp.*).

`GNNI.forward` / `decode_all()` run the fused persistent kernel (GD_PROG_V1_2_2 with GD_FLAG_ALL_ITERS); the per-layer classes
keep the reference's plugin API: GraphConv.propagate() is the CUDA propagate kernel (gd_propagate_fwd: "sum minus self, cat prior"
/ the sum-product check rule) and a subclass's own update() runs in torch on its result."""
import ctypes as C

import torch

from .. import _cabi
from ..graph import graph_from_batched
from ..message_passing import DecoderBase, MessagePassingBase, pack_mlp, _ptr, _require_cuda, _stream


def init_weights(m):
    if type(m) == torch.nn.Linear:
        torch.nn.init.kaiming_normal_(m.weight, a=0, mode='fan_in')
        m.bias.data.fill_(0)


init_weights_2 = init_weights


def _mlp():
    return torch.nn.Sequential(torch.nn.Linear(2, 256).double(), torch.nn.Tanh(), torch.nn.Linear(256, 1).double())


class MessagePassing(MessagePassingBase):
    """propagate() of decoder_v1_2_2.py:52-126: source_to_target = sum-minus-self, cat extra; target_to_source = the sum-product
    check update with the syndrome sign."""
    _gd_program = _cabi.PROG_V1_2_2

    def propagate(self, edge_index, extra=None, size=None, **kwargs):
        return self._propagate(edge_index, extra, size, kwargs)


class GraphConv(MessagePassing):
    def __init__(self, flow, aggr='add', bias=True):
        super(GraphConv, self).__init__(aggr, flow)
        self.flow = flow
        if self.flow == 'source_to_target':
            self.mlp = _mlp()
            self.mlp.apply(init_weights)

    def forward(self, m, edge_index, x):
        x = x if x.dim() == 2 else x.unsqueeze(-1)
        return self.propagate(edge_index=edge_index, size=(x.size(0), x.size(0)), x=m, extra=x)

    def update(self, aggr_out):
        if self.flow == 'source_to_target':
            return self.mlp(aggr_out)
        return aggr_out

    def _gd_hidden(self):
        return 256

    def _gd_update_is_builtin(self):
        # the check phase's identity update is part of the kernel (one feature out); the variable phase's MLP runs in torch
        return self.flow == 'target_to_source' and type(self).update is GraphConv.update


class GNNI(DecoderBase):
    _gd_program = _cabi.PROG_V1_2_2

    def __init__(self, Nc, *, rows=None, cols=None):
        super(GNNI, self).__init__(Nc, rows, cols)
        self.ggc1 = GraphConv("source_to_target")
        self.ggc2 = GraphConv("target_to_source")
        self.mlp = _mlp()
        self.mlp.apply(init_weights_2)
        if rows is not None:
            self.bind_code(rows, cols)

    def _gd_hidden(self):
        return self.mlp[0].out_features

    def _gd_params(self):
        return pack_mlp(self.ggc1.mlp) + pack_mlp(self.mlp)

    def decode_all(self, x, graph=None, return_logits=False):
        """x [B, V+C] CUDA -> prob [Nc, B, V] fp32: the prediction after every iteration."""
        g = graph or self._gd_graph
        if g is None:
            raise ValueError("no Tanner graph bound: call bind_graph(graph) or pass graph=")
        _require_cuda(x, "x")
        if x.dim() != 2 or x.size(1) != g.N:
            raise ValueError("x must be [B, V+C=%d], got %s" % (g.N, tuple(x.shape)))
        B, dev = x.size(0), x.device
        x32 = x.detach().to(torch.float32).contiguous()
        if x32.data_ptr() % 16:
            x32 = x32.clone()
        prob = torch.empty((self.Nc, B, g.V), dtype=torch.float32, device=dev)
        logit = torch.empty((self.Nc, B, g.V), dtype=torch.float32, device=dev) if return_logits else None
        model = _cabi.GdModel(self._gd_program, self._gd_hidden(), int(self.Nc), _cabi.FLAG_ALL_ITERS)
        w = self.packed_weights(dev)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().gd_decode_fwd(g.handle, C.byref(model), _ptr(w), _ptr(x32), _ptr(prob), _ptr(logit),
                                                  None, B, _stream(dev)), "gd_decode_fwd")
        return (prob, logit) if return_logits else prob

    def forward(self, data):
        """`GNNI.forward(data)` of decoder_v1_2_2.py:250-278: a list of Nc tensors P(flip) [B*V, 1]."""
        x, edge_index = data.x, data.edge_index
        _require_cuda(x, "data.x")
        _require_cuda(edge_index, "data.edge_index")
        g, B = graph_from_batched(edge_index, x.size(0), self._gd_rows, self._gd_cols, False, x.device)
        prob = self.decode_all(x.reshape(B, g.N), graph=g)
        return [prob[i].reshape(B * g.V, 1).to(x.dtype) for i in range(self.Nc)]
