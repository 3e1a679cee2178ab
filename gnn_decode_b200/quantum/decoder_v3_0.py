"""Drop-in for the model classes of the reference's quantum/decoder_v3_0.py: every phase is propagate (sum over the node's other
edges, no tanh, cat the node's input: MessagePassing :103-112) -> 2 -> 10 -> 1 ReLU MLP (GraphConv.update :238-247) ->
GRUCell(1,1)(input = previous message, hidden = the update) (:229-232); two read-outs through the same `mlp` (GNNI :259-290):
res = mlp(sum of the final messages at the node) + x and res_p = mlp(sum at the check side of the messages after the LAST variable
phase), both returned over all V+C nodes.  Same class names, signatures and state_dict keys (each GraphConv owns mlp1, mlp2, rnn1
and rnn2 as in the reference; the forward uses ggc1.mlp1 / ggc1.rnn1 and ggc2.mlp2 / ggc2.rnn2 only).

`GNNI.forward` runs the fused persistent kernel (C ABI gd_decode_fwd_aux, GD_PROG_V3_0) and returns the reference's LIST
[sigmoid(-res), sigmoid(-res_p)] of [B*(V+C), 1] tensors; `decode()` returns the variable rows of the first, `decode_aux()` also
the check rows of the second.  The reference compares its loop index with the script's GLOBAL Nc (:267), so it only runs with
GNNI(Nc) built from that constant; here `self.Nc` is used, which is the same thing whenever the reference runs at all."""
import ctypes as C

import torch

from .. import _cabi
from ..graph import graph_from_batched
from ..message_passing import DecoderBase, MessagePassingBase, pack_mlp, _ptr, _require_cuda, _stream
from .QGNNNI_ca import pack_gru


def _mlp(k):
    return torch.nn.Sequential(torch.nn.Linear(k, 10).double(), torch.nn.ReLU(), torch.nn.Linear(10, 1).double())


class MessagePassing(MessagePassingBase):
    """propagate() of decoder_v3_0.py:59-116: sum-minus-self in both flows, cat extra[edge_index[j]], then self.update."""
    _gd_program = _cabi.PROG_V3_0

    def propagate(self, edge_index, extra=None, size=None, **kwargs):
        return self._propagate(edge_index, extra, size, kwargs)


class GraphConv(MessagePassing):
    def __init__(self, flow, aggr='add', bias=True):
        super(GraphConv, self).__init__(aggr, flow)
        self.flow = flow
        self.mlp1 = _mlp(2)
        self.mlp2 = _mlp(2)
        self.rnn1 = torch.nn.GRUCell(1, 1, bias=bias).double()
        self.rnn2 = torch.nn.GRUCell(1, 1, bias=bias).double()

    def forward(self, m, edge_index, x):
        x = x if x.dim() == 2 else x.unsqueeze(-1)
        mes = self.propagate(edge_index=edge_index, size=(x.size(0), x.size(0)), x=m, extra=x)
        if self.flow == 'target_to_source':
            return self.rnn2(m, mes)
        return self.rnn1(m, mes)

    def update(self, aggr_out):
        if self.flow == 'target_to_source':
            return self.mlp2(aggr_out)
        return self.mlp1(aggr_out)

    def _gd_hidden(self):
        return self.mlp1[0].out_features


class GNNI(DecoderBase):
    _gd_program = _cabi.PROG_V3_0

    def __init__(self, Nc, *, rows=None, cols=None):
        super(GNNI, self).__init__(Nc, rows, cols)
        self.ggc1 = GraphConv("source_to_target")
        self.ggc2 = GraphConv("target_to_source")
        self.mlp = _mlp(1)
        if rows is not None:
            self.bind_code(rows, cols)

    def _gd_hidden(self):
        return self.mlp[0].out_features

    def _gd_params(self):
        return (pack_mlp(self.ggc1.mlp1) + pack_gru(self.ggc1.rnn1) + pack_mlp(self.ggc2.mlp2) + pack_gru(self.ggc2.rnn2) +
                pack_mlp(self.mlp))

    def decode_aux(self, x, graph=None, return_logits=False):
        """x [B, V+C] CUDA -> (prob [B, V], prob_chk [B, C]) fp32: the variable rows of the first read-out and the check rows of
        the second; with return_logits also (logit, logit_chk)."""
        g = graph or self._gd_graph
        if g is None:
            raise ValueError("no Tanner graph bound: call bind_graph(graph) or pass graph=")
        _require_cuda(x, "x")
        if x.dim() != 2 or x.size(1) != g.N:
            raise ValueError("x must be [B, V+C=%d], got %s" % (g.N, tuple(x.shape)))
        if self.Nc < 1:
            raise ValueError("decoder_v3_0 needs Nc >= 1 (its second read-out is taken inside the last iteration)")
        B, dev = x.size(0), x.device
        x32 = x.detach().to(torch.float32).contiguous()
        if x32.data_ptr() % 16:
            x32 = x32.clone()
        f32 = dict(dtype=torch.float32, device=dev)
        prob, prob_c = torch.empty((B, g.V), **f32), torch.empty((B, g.C), **f32)
        logit = torch.empty((B, g.V), **f32) if return_logits else None
        logit_c = torch.empty((B, g.C), **f32) if return_logits else None
        model = self.gd_model()
        w = self.packed_weights(dev)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().gd_decode_fwd_aux(g.handle, C.byref(model), _ptr(w), _ptr(x32), _ptr(prob), _ptr(logit), None,
                                                      _ptr(prob_c), _ptr(logit_c), B, _stream(dev)), "gd_decode_fwd_aux")
        return (prob, prob_c, logit, logit_c) if return_logits else (prob, prob_c)

    def forward(self, data):
        """`GNNI.forward(data)` of decoder_v3_0.py:259-290: [sigmoid(-res), sigmoid(-res_p)], each [B*(V+C), 1]."""
        x, edge_index = data.x, data.edge_index
        _require_cuda(x, "data.x")
        _require_cuda(edge_index, "data.edge_index")
        g, B = graph_from_batched(edge_index, x.size(0), self._gd_rows, self._gd_cols, False, x.device)
        xb = x.reshape(B, g.N)
        prob, prob_c = self.decode_aux(xb, graph=g)
        # the rows no message reaches: mlp(0) (+ the node's input for the first read-out), decoder_v3_0.py:275-276
        with torch.no_grad():
            z0 = self.mlp(torch.zeros((1, 1), dtype=self.mlp[0].weight.dtype, device=x.device)).to(x.dtype)
        res_all = torch.cat([prob.to(x.dtype), torch.sigmoid(-1 * (z0 + xb[:, g.V:]))], 1)
        res_p_all = torch.cat([torch.sigmoid(-1 * z0).expand(B, g.V), prob_c.to(x.dtype)], 1)
        return [res_all.reshape(B * g.N, 1), res_p_all.reshape(B * g.N, 1)]
