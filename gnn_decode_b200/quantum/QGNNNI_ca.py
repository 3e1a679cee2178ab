"""Drop-in for the model classes of the reference's quantum/QGNNNI_ca.py (fp32; every phase is
propagate -> 1->20->1 ReLU MLP -> GRUCell(1,1) of (previous message, new message); the check phase
multiplies by the syndrome sign instead of taking tanh; one prediction per iteration = deep
supervision): MessagePassing QGNNNI_ca.py:33-122, GatedGraphConv :177-209, GNNI :211-251.
Same class names, signatures and state_dict keys.  `GNNI.forward` returns the reference's LIST of
Nc predictions; `decode()` returns the last one, `decode_all()` the [Nc, B, V] stack."""
import ctypes as C

import torch

from .. import _cabi
from ..message_passing import DecoderBase, MessagePassingBase, pack_mlp, _ptr, _require_cuda, _stream
from ..graph import graph_from_batched


def _mlp(hidden=20):
    return torch.nn.Sequential(torch.nn.Linear(1, hidden), torch.nn.ReLU(), torch.nn.Linear(hidden, 1))


def pack_gru(cell):
    return [cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh]


class MessagePassing(MessagePassingBase):
    """propagate() of QGNNNI_ca.py:52-112: sum-minus-self in both flows, + extra (source_to_target) or
    * extra (target_to_source), then self.update."""
    _gd_program = _cabi.PROG_GRU_CA

    def propagate(self, edge_index, extra=None, size=None, **kwargs):
        return self._propagate(edge_index, extra, size, kwargs)


class GatedGraphConv(MessagePassing):
    def __init__(self, flow, aggr='add', bias=True):
        super(GatedGraphConv, self).__init__(aggr, flow)
        self.flow = flow
        self.mlp1 = _mlp()
        self.mlp2 = _mlp()
        self.rnn = torch.nn.GRUCell(1, 1, bias=bias)

    def forward(self, m, edge_index, x, results):
        x = x if x.dim() == 2 else x.unsqueeze(-1)
        mes = self.propagate(edge_index=edge_index, size=(x.size(0), x.size(0)), x=m, extra=x)
        m = self.rnn(m, mes)
        if self.flow == 'target_to_source':
            results.append(torch.zeros((x.size(0), 1), dtype=m.dtype, device=m.device).index_add_(0, edge_index[0], m))
        return m, results

    def update(self, x_j):
        return self.mlp2(x_j) if self.flow == 'target_to_source' else self.mlp1(x_j)

    _gd_builtin_update = update

    def _gd_hidden(self):
        return self.mlp1[0].out_features

    def _gd_update_params(self):
        return pack_mlp(self.mlp2 if self.flow == 'target_to_source' else self.mlp1)


class GNNI(DecoderBase):
    _gd_program = _cabi.PROG_GRU_CA

    def __init__(self, Nc, *, rows=None, cols=None):
        super(GNNI, self).__init__(Nc, rows, cols)
        self.ggc1 = GatedGraphConv("source_to_target")
        self.ggc2 = GatedGraphConv("target_to_source")
        self.mlp = _mlp()
        if rows is not None:
            self.bind_code(rows, cols)

    def _gd_hidden(self):
        return self.mlp[0].out_features

    def _gd_params(self):
        return (pack_mlp(self.ggc1.mlp1) + pack_gru(self.ggc1.rnn) + pack_mlp(self.ggc2.mlp2) + pack_gru(self.ggc2.rnn) +
                pack_mlp(self.mlp))

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
        """`GNNI.forward(data)` of QGNNNI_ca.py:222-251: a list of Nc tensors [B*V, 1]."""
        x, edge_index = data.x, data.edge_index
        _require_cuda(x, "data.x")
        _require_cuda(edge_index, "data.edge_index")
        g, B = graph_from_batched(edge_index, x.size(0), self._gd_rows, self._gd_cols, False, x.device)
        prob = self.decode_all(x.reshape(B, g.N), graph=g)
        return [prob[i].reshape(B * g.V, 1).to(x.dtype) for i in range(self.Nc)]
