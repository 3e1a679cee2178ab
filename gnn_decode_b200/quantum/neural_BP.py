"""Drop-in for the model classes of the reference's quantum/neural_BP.py (neural belief propagation:
the sum-product check update with learned per-edge multiplicative weights, 2*Nc un-tied layers, a
gated residual and a weighted read-out): MessagePassing neural_BP.py:43-139, GraphConv :236-260,
GNNI :263-314.  Same class names, signatures and state_dict keys (`layers.{i}.W`, `layers.{i}.W_p`,
`W`, `W_p`, `alpha`).

The reference sizes its parameters from the module global `H` (`int(H.sum())`, :241-242, :269-270);
here the number of edges is the keyword-only argument `n_edges`.  The weight clipping helper
(`WeightClipper`, :353-366) is host-side optimiser glue and is kept as is."""
import ctypes as C

import torch

from .. import _cabi
from ..message_passing import DecoderBase, MessagePassingBase


class MessagePassing(MessagePassingBase):
    """propagate() of neural_BP.py:68-131: target_to_source is the sum-product update (no +-10 clamp,
    eps 1e-20 / 1 - 1e-15) with the syndrome sign; source_to_target is sum-minus-self, cat prior."""
    _gd_program = _cabi.PROG_NEURAL_BP

    def propagate(self, edge_index, extra=None, size=None, **kwargs):
        return self._propagate(edge_index, extra, size, kwargs)


class GraphConv(MessagePassing):
    def __init__(self, flow, aggr='add', bias=True, *, n_edges):
        super(GraphConv, self).__init__(aggr, flow)
        self.flow = flow
        self.n_edges = int(n_edges)
        self.W = torch.nn.Parameter(torch.ones((self.n_edges, 1)).double())
        self.W_p = torch.nn.Parameter(torch.ones((self.n_edges, 1)).double())

    def forward(self, m, edge_index, x, prev=None):
        x = x if x.dim() == 2 else x.unsqueeze(-1)
        if self.flow == 'source_to_target':
            m = m.mul(self.W.repeat(m.size(0) // self.n_edges, 1))
        return self.propagate(edge_index=edge_index, size=(x.size(0), x.size(0)), x=m, extra=x)

    def update(self, aggr_out):
        if self.flow == 'source_to_target':
            w = self.W_p.repeat(aggr_out.size(0) // self.n_edges, 1)
            return aggr_out[:, 0].clone().unsqueeze(1) + aggr_out[:, 1].clone().unsqueeze(1).mul(w)
        return aggr_out

    _gd_builtin_update = update

    def _gd_hidden(self):
        return self.n_edges

    def _gd_update_params(self):
        return [self.W_p] if self.flow == 'source_to_target' else []


class GNNI(DecoderBase):
    _gd_program = _cabi.PROG_NEURAL_BP

    def __init__(self, Nc, *, n_edges, rows=None, cols=None):
        super(GNNI, self).__init__(Nc, rows, cols)
        self.n_edges = int(n_edges)
        self.layers = self._make_layer()
        self.W = torch.nn.Parameter(torch.ones((self.n_edges, 1)).double())
        self.W_p = torch.nn.Parameter((torch.ones((self.n_edges, 1)) * 0.5).double())
        self.alpha = torch.nn.Parameter(torch.Tensor([[0]]).double())
        if rows is not None:
            self.bind_code(rows, cols)

    def _make_layer(self):
        layers = []
        for _ in range(self.Nc):
            layers.append(GraphConv("source_to_target", n_edges=self.n_edges))
            layers.append(GraphConv("target_to_source", n_edges=self.n_edges))
        return torch.nn.Sequential(*layers)

    def bind_code(self, rows, cols):
        super(GNNI, self).bind_code(rows, cols)
        for layer in self.layers:
            layer.bind_code(rows, cols)
        return self

    def _gd_hidden(self):
        return self.n_edges

    def _gd_params(self):
        out = []
        for l in range(self.Nc):
            out += [self.layers[2 * l].W, self.layers[2 * l].W_p]
        return out + [self.W, self.W_p, self.alpha]


class WeightClipper(object):
    """neural_BP.py:353-366: clamp W / W_p into [1e-10, 1e10] after an optimiser step."""

    def __init__(self, frequency=5):
        self.frequency = frequency

    def __call__(self, module):
        for name in ('W', 'W_p'):
            if hasattr(module, name):
                getattr(module, name).data = getattr(module, name).data.clamp(1e-10, 1e10)
