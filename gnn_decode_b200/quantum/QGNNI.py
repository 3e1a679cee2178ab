"""Drop-in for the model classes of the reference's quantum/QGNNI.py (hidden-10 ReLU learned BP on
the toric code): MessagePassing QGNNI.py:35-124, GraphConv :186-214, GNNI :217-252.
Same names, signatures and state_dict keys; see decoder_v2_4.py in this package for the
(non-arithmetic) differences."""
import torch

from .. import _cabi
from ..message_passing import DecoderBase, MessagePassingBase, pack_mlp


def _mlp(hidden=10):
    return torch.nn.Sequential(torch.nn.Linear(1, hidden).double(), torch.nn.ReLU(),
                               torch.nn.Linear(hidden, 1).double())


class MessagePassing(MessagePassingBase):
    """propagate() of QGNNI.py:54-116: source_to_target adds extra, target_to_source does
    tanh(m/2), sum-minus-self and cats the syndrome sign."""
    _gd_program = _cabi.PROG_QGNNI

    def propagate(self, edge_index, extra=None, size=None, **kwargs):
        return self._propagate(edge_index, extra, size, kwargs)


class GraphConv(MessagePassing):
    def __init__(self, flow, aggr='add', bias=True):
        super(GraphConv, self).__init__(aggr, flow)
        self.flow = flow
        self.mlp = _mlp()          # present (and saved) for both flows; only target_to_source uses it

    def forward(self, m, edge_index, x):
        x = x if x.dim() == 2 else x.unsqueeze(-1)
        return self.propagate(edge_index=edge_index, size=(x.size(0), x.size(0)), x=m, extra=x)

    def update(self, aggr_out):
        if self.flow == 'target_to_source':
            h = self.mlp(aggr_out[:, 0].clone().unsqueeze(1))
            return h.mul(aggr_out[:, 1].clone().unsqueeze(1))
        return aggr_out

    _gd_builtin_update = update

    def _gd_hidden(self):
        return self.mlp[0].out_features

    def _gd_update_params(self):
        return pack_mlp(self.mlp) if self.flow == 'target_to_source' else []


class GNNI(DecoderBase):
    _gd_program = _cabi.PROG_QGNNI

    def __init__(self, Nc, *, rows=None, cols=None):
        super(GNNI, self).__init__(Nc, rows, cols)
        self.ggc1 = GraphConv("source_to_target")
        self.ggc2 = GraphConv("target_to_source")
        self.mlp = _mlp()
        if rows is not None:
            self.bind_code(rows, cols)

    def _gd_hidden(self):
        return self.mlp[0].out_features

    def _gd_params(self):
        return pack_mlp(self.ggc2.mlp) + pack_mlp(self.mlp)
