"""Drop-in for the model classes of the reference's classical/CGNNI.py (learned BP for BCH(63,45) /
the toy 4x8 LDPC over AWGN, fp32, hidden-10 ReLU): MessagePassing CGNNI.py:33-122,
GatedGraphConv :212-242, GNNI :248-284.  state_dict keys match, including the parameters the
reference creates but never uses in forward (ggc*.mlp1, ggc*.rnn, ggc1.mlp2), so
classical/model/decoder_parameters_epoch{6..54}.pkl load unchanged."""
import torch

from .. import _cabi
from ..message_passing import DecoderBase, MessagePassingBase, pack_mlp


def _mlp(hidden=10):
    return torch.nn.Sequential(torch.nn.Linear(1, hidden), torch.nn.ReLU(), torch.nn.Linear(hidden, 1))


class MessagePassing(MessagePassingBase):
    """propagate(edge_index, post, size=None, **kwargs) of CGNNI.py:52-112: optional tanh(m/2),
    sum-minus-self, `+ post[edge_index[j]]` when post is given, then self.update."""
    _gd_program = _cabi.PROG_CGNNI

    def propagate(self, edge_index, post, size=None, **kwargs):
        return self._propagate(edge_index, post, size, kwargs)


class GatedGraphConv(MessagePassing):
    def __init__(self, flow, aggr='add', bias=True):
        super(GatedGraphConv, self).__init__(aggr, flow)
        self.flow = flow
        self.mlp1 = _mlp()
        self.mlp2 = _mlp()
        self.rnn = torch.nn.GRUCell(1, 1, bias=bias)     # unused by forward, kept for the checkpoints

    def forward(self, m, edge_index, x=None):
        size = None
        if x is not None:
            x = x if x.dim() == 2 else x.unsqueeze(-1)
            size = (x.size(0), x.size(0))
        return self.propagate(edge_index=edge_index, size=size, x=m, post=x)

    def update(self, aggr_out):
        if self.flow == 'target_to_source':
            return self.mlp2(aggr_out)
        return aggr_out

    _gd_builtin_update = update

    def _gd_hidden(self):
        return self.mlp2[0].out_features

    def _gd_update_params(self):
        return pack_mlp(self.mlp2) if self.flow == 'target_to_source' else []


class GNNI(DecoderBase):
    _gd_program = _cabi.PROG_CGNNI

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
        return pack_mlp(self.ggc2.mlp2) + pack_mlp(self.mlp)
