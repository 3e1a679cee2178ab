"""Drop-in for the model classes of the reference's quantum/decoder_v2_4.py ("BP free decoder,
with phase one & two modified to mlp"): same class names, constructor / forward / propagate /
update signatures and state_dict keys (`ggc1.mlp.0.weight` ...), so the shipped checkpoints
quantum/new_model/decoder_parameters_epoch*.pkl load unchanged.

Reference: MessagePassing decoder_v2_4.py:66-158, GraphConv :230-257, GNNI :260-294.
Differences, all outside the arithmetic:
  * kernels compute in fp32 (the reference in fp64); fp64 inputs are cast once and the result is
    cast back (logits agree with the fp64 reference to ~1e-6 relative);
  * rows / cols / BATCH_SIZE are not module globals: they are inferred from edge_index and x, or
    given as keyword arguments / bind_code();
  * no import-time dataset generation or checkpoint load.
"""
import torch

from .. import _cabi
from ..message_passing import DecoderBase, MessagePassingBase, pack_mlp, special_args, __size_error_msg__  # noqa: F401


def init_weights(m):
    """kaiming-normal weights, zero bias (decoder_v2_4.py:211-215)."""
    if type(m) == torch.nn.Linear:
        torch.nn.init.kaiming_normal_(m.weight, a=0, mode='fan_in')
        m.bias.data.fill_(0)


init_weights_2 = init_weights


def _mlp(n_in, hidden=128):
    return torch.nn.Sequential(torch.nn.Linear(n_in, hidden).double(), torch.nn.Softplus(),
                               torch.nn.Linear(hidden, 1).double())


class MessagePassing(MessagePassingBase):
    """propagate() of decoder_v2_4.py:85-148: [tanh(m/2) for target_to_source], sum-minus-self over
    edge_index[j], cat with extra[edge_index[j]], then self.update."""
    _gd_program = _cabi.PROG_V2_4

    def propagate(self, edge_index, extra=None, size=None, **kwargs):
        return self._propagate(edge_index, extra, size, kwargs)


class GraphConv(MessagePassing):
    def __init__(self, flow, aggr='add', bias=True):
        super(GraphConv, self).__init__(aggr, flow)
        self.flow = flow
        self.mlp = _mlp(2 if self.flow == 'source_to_target' else 1)
        self.mlp.apply(init_weights)

    def forward(self, m, edge_index, x, prev=None):
        x = x if x.dim() == 2 else x.unsqueeze(-1)
        return self.propagate(edge_index=edge_index, size=(x.size(0), x.size(0)), x=m, extra=x)

    def update(self, aggr_out):
        # used only when a subclass calls it explicitly: the built-in update runs fused in the kernel
        if self.flow == 'source_to_target':
            return self.mlp(aggr_out)
        return self.mlp(aggr_out[:, 0].clone().unsqueeze(1)).mul(aggr_out[:, 1].clone().unsqueeze(1))

    _gd_builtin_update = update

    def _gd_hidden(self):
        return self.mlp[0].out_features

    def _gd_update_params(self):
        return pack_mlp(self.mlp)


class GNNI(DecoderBase):
    _gd_program = _cabi.PROG_V2_4

    def __init__(self, Nc, *, rows=None, cols=None):
        super(GNNI, self).__init__(Nc, rows, cols)
        self.ggc1 = GraphConv("source_to_target")
        self.ggc2 = GraphConv("target_to_source")
        self.mlp = _mlp(1)
        self.mlp.apply(init_weights_2)
        if rows is not None:
            self.bind_code(rows, cols)

    def _gd_hidden(self):
        return self.mlp[0].out_features

    def _gd_params(self):
        return pack_mlp(self.ggc1.mlp) + pack_mlp(self.ggc2.mlp) + pack_mlp(self.mlp)
