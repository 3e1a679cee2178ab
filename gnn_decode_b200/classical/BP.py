"""Drop-in for the model classes of the reference's classical/BP.py (parameter-free sum-product BP,
fp32): MessagePassing BP.py:33-131, GatedGraphConv :215-229, GNNI :231-259."""
from .. import _cabi
from ..message_passing import DecoderBase, MessagePassingBase


class MessagePassing(MessagePassingBase):
    _gd_program = _cabi.PROG_BP_CLASSICAL

    def propagate(self, edge_index, extra=None, size=None, **kwargs):
        return self._propagate(edge_index, extra, size, kwargs)

    _gd_builtin_update = MessagePassingBase.update


class GatedGraphConv(MessagePassing):
    def __init__(self, flow, aggr='add', bias=True):
        super(GatedGraphConv, self).__init__(aggr, flow)
        self.flow = flow

    def forward(self, m, edge_index, x=None):
        size = None
        if x is not None:
            x = x if x.dim() == 2 else x.unsqueeze(-1)
            size = (x.size(0), x.size(0))
        return self.propagate(edge_index=edge_index, size=size, x=m, extra=x)


class GNNI(DecoderBase):
    _gd_program = _cabi.PROG_BP_CLASSICAL

    def __init__(self, Nc, *, rows=None, cols=None):
        super(GNNI, self).__init__(Nc, rows, cols)
        self.ggc1 = GatedGraphConv("source_to_target")
        self.ggc2 = GatedGraphConv("target_to_source")
        if rows is not None:
            self.bind_code(rows, cols)
