"""Drop-in for the model classes of the reference's quantum/BP.py (parameter-free sum-product BP
with the syndrome sign folded into the check update): MessagePassing BP.py:35-132,
GatedGraphConv :179-188, GNNI :191-219.

fp32 note: the reference clamps the check product to +-(1 - 1e-12), which is 1.0f in fp32; the
kernel evaluates the same clamped expression as log1p(p) - log(max(-expm1(ext), 1e-12)), which
keeps fp64-like accuracy in the saturated regime."""
from .. import _cabi
from ..message_passing import DecoderBase, MessagePassingBase


class MessagePassing(MessagePassingBase):
    _gd_program = _cabi.PROG_BP_QUANTUM

    def propagate(self, edge_index, extra=None, size=None, **kwargs):
        return self._propagate(edge_index, extra, size, kwargs)

    _gd_builtin_update = MessagePassingBase.update   # identity update: nothing to fuse


class GatedGraphConv(MessagePassing):
    def __init__(self, flow, aggr='add', bias=True):
        super(GatedGraphConv, self).__init__(aggr, flow)

    def forward(self, m, edge_index, x):
        x = x if x.dim() == 2 else x.unsqueeze(-1)
        return self.propagate(edge_index=edge_index, size=(x.size(0), x.size(0)), x=m, extra=x)


class GNNI(DecoderBase):
    _gd_program = _cabi.PROG_BP_QUANTUM

    def __init__(self, Nc, *, rows=None, cols=None):
        super(GNNI, self).__init__(Nc, rows, cols)
        self.ggc1 = GatedGraphConv("source_to_target")
        self.ggc2 = GatedGraphConv("target_to_source")
        if rows is not None:
            self.bind_code(rows, cols)
