"""Drop-in for the model classes of the reference's quantum/decoder_v2_4_1.py: decoder_v2_4 with 2*Nc UN-TIED GraphConv
layers (each source_to_target layer owns a 1 -> 128 -> 1 Softplus `mlp1` and per-edge-type tables `W`, `W_p`; each
target_to_source layer a 1 -> 128 -> 1 `mlp`), one-hot edge types, and a gated residual
`m = chk * sigmoid(alpha) + m_p * sigmoid(beta)`.  Reference: GraphConv decoder_v2_4_1.py:255-288, GNNI :291-349.
Same class names, signatures and state_dict keys (44 tensors for Nc = 3).

How it runs: `GNNI.forward` / `decode()` is the fused persistent kernel (GD_PROG_V2_4_1: the layer's two MLPs and type tables are
re-staged in shared memory at the start of every iteration, edge types are derived from the graph inside the kernel).  The per-layer
classes keep the reference's plugin API: the propagate() of every layer -- tanh pre, "sum over siblings minus self", cat with
extra[edge_index[j]] (decoder_v2_4_1.py:134-146, identical to decoder_v2_4's) -- is the CUDA propagate kernel (C ABI
gd_propagate_fwd) and the layer's `update()` (the un-tied MLP and type tables) runs in torch on its result (`forward_layers`).

The reference derives the edge types from H at import time (`feat_onehot`, :211-236: position of the edge among its
check's four edges, + 4 for the second half of the checks); here they are computed from edge_index on first use.  Every
check must have exactly four edges (the toric code), as in the reference."""
import torch

from .. import _cabi
from ..graph import graph_from_batched
from ..message_passing import DecoderBase, MessagePassingBase, pack_mlp, _require_cuda

nb_digits = 8


def init_weights(m):
    if type(m) == torch.nn.Linear:
        torch.nn.init.kaiming_normal_(m.weight, a=0, mode='fan_in')
        m.bias.data.fill_(0)


init_weights_2 = init_weights


def _mlp():
    return torch.nn.Sequential(torch.nn.Linear(1, 128).double(), torch.nn.Softplus(), torch.nn.Linear(128, 1).double())


def edge_types(edge_index, C):
    """[E] long: position of each edge among its check's edges in ascending variable order, + 4 for checks >= C / 2
    (decoder_v2_4_1.py:211-236)."""
    ei = edge_index.detach().to("cpu", torch.int64)
    var, chk = ei[0], ei[1]
    t = torch.zeros(ei.size(1), dtype=torch.long)
    for c in range(C):
        es = (chk == c).nonzero().reshape(-1)
        if es.numel() != 4:
            raise ValueError("decoder_v2_4_1 needs every check to have 4 edges (check %d has %d)" % (c, es.numel()))
        order = es[torch.argsort(var[es], stable=True)]
        t[order] = torch.arange(4) + (4 if c > C / 2 - 1 else 0)
    return t


class MessagePassing(MessagePassingBase):
    _gd_program = _cabi.PROG_V2_4            # the propagate() of this script IS decoder_v2_4's

    def propagate(self, edge_index, extra=None, size=None, **kwargs):
        return self._propagate(edge_index, extra, size, kwargs)


class GraphConv(MessagePassing):
    def __init__(self, flow, aggr='add', bias=True):
        super(GraphConv, self).__init__(aggr, flow)
        self.flow = flow
        if self.flow == 'source_to_target':
            self.mlp1 = _mlp()
            self.mlp1.apply(init_weights)
        else:
            self.mlp = _mlp()
            self.mlp.apply(init_weights)
        self.W = torch.nn.Parameter(torch.ones((nb_digits, 1)).double())
        self.W_p = torch.nn.Parameter(torch.ones((nb_digits, 1)).double())
        self._types = None                   # [B*E] long on the device, set by GNNI.forward (the reference's global feat_onehot)

    def forward(self, m, edge_index, x, prev=None):
        x = x if x.dim() == 2 else x.unsqueeze(-1)
        if self.flow == 'source_to_target':
            m = m.mul(self.W[self._types])                       # == matmul(m.mul(feat_onehot), W)
        return self.propagate(edge_index=edge_index, size=(x.size(0), x.size(0)), x=m, extra=x)

    def update(self, aggr_out):
        if self.flow == 'source_to_target':
            prior = aggr_out[:, 1].clone().unsqueeze(1).mul(self.W_p[self._types])
            return self.mlp1(aggr_out[:, 0].clone().unsqueeze(1)) + prior
        return self.mlp(aggr_out[:, 0].clone().unsqueeze(1)).mul(aggr_out[:, 1].clone().unsqueeze(1))

    def _gd_hidden(self):
        return 128


class GNNI(DecoderBase):
    _gd_program = _cabi.PROG_V2_4_1

    def __init__(self, Nc, *, rows=None, cols=None):
        super(GNNI, self).__init__(Nc, rows, cols)
        self.layers = self._make_layer()
        self.mlp = _mlp()
        self.mlp.apply(init_weights_2)
        self.W = torch.nn.Parameter(torch.ones((nb_digits, 1)).double())
        self.W_p = torch.nn.Parameter((torch.ones((nb_digits, 1)) * 0.5).double())
        self.alpha = torch.nn.Parameter(torch.Tensor([[4]]).double())
        self.beta = torch.nn.Parameter(torch.Tensor([[-4]]).double())
        self._types_cache = {}
        if rows is not None:
            self.bind_code(rows, cols)

    def _make_layer(self):
        layers = []
        for _ in range(self.Nc):
            layers.append(GraphConv("source_to_target"))
            layers.append(GraphConv("target_to_source"))
        return torch.nn.Sequential(*layers)

    def bind_code(self, rows, cols):
        self._gd_rows, self._gd_cols = int(rows), int(cols)
        for layer in self.layers:
            layer.bind_code(rows, cols)
        return self

    def _gd_hidden(self):
        return self.mlp[0].out_features

    def _gd_params(self):
        ps = []
        for i in range(0, len(self.layers), 2):
            ps += pack_mlp(self.layers[i].mlp1) + [self.layers[i].W, self.layers[i].W_p] + pack_mlp(self.layers[i + 1].mlp)
        return ps + pack_mlp(self.mlp) + [self.W, self.W_p, self.alpha, self.beta]

    def forward_layers(self, data):
        """The reference's forward spelled out on the per-layer classes (CUDA propagate kernel + torch update() per layer):
        data.x [B*(V+C), 1], data.edge_index [2, B*E] (check ids not yet offset) -> P(flip) [B*V, 1]."""
        x, ei0 = data.x, data.edge_index
        _require_cuda(x, "data.x")
        _require_cuda(ei0, "data.edge_index")
        g, B = graph_from_batched(ei0, x.size(0), self._gd_rows, self._gd_cols, False, x.device)
        rows, N = g.V, g.N
        if self._gd_rows is None:
            self.bind_code(g.V, g.C)
        key = (id(g), B, str(x.device))
        types = self._types_cache.get(key)
        if types is None:
            types = edge_types(g.edge_index, g.C).repeat(B).to(x.device)
            self._types_cache = {key: types}
        for layer in self.layers:
            layer._types = types
        edge_index = torch.cat([ei0[0].unsqueeze(0), ei0[1].unsqueeze(0).add(rows)], dim=0)
        m = torch.zeros((edge_index.size(1), 1), dtype=torch.float64, device=x.device)
        idx = (torch.arange(B, device=x.device).unsqueeze(1) * N + torch.arange(rows, device=x.device).unsqueeze(0)).reshape(-1)
        for i in range(0, len(self.layers), 2):
            m_p = m.clone()
            m = self.layers[i](m, edge_index, x)
            m = torch.matmul(self.layers[i + 1](m, edge_index, x), torch.sigmoid(self.alpha)) + \
                torch.matmul(m_p, torch.sigmoid(self.beta))
        m = self.mlp(m).mul(self.W[types])
        prior = x[edge_index[0]].mul(self.W_p[types])
        zeros = torch.zeros((x.size(0), 1), dtype=m.dtype, device=m.device)
        res = zeros.clone().index_add_(0, edge_index[0], m)[idx] + zeros.index_add_(0, edge_index[0], prior)[idx]
        return torch.sigmoid(-1 * res)
