"""Stand-in for torch_geometric.utils.scatter_ (PyG <= 1.4), 'add' only -- the only
aggregation the reference instantiates (call sites classical/CGNNI.py:103,105,273;
quantum/QGNNI.py:105,107,242; quantum/BP.py:111,113,119,209)."""
import torch


def scatter_(name, src, index, dim_size=None):
    assert name == 'add', "oracle shim: only aggr='add' is ever used by the reference"
    if dim_size is None:
        dim_size = int(index.max()) + 1
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype)
    return out.index_add_(0, index, src)
