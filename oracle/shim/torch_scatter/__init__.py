"""Stand-in for torch_scatter 1.x scatter_add(src, index, dim, out, dim_size, fill_value)
(call sites quantum/decoder_v2_4.py:28,30,43).  Textbook scatter-add, sequential edge order."""
import torch


def scatter_add(src, index, dim=-1, out=None, dim_size=None, fill_value=0):
    assert dim == 0 and out is None and fill_value == 0
    if dim_size is None:
        dim_size = int(index.max()) + 1
    res = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype)
    return res.index_add_(0, index, src)


def scatter_mean(*a, **k):  # referenced by getattr() only for aggr='mean' (dead code)
    raise NotImplementedError


def scatter_max(*a, **k):
    raise NotImplementedError
