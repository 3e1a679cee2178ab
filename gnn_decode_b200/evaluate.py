"""On-GPU hard-decision evaluation with the semantics of the reference's LossFunc.forward(train=0)
(quantum/neural_BP.py:338-348): residual-syndrome failures and logical failures."""
import ctypes as ct

import numpy as np
import torch

from . import _cabi


def count_failures(graph, err, hard, logical=None, counts=None):
    """err, hard: [B, V] uint8 CUDA tensors.  logical: [K, V] 0/1 array (as `logical` from
    H_Prep.get_logical / codes.css_logicals).  Returns a CUDA int64 tensor
    [syndrome_failures, logical_failures_among_syndrome_ok, total_failures] (accumulated into
    `counts` when given)."""
    dev = graph.device
    if counts is None:
        counts = torch.zeros(3, dtype=torch.int64, device=dev)
    K, ldev = 0, None
    if logical is not None:
        key = np.ascontiguousarray(np.asarray(logical, dtype=np.uint8))
        ldev = graph._logical_dev.get(key.tobytes())
        if ldev is None:
            ldev = torch.from_numpy(key).to(dev)
            graph._logical_dev[key.tobytes()] = ldev
        K = key.shape[0]
        if key.shape[1] != graph.V:
            raise ValueError("logical must be [K, V=%d]" % graph.V)
    if err.shape != hard.shape or err.dtype != torch.uint8 or hard.dtype != torch.uint8:
        raise ValueError("err and hard must be uint8 tensors of the same [B, V] shape")
    st = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().gd_eval_failures(graph.handle, ct.c_void_p(ldev.data_ptr()) if K else None, K,
                                                 ct.c_void_p(err.contiguous().data_ptr()),
                                                 ct.c_void_p(hard.contiguous().data_ptr()), err.size(0),
                                                 ct.c_void_p(counts.data_ptr()), ct.c_void_p(st)), "gd_eval_failures")
    return counts
