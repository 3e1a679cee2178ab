"""On-GPU synthetic syndrome sampler (Philox4x32-10), with the layout and distribution of the
reference's gen_syn (quantum/error_generate.py:252-278) -- see csrc/gd_sampler.cu."""
import ctypes as ct

import numpy as np
import torch

from . import _cabi

NOISE_IID_XZ, NOISE_DEPOLARIZING, NOISE_AWGN_ZERO, NOISE_AWGN_ONE = 0, 1, 2, 3


def sample_syndromes(graph, B, p_list, noise=NOISE_IID_XZ, seed=1234, first_sample=0, x_out=None, err_out=None):
    """Returns (x [B, V+C] fp32 = [prior | (-1)^syndrome], err [B, V] uint8) on graph.device.
    noise 2 / 3 (classical, CGNNI.py:125-159): p_list = SNR in dB, x = [channel LLR | 0], err = codeword.
    Sample s depends only on (seed, first_sample + s): shards of a batch can be drawn on
    different ranks and are bit-identical to the single-GPU draw."""
    dev = graph.device
    x = torch.empty((B, graph.N), dtype=torch.float32, device=dev) if x_out is None else x_out
    err = torch.empty((B, graph.V), dtype=torch.uint8, device=dev) if err_out is None else err_out
    p = np.ascontiguousarray(np.asarray(p_list, dtype=np.float32))
    st = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().gd_sample(graph.handle, int(noise), ct.c_void_p(p.ctypes.data), int(p.size),
                                          ct.c_uint64(seed), ct.c_uint64(first_sample), ct.c_void_p(x.data_ptr()),
                                          ct.c_void_p(err.data_ptr()), int(B), ct.c_void_p(st)), "gd_sample")
    return x, err
