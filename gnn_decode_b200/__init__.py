"""gnn_decode_b200 -- B200-native (sm_100a) message-passing hot path of ironmanaudi/GNN-decode.

Layout (only what the hot path needs):
  csrc/               hand-written CUDA kernels + the C ABI (include/gnn_decode.h)
  _cabi.py            ctypes binding of the C ABI (no CPU fallback: raises if the .so is missing)
  graph.py            TannerGraph: destination-sorted CSR/CSC tables on the device
  message_passing.py  host mirror of the reference's MessagePassing / GNNI plugin interface
  quantum/, classical/  drop-ins named after the reference scripts (decoder_v2_4, QGNNI, BP, CGNNI)
  codes.py            parity-check matrices (toric = reference generator; rotated surface, HGP, BCH)
  sampler.py          on-GPU Philox syndrome sampler; evaluate.py: on-GPU failure counters
  dist.py             batch sharding over ranks (inference: no collective; training: grad allreduce)
"""
from . import _cabi  # noqa: F401
from .graph import TannerGraph, graph_from_batched  # noqa: F401

__version__ = "0.1.0"
