"""Developer timing of the fused decoder on one GPU (not the driver's bench.py)."""
import ctypes as C
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_decode_b200 import _cabi, codes
from gnn_decode_b200.graph import TannerGraph
from gnn_decode_b200.quantum import decoder_v2_4, QGNNI, BP, neural_BP, QGNNNI_ca
from gnn_decode_b200.classical import CGNNI


def microbench():
    lib = _cabi.lib()
    out = {}
    for kind, name in enumerate(["ex2", "ex2+lg2", "ffma", "softplus_unit", "ffma2", "rcp", "ex2+rcp+lg2"]):
        r = (C.c_double * 2)()
        _cabi.check(lib.gd_microbench(kind, 4096, 0, r))
        out[name] = r[0]
        print("microbench %-14s %.3f T units/s (%.3f ms)" % (name, r[0] / 1e12, r[1]))
    return out


def time_decode(dec, g, B, iters=5, warm=2):
    dev = g.device
    x = torch.randn(B, g.N, device=dev)
    x[:, g.V:] = torch.sign(x[:, g.V:])
    x[:, :g.V] = 2.0 + 0.5 * x[:, :g.V]
    if os.environ.get("GD_AUTOTUNE"):
        print("   autotuned:", dec.autotune(x, graph=g))
    for _ in range(warm):
        dec.decode(x, graph=g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        dec.decode(x, graph=g)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    peaks = microbench()
    cases = [("v2_4 rot-d5", decoder_v2_4.GNNI(15), codes.rotated_surface_pcm(5), 65536),
             ("v2_4 toric-L5", decoder_v2_4.GNNI(15), codes.toric_pcm(5), 65536),
             ("v2_4 rot-d11", decoder_v2_4.GNNI(15), codes.rotated_surface_pcm(11), 16384),
             ("v2_4 toric-L11", decoder_v2_4.GNNI(15), codes.toric_pcm(11), 8192),
             ("qgnni toric-L5", QGNNI.GNNI(25), codes.toric_pcm(5), 65536),
             ("cgnni bch", CGNNI.GNNI(25), codes.bch_63_45_pcm(), 65536),
             ("cgnni ldpc", CGNNI.GNNI(25), codes.ldpc_toy_pcm(), 1 << 20),
             ("bp_q toric-L5", BP.GNNI(10), codes.toric_pcm(5), 65536),
             ("bp_q hgp1600", BP.GNNI(20), codes.hgp_pcm(), 16384),
             ("qgnni hgp1600", QGNNI.GNNI(20), codes.hgp_pcm(), 16384),
             ("v2_4 hgp1600", decoder_v2_4.GNNI(3), codes.hgp_pcm(), 8192),
             ("neural_bp toric-L5", neural_BP.GNNI(15, n_edges=192), codes.toric_pcm(5), 65536),
             ("gru_ca toric-L5", QGNNNI_ca.GNNI(25), codes.toric_pcm(5), 65536)]
    only = sys.argv[1:] 
    for name, dec, pcm, B in cases:
        if only and not any(o in name for o in only):
            continue
        g = TannerGraph.from_pcm(pcm, dev)
        if "v2_4" in name:      # the shipped checkpoint (quantum/new_model epoch3), as bench.py uses
            z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                                     "v2_4_toricL5_epoch3.npz"))
            dec.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")})
        dec = dec.to(dev).eval()
        info = g.launch_info(dec.gd_model(), B)
        ms = time_decode(dec, g, B)
        syn = B / (ms * 1e-3)
        line = "%-16s B=%-8d %8.3f ms  %10.3f M syn/s  %s" % (name, B, ms, syn / 1e6, info)
        if "v2_4" in name:
            T = dec.Nc
            units = g.E * (T * 256 + 128)
            line += "  softplus-units/s %.3f T (%.1f%% of microbench unit rate, %.1f%% of 2-MUFU rate)" % (
                syn * units / 1e12, 100 * syn * units / peaks["softplus_unit"], 100 * syn * units / peaks["ex2+lg2"])
        print(line, flush=True)
