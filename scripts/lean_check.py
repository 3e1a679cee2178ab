"""Developer check of the check-owner table kernel (csrc/gd_lean.cu) on one GPU: agreement with the edge-owner kernel
(GD_NO_LEAN) and with the fp64 oracle on sampled syndromes, timing of both, and a geometry sweep through gd_set_option.

    python scripts/lean_check.py [--sweep] [--codes rot5,toric5,rot7,rot11]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnn_decode_b200 import codes, options  # noqa: E402
from gnn_decode_b200.graph import TannerGraph  # noqa: E402
from gnn_decode_b200.quantum import decoder_v2_4  # noqa: E402
from gnn_decode_b200.sampler import sample_syndromes  # noqa: E402
from oracle import restate  # noqa: E402  (developer script: the oracle is the checker here)

P10 = [0.01, 0.02, 0.03, 0.04, 0.05, 0.06, 0.07, 0.08, 0.09, 0.1]
CODES = {"rot5": (lambda: codes.rotated_surface_pcm(5), 1, 65536), "toric5": (lambda: codes.toric_pcm(5), 0, 65536),
         "rot7": (lambda: codes.rotated_surface_pcm(7), 1, 65536), "toric7": (lambda: codes.toric_pcm(7), 0, 32768),
         "rot11": (lambda: codes.rotated_surface_pcm(11), 1, 65536), "toric11": (lambda: codes.toric_pcm(11), 0, 65536)}


def weights():
    z = np.load(os.path.join(ROOT, "tests", "golden", "v2_4_toricL5_epoch3.npz"))
    return {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--codes", default="rot5,toric5,rot7,rot11")
    ap.add_argument("--parts", type=int, default=-1, help="GD_LEAN_PARTS: 1 = decode connected components apart, 0 = never")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "lean_check.json"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    w = weights()
    if args.parts >= 0:
        options.set_option("GD_LEAN_PARTS", args.parts)
    res = {}
    for name in args.codes.split(","):
        mk, noise, B = CODES[name]
        pcm = mk()
        g = TannerGraph.from_pcm(pcm, dev)
        dec = decoder_v2_4.GNNI(15)
        dec.load_state_dict(w)
        dec = dec.to(dev).eval()
        dec.bind_graph(g)
        x, _ = sample_syndromes(g, B, P10 if "5" in name else P10[:5], noise=noise, seed=1234)
        model = dec.gd_model()
        info = g.launch_info(model, B)
        prob, logit, hard = dec.decode(x, return_logits=True, return_hard=True)
        torch.cuda.synchronize()
        prob2 = dec.decode(x)
        rec = {"launch": info, "deterministic": bool(torch.equal(prob, prob2))}
        with options.option("GD_NO_LEAN"):
            info_old = g.launch_info(model, B)
            p_old, l_old, h_old = dec.decode(x, return_logits=True, return_hard=True)
            t_old = timed(lambda: dec.decode(x))
        t_new = timed(lambda: dec.decode(x))
        d = (logit - l_old).abs()
        decided = l_old.abs() > 1e-3
        rec.update({"launch_old": info_old, "ms_lean": t_new, "ms_old": t_old, "Msyn_s_lean": B / t_new / 1e3, "Msyn_s_old": B / t_old / 1e3,
                    "max_abs_dlogit_vs_old": d.max().item(), "hard_mismatch_vs_old": int((hard[decided] != h_old[decided]).sum().item()),
                    "rms_logit": l_old.pow(2).mean().sqrt().item()})
        # oracle on a subset
        idx = torch.from_numpy(np.random.RandomState(1).choice(B, 64, replace=False))
        ei = torch.from_numpy(codes.edge_index_of(pcm))
        ref = restate.decode("v2_4", ei, g.V, g.C, x[idx.to(dev)].double().cpu(), w, T=15)["logit"]
        for tag, l in (("lean", logit), ("old", l_old)):
            e = (l[idx.to(dev)].double().cpu() - ref).abs()
            rec["max_abs_vs_oracle_" + tag] = e.max().item()
            rec["max_rel_vs_oracle_" + tag] = (e / ref.abs().clamp_min(0.1)).max().item()
        print(name, json.dumps(rec), flush=True)
        res[name] = rec
        if args.sweep:
            sw = []
            for R in (2, 3, 4, 5, 6, 8, 10, 12, 15, 16, 20, 24, 28, 30, 32):
                for G in (1, 2, 3, 4, 5, 6, 8):
                    if R * G > 32:
                        continue
                    for K in (12,):
                        options.set_option("GD_LEAN_R", R); options.set_option("GD_LEAN_G", G)
                        li = g.launch_info(model, B)
                        if li["threads"] != 32 * R * G:          # geometry refused (does not fit): the fallback kernel answered
                            continue
                        try:
                            t = timed(lambda: dec.decode(x), reps=5, warm=2)
                        except Exception as ex:  # noqa: BLE001
                            print("   R=%d G=%d failed: %s" % (R, G, ex), flush=True)
                            continue
                        sw.append({"R": R, "G": G, "K": K, "ms": t, "Msyn_s": B / t / 1e3, "smem": li["smem_bytes"]})
                        print("   R=%2d G=%d K=%d  %.3f ms  %.1f M/s" % (R, G, K, t, B / t / 1e3), flush=True)
            for o in ("GD_LEAN_R", "GD_LEAN_G"):
                options.unset_option(o)
            rec["sweep"] = sw
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
