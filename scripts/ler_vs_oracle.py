"""Logical error rate of the GPU decoder against the ORACLE (the reference's fp64 arithmetic, oracle/restate.py) on the same
Philox syndromes of the headline workload (decoder_v2_4, rotated surface code d = 5, depolarizing, reference checkpoint):
hard decisions compared bit by bit, failure counters (neural_BP.py:338-348 semantics) computed for both.

    python scripts/ler_vs_oracle.py [n_samples=262144] [d=5] > profiles/r02_ler_vs_oracle.txt

Developer script / test helper: the oracle is the checker here, never the thing measured."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnn_decode_b200 import codes  # noqa: E402
from gnn_decode_b200.evaluate import count_failures  # noqa: E402
from gnn_decode_b200.graph import TannerGraph  # noqa: E402
from gnn_decode_b200.quantum import decoder_v2_4  # noqa: E402
from gnn_decode_b200.sampler import sample_syndromes  # noqa: E402
from oracle import restate  # noqa: E402

P10 = [0.01, 0.02, 0.03, 0.04, 0.05, 0.06, 0.07, 0.08, 0.09, 0.1]


def run(n, d=5, seed=2025, chunk=4096, log=None):
    dev = torch.device("cuda", 0)
    Hz, Hx = codes.rotated_surface_checks(d)
    pcm = codes.css_pcm(Hz, Hx)
    logical = codes.css_logicals(Hz, Hx)
    g = TannerGraph.from_pcm(pcm, dev)
    z = np.load(os.path.join(ROOT, "tests", "golden", "v2_4_toricL5_epoch3.npz"))
    w = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}
    dec = decoder_v2_4.GNNI(15)
    dec.load_state_dict(w)
    dec = dec.to(dev).eval().bind_graph(g)
    x, err = sample_syndromes(g, n, P10 if d <= 7 else P10[:5], noise=1, seed=seed)
    t0 = time.time()
    prob, logit, hard = dec.decode(x, return_logits=True, return_hard=True)
    torch.cuda.synchronize()
    t_gpu = time.time() - t0
    ei = torch.from_numpy(codes.edge_index_of(pcm))
    torch.set_num_threads(os.cpu_count() or 1)
    hard_ref = torch.empty((n, g.V), dtype=torch.uint8)
    logit_ref = torch.empty((n, g.V), dtype=torch.float64)
    xc = x.cpu().double()
    t0 = time.time()
    for lo in range(0, n, chunk):
        r = restate.decode("v2_4", ei, g.V, g.C, xc[lo:lo + chunk], w, T=15)
        logit_ref[lo:lo + chunk] = r["logit"]
        hard_ref[lo:lo + chunk] = (r["prob"] > 0.5).to(torch.uint8)
    t_cpu = time.time() - t0
    cnt_gpu = count_failures(g, err, hard, logical).tolist()
    cnt_ref = count_failures(g, err, hard_ref.to(dev), logical).tolist()
    diff = hard.cpu() != hard_ref
    dl = (logit.double().cpu() - logit_ref).abs()
    big = logit_ref.abs() > 0.1
    out = {"n": n, "bits": hard.numel(), "hard_mismatches": int(diff.sum()),
           "max_abs_ref_logit_at_mismatch": float(logit_ref.abs()[diff].max()) if bool(diff.any()) else 0.0,
           "max_abs_dlogit": float(dl.max()), "max_rel_dlogit_above_0.1": float((dl[big] / logit_ref.abs()[big]).max()),
           "worst_over_bar": float((dl / (1e-4 * logit_ref.abs().clamp_min(1.0))).max()),
           "failures_gpu": cnt_gpu, "failures_oracle": cnt_ref, "t_gpu_s": t_gpu, "t_cpu_s": t_cpu}
    if log:
        print("decoder_v2_4, rotated surface code d = %d, depolarizing p in %s, T = 15, checkpoint quantum/new_model epoch3" % (d, P10 if d <= 7 else P10[:5]), file=log)
        print("%d Philox syndromes (seed %d), %d hard decisions: GPU (fp32 tables) vs oracle (fp64, oracle/restate.py)" % (n, seed, hard.numel()), file=log)
        print("  differing hard decisions: %d   (largest |reference logit| where they differ: %.3g)" %
              (out["hard_mismatches"], out["max_abs_ref_logit_at_mismatch"]), file=log)
        print("  logits: max |d| %.3g, max relative error on |logit| > 0.1: %.3g, worst |d| / (1e-4 max(|ref|, 1)): %.3f" %
              (out["max_abs_dlogit"], out["max_rel_dlogit_above_0.1"], out["worst_over_bar"]), file=log)
        print("  failures [residual syndrome, logical among syndrome-ok, total]: GPU %s   oracle %s" % (cnt_gpu, cnt_ref), file=log)
        print("  logical error rate: GPU %.6f   oracle %.6f   (binomial sigma %.6f)" %
              (cnt_gpu[2] / n, cnt_ref[2] / n, (cnt_ref[2] / n * (1 - cnt_ref[2] / n) / n) ** 0.5), file=log)
        print("  time: GPU decode %.3f s (first call, includes table build), oracle %.1f s on %d host threads" % (t_gpu, t_cpu, os.cpu_count() or 1), file=log)
    return out


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 262144, d=int(sys.argv[2]) if len(sys.argv) > 2 else 5, log=sys.stdout)
