"""Developer check: the table kernel on wide message domains (the collapsed epoch-67 checkpoint, freshly initialised weights) against
the fp64 oracle on every row, next to the edge-owner kernel; and which forward/backward a training step takes (stash header)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import Golden, logit_worst  # noqa: E402
from gnn_decode_b200 import codes, options  # noqa: E402
from gnn_decode_b200.graph import TannerGraph  # noqa: E402
from gnn_decode_b200.quantum import decoder_v2_4  # noqa: E402
from gnn_decode_b200.sampler import sample_syndromes  # noqa: E402
from gnn_decode_b200.train import FusedTrainer  # noqa: E402
from oracle import restate  # noqa: E402

DEV = torch.device("cuda", 0)
P10 = [0.01, 0.02, 0.03, 0.04, 0.05, 0.06, 0.07, 0.08, 0.09, 0.1]


def main():
    pcm = codes.rotated_surface_pcm(5)
    g = TannerGraph.from_pcm(pcm, DEV)
    ei = torch.from_numpy(codes.edge_index_of(pcm))
    x, _ = sample_syndromes(g, 2000, P10[:6], noise=1, seed=8)
    for name in ("epoch67", "epoch3", "fresh0", "fresh3"):
        if name.startswith("fresh"):
            torch.manual_seed(int(name[5:]))
            dec = decoder_v2_4.GNNI(15)
        else:
            dec = decoder_v2_4.GNNI(15)
            dec.load_state_dict(Golden("v2_4_toricL4_epoch67" if name == "epoch67" else "v2_4_toricL5_epoch3").weights)
        dec = dec.to(DEV).eval().bind_graph(g)
        w = {k: v.detach().cpu() for k, v in dec.state_dict().items()}
        ref = restate.decode("v2_4", ei, g.V, g.C, x.double().cpu(), w, T=15)["logit"]
        with options.option("GD_NO_LEAN"):
            l_old = dec.decode(x, return_logits=True)[1]
            with options.option("GD_NO_VTAB"):
                l_nv = dec.decode(x, return_logits=True)[1]
                with options.option("GD_NO_CTAB"):
                    l_dir = dec.decode(x, return_logits=True)[1]
        print(name, "edge-owner without variable tables: worst %.3f | all-direct: worst %.3f" % (logit_worst(l_nv.cpu(), ref)[0], logit_worst(l_dir.cpu(), ref)[0]))
        ls = [dec.decode(x, return_logits=True)[1] for _ in range(4)]
        print(name, "edge-owner: worst %.3f (max abs %.2e)" % logit_worst(l_old.cpu(), ref),
              "| table calls:", ["%.3f%s" % (logit_worst(l.cpu(), ref)[0], "=" if torch.equal(l, l_old) else "") for l in ls],
              "| max |logit| %.1f" % ref.abs().max().item(), flush=True)
    # training: which kernels take the step
    for d, B in ((7, 4096), (5, 8192)):
        Hz, Hx = codes.rotated_surface_checks(d)
        pcm = codes.css_pcm(Hz, Hx)
        g = TannerGraph.from_pcm(pcm, DEV)
        torch.manual_seed(0)
        dec = decoder_v2_4.GNNI(15).to(DEV).train().bind_graph(g)
        tr = FusedTrainer(dec, g, codes.css_logicals(Hz, Hx), lr=3e-4, weight_decay=1e-9)
        x, err = sample_syndromes(g, B, [0.01, 0.03, 0.05, 0.08], noise=1, seed=1)
        for it in range(6):
            loss = tr.step(x, err)
            torch.cuda.synchronize()
            stash = tr._bufs[B]["stash"]
            hdr = stash[(15 + 1) * 2 * g.E * B:(15 + 1) * 2 * g.E * B + 8].view(torch.int32).cpu().numpy()
            fm = np.array([hdr[4]], np.int32).view(np.float32)[0]
            print("train d=%d step %d: loss %.2f status %d old_count %d n_slots %d fmax %.3f vt_n %d" %
                  (d, it, loss.item(), hdr[0], hdr[1], hdr[2], fm, hdr[5]), flush=True)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(20):
            tr.step(x, err)
        ev[1].record()
        torch.cuda.synchronize()
        print("   %.3f ms/step" % (ev[0].elapsed_time(ev[1]) / 20), flush=True)
        with options.option("GD_NO_LEAN"):
            for _ in range(3):
                tr.step(x, err)
            ev[0].record()
            for _ in range(20):
                tr.step(x, err)
            ev[1].record()
            torch.cuda.synchronize()
        print("   edge-owner kernels: %.3f ms/step" % (ev[0].elapsed_time(ev[1]) / 20), flush=True)


if __name__ == "__main__":
    main()
