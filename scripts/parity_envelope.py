"""Measured error envelope of the CUDA paths against the fp64 oracle / the reference's own outputs, over every golden case
(logits of all programs, resident and streamed) and the gradient fixtures -- the numbers the parity bars in tests/ are set
from.  Developer script (the oracle is the checker); writes a text table.

    python scripts/parity_envelope.py > profiles/r02_parity_envelope.txt
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import Golden, golden_cases, logit_bound, make_decoder  # noqa: E402
from gnn_decode_b200 import options  # noqa: E402
from gnn_decode_b200.graph import TannerGraph  # noqa: E402
from oracle import restate  # noqa: E402

DEV = torch.device("cuda", 0)


def row(tag, got, ref, rtol):
    ref = ref.double().cpu()
    err = (got.double().cpu() - ref).abs()
    big = ref.abs() > 0.1
    rel = (err[big] / ref.abs()[big]).max().item() if bool(big.any()) else 0.0
    print("%-44s max|d| %.3g  max rel(|l|>0.1) %.3g  worst/bar(rtol %.0e) %.3f  rms|l| %.3g" %
          (tag, err.max().item(), rel, rtol, (err / logit_bound(ref, rtol)).max().item(), ref.pow(2).mean().sqrt().item()))


def main():
    print("# logits: CUDA fp32 vs fp64 oracle (oracle/restate.py), bar = rtol * max(|ref|, 1)")
    for name in golden_cases():
        g = Golden(name)
        mod, dec = make_decoder(g)
        dec = dec.to(DEV).eval()
        tg = TannerGraph(g.edge_index, g.V, g.C, DEV)
        ref = restate.decode(g.program, g.edge_index, g.V, g.C, g.x, g.weights, T=g.T, dtype=torch.float64)["logit"]
        rtol = 2e-3 if g.program.startswith("bp") else 1e-4
        x = g.x.repeat(4, 1).to(DEV)
        _, logit, _ = dec.decode(x, graph=tg, return_logits=True, return_hard=True)
        row(name + " [default path]", logit[:g.B], ref, rtol)
        if g.program == "v2_4":        # the table kernel may spend its first calls raising the table resolution: steady state
            for _ in range(3):
                _, logit, _ = dec.decode(x, graph=tg, return_logits=True, return_hard=True)
            row(name + " [default path, 4th call]", logit[:g.B], ref, rtol)
        if g.program == "v2_4":
            with options.option("GD_NO_LEAN"):
                _, l2, _ = dec.decode(x, graph=tg, return_logits=True, return_hard=True)
            row(name + " [edge-owner kernel]", l2[:g.B], ref, rtol)
            with options.option("GD_NO_LEAN"), options.option("GD_NO_VTAB"), options.option("GD_NO_CTAB"):
                _, l3, _ = dec.decode(x, graph=tg, return_logits=True, return_hard=True)
            row(name + " [all-direct evaluation]", l3[:g.B], ref, rtol)
        with options.option("GD_FORCE_STREAMED"):
            try:
                _, l4, _ = dec.decode(x, graph=tg, return_logits=True, return_hard=True)
                row(name + " [streamed]", l4[:g.B], ref, rtol)
            except Exception as ex:  # noqa: BLE001
                print("%-44s streamed: %s" % (name, ex))
    print("# widened programs (fused kernels) vs the fp64 oracle")
    import importlib
    ext = {"ext_neural_bp_toricL4": "quantum.neural_BP", "ext_gru_ca_toricL4": "quantum.QGNNNI_ca", "ext_v1_1_onehot_toricL4": None,
           "ext_v2_4_1_toricL4": "quantum.decoder_v2_4_1", "ext_v3_0_toricL4": "quantum.decoder_v3_0", "ext_v1_2_2_toricL4": "quantum.decoder_v1_2_2"}
    for name, modname in ext.items():
        if modname is None:
            continue
        g = Golden(name)
        mod = importlib.import_module("gnn_decode_b200." + modname)
        dec = mod.GNNI(g.T, n_edges=g.E) if g.program == "neural_bp" else mod.GNNI(g.T)
        dec.load_state_dict(g.weights, strict=True)
        dec = dec.to(DEV).eval()
        tg = TannerGraph(g.edge_index, g.V, g.C, DEV)
        ref = restate.decode(g.program, g.edge_index, g.V, g.C, g.x, g.weights, T=g.T, dtype=torch.float64)
        rtol = 2e-3 if g.program == "neural_bp" else 1e-4
        x = g.x.to(DEV)
        if g.program == "v3_0":
            _, _, l, lc = dec.decode_aux(x, graph=tg, return_logits=True)
            row(name + " [variable read-out]", l, ref["logit"], rtol)
            row(name + " [check read-out]", lc, ref["logit_chk"], rtol)
        elif g.program in ("v1_2_2", "gru_ca"):
            _, la = dec.decode_all(x, graph=tg, return_logits=True)
            row(name + " [all %d iterations]" % g.T, la, ref["all_logit"], rtol)
        else:
            row(name, dec.decode(x, graph=tg, return_logits=True)[1], ref["logit"], rtol)
    print("# gradients: CUDA backward vs the reference's loss.backward() fixtures (fp64): max |dg| / max |g| per tensor")
    import test_training_gpu as tt
    from gnn_decode_b200.quantum import decoder_v2_4
    for name in ("grad_v2_4_toricL4_epoch1", "grad_v2_4_toricL5_epoch3_T6"):
        case = tt._Case(name)
        dec = decoder_v2_4.GNNI(case.T)
        dec.load_state_dict(case.weights)
        dec = dec.to(DEV).train()
        loss, grads, pred = tt._train_step(case, dec)
        worst = 0.0
        for k, gref in case.grads.items():
            e = (grads[k].double().cpu() - gref).abs().max().item() / gref.abs().max().item()
            worst = max(worst, e)
            print("  %-40s %-24s %.3g" % (name, k, e))
        print("  %-40s loss rel err %.3g   worst tensor %.3g" % (name, abs(loss - case.loss) / abs(case.loss), worst))


if __name__ == "__main__":
    main()
