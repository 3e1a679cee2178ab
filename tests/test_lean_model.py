"""The algorithm of the check-owner table kernel (csrc/gd_lean.cu), emulated in numpy (oracle/lean_model.py: same tables,
same fp32 recurrence), against the fp64 oracle on the golden inputs -- checks on the CPU that (a) the per-edge recurrence
    t_e = g_p(m_sib(e)),  m_e += s_c f2(sum of the other t of the check),  logit_v = prior + sum f3(m_e)
IS decoder_v2_4's GNNI.forward on graphs with variable degree <= 2 / check degree <= 4, and (b) the table sizes the kernel
ships with (128 / 512 / 2048 pieces) meet their error budgets on the shipped checkpoints."""
import numpy as np
import pytest
import torch

from conftest import Golden, logit_worst
from oracle import lean_model, restate


@pytest.mark.parametrize("name", ["v2_4_toricL4_epoch1", "v2_4_toricL5_epoch3", "v2_4_toricL7_epoch3_T5"])
def test_table_recurrence_matches_the_oracle(name):
    g = Golden(name)
    w = {k: v.numpy() for k, v in g.weights.items()}
    out = lean_model.decode(g.edge_index.numpy(), g.V, g.C, g.x.numpy(), w, g.T)
    ref = restate.decode("v2_4", g.edge_index, g.V, g.C, g.x, g.weights, T=g.T)["logit"]
    worst, max_err = logit_worst(torch.from_numpy(out["logit"]), ref)
    assert worst <= 0.5, (worst, max_err)                       # measured: <= 0.06 of the bar
    e = out["errs"]
    fmax = e["Rm"] / g.T / 1.02
    assert e["c"] <= lean_model.BUDGET_C + 6e-8 * fmax           # check table: 1e-7 + half an fp32 ulp of its values
    assert e["r"] <= lean_model.BUDGET_R + 1e-7 * e["f3max"]
    assert max(e["v"].values()) <= e["budget_v"] == lean_model.BUDGET_V and e["vt_n"] == 512


def test_collapsed_checkpoint_needs_finer_variable_tables():
    """epoch67 (T max|mlp2| = 70, max|mlp3'| = 20: a table error is amplified ~4000 x on its way to the logits, the shipped
    checkpoints: ~200 x).  512 pieces on [-72, 72] miss even the plain 1e-6 budget; the piece-width rule's 1024 meet that (3.8e-7) but
    not the logit bar (1.1 x) -- the amplification-aware budget (1e-7 here) and piece width (0.097 instead of 0.172) make it 2048 pieces, which do (0.43 x)."""
    g = Golden("v2_4_toricL4_epoch67")
    w = {k: v.numpy() for k, v in g.weights.items()}
    ref = restate.decode("v2_4", g.edge_index, g.V, g.C, g.x, g.weights, T=g.T)["logit"]
    out = lean_model.decode(g.edge_index.numpy(), g.V, g.C, g.x.numpy()[:4], w, g.T, adaptive=False)
    assert max(out["errs"]["v"].values()) > lean_model.BUDGET_V
    out = lean_model.decode(g.edge_index.numpy(), g.V, g.C, g.x.numpy(), w, g.T, vt_n=1024, adaptive=False)
    e = out["errs"]
    assert e["budget_v"] < max(e["v"].values()) <= lean_model.BUDGET_V
    assert logit_worst(torch.from_numpy(out["logit"]), ref)[0] > 1.0          # the reason the budget has to know about the gain
    out = lean_model.decode(g.edge_index.numpy(), g.V, g.C, g.x.numpy(), w, g.T)
    e = out["errs"]
    assert e["vt_n"] == 2048 and max(e["v"].values()) <= e["budget_v"] < 2e-7
    assert e["vt_n_first"] == 2048            # ... at once: the piece-width rule knows the amplification (estimated in the prep pass)
    worst, max_err = logit_worst(torch.from_numpy(out["logit"]), ref)
    assert worst <= 0.5, (worst, max_err)


def test_piece_rule_keeps_the_shipped_checkpoints_on_the_base_tables():
    assert lean_model.vt_pieces(512, 34.0) == 512 and lean_model.vt_pieces(512, 43.9) == 512
    assert lean_model.vt_pieces(512, 71.0) == 1024 and lean_model.vt_pieces(512, 96.0) == 2048
    assert lean_model.vt_pieces(512, 1e4) == 4096                # capped: the error check then sends the batch to the edge-owner kernel


def test_midpoint_error_estimate_is_the_true_error():
    """The kernel measures a table's INTERPOLATION error at the interval midpoints only (the Hermite remainder
    (t (1 - t))^2 f''''/24 peaks there).  On a dense grid the fp32 look-up adds what any fp32 evaluation adds: the interval
    coordinate w = x n/6 + off is rounded at magnitude ~n (half an ulp of 128 = 3.8e-6 intervals = 1.8e-7 in x), times |f'| <= ~2,
    plus the rounding of the value itself (|f| ~ 2-3: 1.2e-7)."""
    g = Golden("v2_4_toricL5_epoch3")
    gw = lambda k: g.weights[k].numpy().astype(np.float32).astype(np.float64)
    a, c, w2, b2 = gw("ggc2.mlp.0.weight")[:, 0], gw("ggc2.mlp.0.bias"), gw("ggc2.mlp.2.weight")[0], float(gw("ggc2.mlp.2.bias")[0])
    n = 128
    coef, err_mid, _ = lean_model.build_table(a, c, w2, b2, 3.0, n)
    x = np.linspace(-3.0, 3.0, 20001)
    f, _ = lean_model.mlp_f_df(a, c, w2, b2, x)
    inv_h = n / 6.0
    got = lean_model.eval_table(coef, inv_h, 3.0 * inv_h - 0.5, x)
    assert np.abs(got.astype(np.float64) - f).max() <= err_mid + 6e-7
