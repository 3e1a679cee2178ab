"""The cubic tables of the decoder_v2_4 kernels, restated in numpy (oracle/tables.py), on the shipped
checkpoints: whenever the kernel's a-priori bound accepts a table, the table really is that accurate -- for the check-phase
MLP (budget 1e-7 on |ext| <= 3), the read-out MLP (4e-6 on |m| <= T max|mlp2|) and the per-prior sections of the
variable-phase MLP (2e-7 on the budget-derived domain)."""
import os

import numpy as np
import pytest

from oracle import tables

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CKPTS = ["v2_4_toricL4_epoch1", "v2_4_toricL5_epoch3", "v2_4_toricL4_epoch67"]


def _w(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = lambda k: z["w:" + k].astype(np.float64)
    return {"v": (g("ggc1.mlp.0.weight"), g("ggc1.mlp.0.bias"), g("ggc1.mlp.2.weight")[0], float(g("ggc1.mlp.2.bias")[0])),
            "c": (g("ggc2.mlp.0.weight")[:, 0], g("ggc2.mlp.0.bias"), g("ggc2.mlp.2.weight")[0], float(g("ggc2.mlp.2.bias")[0])),
            "r": (g("mlp.0.weight")[:, 0], g("mlp.0.bias"), g("mlp.2.weight")[0], float(g("mlp.2.bias")[0])),
            "T": int(z["T"]), "priors": np.unique(z["x"][:, :int(z["V"])])}


def _max_err(f, coef, R, per_interval=9):
    n = coef.shape[0]
    x = (-R + (2.0 * R / n) * (np.arange(n)[:, None] + np.linspace(0.0, 1.0, per_interval)[None, :])).ravel()
    # the look-up clamps the interval coordinate to n - 0.001 (an exactly saturated input, x == R, is evaluated 0.001 h short of
    # the end: |f'| h 1e-3 ~ 1e-5 there, see test_saturated_end_point); everything up to that point is checked here
    x = np.clip(x, -R, R - 0.002 * (2.0 * R / n))
    return float(np.abs(tables.evaluate(coef, R, x) - f(x)).max())


@pytest.mark.parametrize("name", CKPTS)
def test_check_phase_and_readout_tables_meet_their_budgets(name):
    w = _w(name)
    f = tables.mlp_1in(*w["c"])
    R, n = 3.0, 512                                     # surface / toric codes: check degree 4
    assert tables.bound(w["c"][0], w["c"][2], 2 * R / n) <= 1e-7
    coef = tables.build(f, R, n)
    assert _max_err(f, coef, R) <= 1e-7
    fmax = float(np.abs(coef[:, 0]).max())
    Rr, nr = w["T"] * (fmax * 1.02 + 1e-6), 2048        # |m| <= T max|mlp2|
    fr = tables.mlp_1in(*w["r"])
    b = tables.bound(w["r"][0], w["r"][2], 2 * Rr / nr)
    if b <= 4e-6:                                       # the kernel uses the table only then
        assert _max_err(fr, tables.build(fr, Rr, nr), Rr) <= b + 1e-12


@pytest.mark.parametrize("name", CKPTS)
def test_variable_phase_sections_meet_their_budget(name):
    w = _w(name)
    w1, b1, w2, b2 = w["v"]
    n, budget = 512, 2e-7
    R = min(tables.budget_half_width(w1[:, 0], w2, n, budget), 1e4)
    assert R >= 0.5                                     # else the kernel keeps the direct evaluation
    assert tables.bound(w1[:, 0], w2, 2 * R / n) <= budget * (1 + 1e-9)
    for prior in w["priors"][:6]:
        f = tables.mlp_2in_section(w1, b1, w2, b2, float(prior))
        err = _max_err(f, tables.build(f, R, n), R)
        assert err <= budget, (prior, err)
        # with fp32 node values (what the kernel stores) the table sits at the fp32 rounding floor of the function itself
        err32 = _max_err(f, tables.build(f, R, n, np.float32), R)
        scale = float(np.abs(f(np.linspace(-R, R, 257))).max())
        assert err32 <= budget + 4 * np.finfo(np.float32).eps * scale, (prior, err32, scale)


def test_bound_is_a_bound_on_random_mlps():
    rng = np.random.default_rng(0)
    for _ in range(20):
        h = 128
        w1 = rng.standard_normal(h) * rng.uniform(0.2, 3.0)
        b1 = rng.standard_normal(h)
        w2 = rng.standard_normal(h) * rng.uniform(0.05, 0.5)
        f = tables.mlp_1in(w1, b1, w2, 0.1)
        n = 256
        R = min(tables.budget_half_width(w1, w2, n, 1e-6), 50.0)
        err = _max_err(f, tables.build(f, R, n), R)
        assert err <= tables.bound(w1, w2, 2 * R / n) * (1 + 1e-6) + 1e-13


def test_saturated_end_point():
    """x == +R exactly (all sibling tanh saturated at +-1) is looked up 0.001 intervals short of the end of the table: the
    error there is |f'(R)| * h * 1e-3, ~1e-5 absolute for the shipped weights -- 5e-6 of the parity bar's scale."""
    w = _w("v2_4_toricL5_epoch3")
    f = tables.mlp_1in(*w["c"])
    R, n = 3.0, 512
    coef = tables.build(f, R, n)
    h = 2 * R / n
    slope = abs(float(f(R) - f(R - 1e-6))) / 1e-6
    err = abs(float(tables.evaluate(coef, R, np.array([R]))[0] - f(R)))
    assert err <= 1.1e-3 * h * slope + 1e-7
    assert err <= 2e-5
