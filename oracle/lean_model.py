"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy emulation of the check-owner table kernel
(gnn_decode_b200/csrc/gd_lean.cu) -- the same tables (double-precision nodes and analytic derivatives, cubic Hermite pieces
re-centred on the interval midpoints and rounded to fp32, piece -1 and piece n included) and the same fp32 per-edge recurrence

    t_e  = g_p(m_sib(e))   (g_p(0) where the variable has no sibling edge)          g_p = tanh(ggc1.mlp([., prior]) / 2)
    m_e += s_c * f2(sum of t over the other edges of the check)                       f2  = ggc2.mlp
    logit_v = prior + sum_{e at v} f3(m_e)                                            f3  = mlp

that the kernel runs for decoder_v2_4 (quantum/decoder_v2_4.py:272-294) on graphs with variable degree <= 2 and check
degree <= 4.  Used on CPU to check the algorithm and its table budgets against oracle/restate.py (fp64)."""
import numpy as np

BUDGET_C, BUDGET_V, BUDGET_R = 1e-7, 1e-6, 4e-6


def _softplus_pair(z):
    """softplus(z) and sigmoid(z) with torch.nn.Softplus(beta=1, threshold=20) semantics."""
    e = np.exp(-np.abs(z))
    sp = np.where(z > 20.0, z, np.maximum(z, 0.0) + np.log1p(e))
    sg = np.where(z > 20.0, 1.0, np.where(z >= 0.0, 1.0 / (1.0 + e), e / (1.0 + e)))
    return sp, sg


def mlp_f_df(a, c, w2, b2, x):
    """f(x) = b2 + sum_k w2_k softplus(a_k x + c_k) and f'(x), in double."""
    z = np.multiply.outer(np.asarray(x, np.float64), a) + c
    sp, sg = _softplus_pair(z)
    return sp @ w2 + b2, sg @ (w2 * a)


def build_table(a, c, w2, b2, Rdom, n, tanh_fold=False):
    """Pieces -1 .. n (n + 2 rows of 4 fp32 coefficients in tau in [-0.5, 0.5]) and the midpoint error of every interval."""
    h = 2.0 * Rdom / n
    x = -Rdom + h * np.arange(-1, n + 2)
    f, df = mlp_f_df(a, c, w2, b2, x)
    fmax = float(np.abs(f).max())
    build_table.last_dmax = float(np.abs(df).max())              # max |f'| over the nodes (the kernel keeps it for mlp2 and mlp3)
    if tanh_fold:
        g = np.tanh(0.5 * f)
        df = 0.5 * (1.0 - g * g) * df
        f = g
    d = h * df
    f0, d0, f1, d1 = f[:-1], d[:-1], f[1:], d[1:]
    c2 = 3.0 * (f1 - f0) - 2.0 * d0 - d1
    c3 = 2.0 * (f0 - f1) + d0 + d1
    coef = np.stack([f0 + 0.5 * d0 + 0.25 * c2 + 0.125 * c3, d0 + c2 + 0.75 * c3, c2 + 1.5 * c3, c3], 1).astype(np.float32)
    xm = x[:-1] + 0.5 * h
    fm, _ = mlp_f_df(a, c, w2, b2, xm)
    if tanh_fold:
        fm = np.tanh(0.5 * fm)
    err = float(np.abs(coef[:, 0].astype(np.float64) - fm).max())
    return coef, err, fmax


def eval_table(coef, inv_h, off, x, clamp=None):
    """fp32 look-up as in lean_cubic: w = x * inv_h + off, piece = round(w), tau = w - piece."""
    w = np.float32(x) * np.float32(inv_h) + np.float32(off)
    w = w.astype(np.float32)
    if clamp is not None:
        w = np.clip(w, np.float32(clamp[0]), np.float32(clamp[1]))
    i = np.rint(w).astype(np.int64)
    tau = (w - i.astype(np.float32)).astype(np.float32)
    c = coef[i + 1]
    r = c[:, 3]
    for k in (2, 1, 0):
        r = (r * tau + c[:, k]).astype(np.float32)
    return r


def vt_pieces(base, Rm, mult=1, gain=0.0):
    """Pieces of the variable-phase tables (gd_lean.cu: lean_vt_pieces): the base count, doubled (up to 8x) until a piece is
    narrower than 0.172 -- narrower by (400 / gain)^(1/4) for a model that amplifies table errors by more than 400 -- and there are
    at least base * mult of them."""
    wmax = 0.172 * (400.0 / gain) ** 0.25 if gain > 400.0 else 0.172
    n = base
    while n < 8 * base and (2.0 * Rm / n > wmax or n < base * mult):
        n *= 2
    return n


def budget_v(T, d2max, d3max):
    """Error budget of the variable-phase tables (gd_lean.cu: lean_decode_kernel): 1e-6, less when the model amplifies an error in
    t by more than 400 on its way to the logits (first order: 6 T max|mlp2'| max|mlp3'|)."""
    gain = 6.0 * T * d2max * d3max
    return 4e-4 / gain if gain > 400.0 else BUDGET_V


def decode(edge_index, V, C, x, w, T, ct_n=128, vt_n=512, rt_n=2048, adaptive=True):
    """x [B, V+C] with one prior per syndrome and +-1 check inputs -> dict(logit, errs).  w: reference state_dict (numpy).
    vt_n is the BASE piece count of the variable-phase tables; adaptive=False keeps it whatever the message domain, otherwise the
    kernel's rule applies: the piece-width rule first, then doubled while a table of the batch misses budget_v (the kernel does that
    from one call to the next; the calls in between are served by the edge-owner kernel)."""
    ei = np.asarray(edge_index)
    var, chk = ei[0], ei[1]
    E = ei.shape[1]
    x = np.asarray(x, np.float64)
    B = x.shape[0]
    prior = x[:, 0].astype(np.float32)
    assert np.all(x[:, :V] == x[:, :1]) and np.all(np.abs(x[:, V:]) == 1.0)
    sgn = x[:, V:].astype(np.float32)
    g = lambda k: np.asarray(w[k], np.float32).astype(np.float64)          # the kernel sees the fp32 copy of the weights
    W1, b1, w2v, b2v = g("ggc1.mlp.0.weight"), g("ggc1.mlp.0.bias"), g("ggc1.mlp.2.weight")[0], float(g("ggc1.mlp.2.bias")[0])
    a2, c2_, w22, b22 = g("ggc2.mlp.0.weight")[:, 0], g("ggc2.mlp.0.bias"), g("ggc2.mlp.2.weight")[0], float(g("ggc2.mlp.2.bias")[0])
    a3, c3_, w23, b23 = g("mlp.0.weight")[:, 0], g("mlp.0.bias"), g("mlp.2.weight")[0], float(g("mlp.2.bias")[0])
    ctab, err_c, fmax = build_table(a2, c2_, w22, b22, 3.0, ct_n)
    d2max = build_table.last_dmax
    fmax32 = np.nextafter(np.float32(fmax), np.float32(np.inf)) if np.float32(fmax) < fmax else np.float32(fmax)
    Rm = float(np.float32(T) * (fmax32 * np.float32(1.02) + np.float32(1e-6)))
    rtab, err_r, f3max = build_table(a3, c3_, w23, b23, Rm, rt_n)
    d3max = build_table.last_dmax
    bv = budget_v(T, d2max, d3max)
    if adaptive:
        base, mult = vt_n, 1
        # the kernel's first choice uses max |mlp3'| on a coarse grid over [-352, 352] (the prep kernel's estimate)
        d3_est = float(np.abs(mlp_f_df(a3, c3_, w23, b23, np.linspace(-352.0, 352.0, 1025))[1]).max())
        errs_first = vt_pieces(base, Rm, 1, 6.0 * T * d2max * d3_est)
        while True:
            vt_n = vt_pieces(base, Rm, mult, 6.0 * T * d2max * d3_est)
            worst = max(build_table(W1[:, 0], W1[:, 1] * float(p) + b1, w2v, b2v, Rm, vt_n, tanh_fold=True)[1] for p in np.unique(prior))
            if worst <= bv or vt_n >= 8 * base:
                break
            mult = 2 * vt_n // base
    errs = {"c": err_c, "r": err_r, "v": {}, "Rm": Rm, "f3max": f3max, "vt_n": vt_n, "budget_v": bv, "vt_n_first": errs_first if adaptive else vt_n}
    # sibling edge of each edge at its variable, other edges of each edge at its check
    sib = -np.ones(E, np.int64)
    for v in range(V):
        es = np.nonzero(var == v)[0]
        assert len(es) <= 2
        if len(es) == 2:
            sib[es[0]], sib[es[1]] = es[1], es[0]
    others = []
    for e in range(E):
        es = [q for q in np.nonzero(chk == chk[e])[0] if q != e]
        assert len(es) <= 3
        others.append(es)
    ct_inv_h, ct_off = ct_n / 6.0, 3.0 * (ct_n / 6.0) - 0.5
    vt_inv_h = 0.5 * vt_n / Rm
    vt_off = Rm * vt_inv_h - 0.5
    rt_inv_h = 0.5 * rt_n / Rm
    rt_off = Rm * rt_inv_h - 0.5
    logit = np.zeros((B, V), np.float32)
    for p in np.unique(prior):
        rows = np.nonzero(prior == p)[0]
        vtab, err_v, _ = build_table(W1[:, 0], W1[:, 1] * float(p) + b1, w2v, b2v, Rm, vt_n, tanh_fold=True)
        errs["v"][float(p)] = err_v
        n = len(rows)
        m = np.zeros((n, E), np.float32)
        t0 = eval_table(vtab, vt_inv_h, vt_off, np.zeros(n, np.float32), (-1.4, vt_n + 0.4))
        for _ in range(T):
            t = np.empty((n, E), np.float32)
            for e in range(E):
                t[:, e] = t0 if sib[e] < 0 else eval_table(vtab, vt_inv_h, vt_off, m[:, sib[e]], (-1.4, vt_n + 0.4))
            mn = np.empty_like(m)
            for e in range(E):
                ext = np.zeros(n, np.float32)
                for q in others[e]:
                    ext = (ext + t[:, q]).astype(np.float32)
                o = eval_table(ctab, ct_inv_h, ct_off, ext)
                mn[:, e] = (m[:, e] + o * sgn[rows, chk[e]]).astype(np.float32)
            m = mn
        lg = np.repeat(np.float32(p), n * V).reshape(n, V).astype(np.float32)
        for e in range(E):
            lg[:, var[e]] = (lg[:, var[e]] + eval_table(rtab, rt_inv_h, rt_off, m[:, e], (-1.4, rt_n + 0.4))).astype(np.float32)
        logit[rows] = lg
    return {"logit": logit, "errs": errs}
