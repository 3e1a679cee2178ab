"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the cubic Hermite tables the decoder_v2_4 kernels
build in shared memory (gnn_decode_b200/csrc/gd_math.cuh: cubic_tab_build / cubic_tab_eval / cubic_tab_bound) and of the
a-priori interpolation-error bound they check before using a table.  The functions tabulated are the reference's MLPs
(quantum/decoder_v2_4.py:237-243 ggc1.mlp 2->128->1 and ggc2.mlp 1->128->1, :267-269 mlp 1->128->1, Softplus)."""
import numpy as np


def softplus(u):
    return np.logaddexp(0.0, u)


def mlp_1in(w1, b1, w2, b2):
    """y(x) = b2 + sum_k w2[k] softplus(w1[k] x + b1[k]); w1 [h], b1 [h], w2 [h]."""
    return lambda x: softplus(np.multiply.outer(np.asarray(x, np.float64), w1) + b1) @ w2 + b2


def mlp_2in_section(w1, b1, w2, b2, prior):
    """The variable-phase MLP at a fixed second input: f_p(ext) = mlp([ext, prior]); w1 [h, 2]."""
    return mlp_1in(w1[:, 0], w1[:, 1] * prior + b1, w2, b2)


def bound(w1_first_input, w2, step):
    """cubic_tab_bound: h^4 / 384 * max|f''''| with |softplus''''| <= 1/8."""
    m4 = float((np.abs(w2) * w1_first_input ** 4).sum())
    return step ** 4 * (0.125 / 384.0) * m4


def budget_half_width(w1_first_input, w2, n, budget):
    """Largest domain half-width R whose n-interval table meets `budget` (the kernel's vtab_R before clamping)."""
    m4 = float((np.abs(w2) * w1_first_input ** 4).sum())
    hmax = (budget * 384.0 / (0.125 * max(m4, 1e-20))) ** 0.25
    return 0.5 * n * hmax


def build(f, R, n, node_dtype=np.float64):
    """Coefficients [n, 4] of f on [-R, R]: node values f(x_i) (optionally rounded to node_dtype, like the kernel's fp32
    nodes), h f'(x_i) from 4th-order central differences of the node values, cubic Hermite form in t in [0, 1]."""
    h = 2.0 * R / n
    x = -R + h * np.arange(-2, n + 4)
    F = f(x).astype(node_dtype).astype(np.float64)
    a, b, f0, f1, c, d = F[0:n], F[1:n + 1], F[2:n + 2], F[3:n + 3], F[4:n + 4], F[5:n + 5]
    d0 = (8.0 * (f1 - b) - (c - a)) / 12.0
    d1 = (8.0 * (c - f0) - (d - b)) / 12.0
    return np.stack([f0, d0, 3.0 * (f1 - f0) - 2.0 * d0 - d1, 2.0 * (f0 - f1) + d0 + d1], axis=1)


def evaluate(coef, R, x):
    n = coef.shape[0]
    u = np.clip((np.asarray(x, np.float64) + R) * (n / (2.0 * R)), 0.0, n - 0.001)
    i = np.floor(u).astype(np.int64)
    t = u - i
    c = coef[i]
    return c[:, 0] + t * (c[:, 1] + t * (c[:, 2] + t * c[:, 3]))
