"""Parity-check matrices for the decoders (host side, NumPy).

Conventions follow the reference (quantum/error_generate.py:12-14 and SURVEY.md section 9):
  * a quantum CSS code on n qubits has V = 2n variable "slots": X-error slots [0, n) then Z-error
    slots [n, 2n); the parity-check matrix `pcm` is [C, V] with the Z-type checks (which detect X
    errors) in the first rows acting on the X slots, and the X-type checks in the last rows acting
    on the Z slots;
  * the model code works with H = pcm.T ([V, C], the reference's `H`), whose row-major COO is the
    per-graph edge_index.

toric_pcm(L) reproduces the reference's generate_PCM(2L^2-2, L)[0] bit for bit (tests pin it against
fixtures produced by the reference); rotated surface and hypergraph-product codes are NOT in the
reference and are supplied here because BASELINE.json's configs name them.
"""
import numpy as np


# ---- GF(2) linear algebra ---------------------------------------------------------------------
def gf2_rref(A):
    """Reduced row echelon form over GF(2). Returns (R, pivot_columns)."""
    A = (np.array(A, dtype=np.uint8) & 1).copy()
    rows, cols = A.shape
    piv, r = [], 0
    for c in range(cols):
        if r == rows:
            break
        nz = np.flatnonzero(A[r:, c])
        if nz.size == 0:
            continue
        p = r + nz[0]
        if p != r:
            A[[r, p]] = A[[p, r]]
        others = np.flatnonzero(A[:, c])
        others = others[others != r]
        A[others] ^= A[r]
        piv.append(c)
        r += 1
    return A[:r], piv


def gf2_rank(A):
    return len(gf2_rref(A)[1])


def gf2_nullspace(A):
    """Basis (rows) of {x : A x = 0 mod 2}."""
    A = np.array(A, dtype=np.uint8) & 1
    R, piv = gf2_rref(A)
    cols = A.shape[1]
    pset = set(piv)
    free = [c for c in range(cols) if c not in pset]
    basis = np.zeros((len(free), cols), dtype=np.uint8)
    if free:
        basis[np.arange(len(free)), free] = 1
        if piv:
            basis[:, piv] = R[:, free].T          # x_pivot = sum of the free columns it depends on
    return basis


def _row_ints(A):
    """Rows of a 0/1 matrix as Python ints (bit c of the int = column c): XOR and lowest-set-bit on arbitrary-precision
    ints are far faster than per-row NumPy calls at these sizes."""
    A = np.ascontiguousarray(np.array(A, dtype=np.uint8) & 1)
    if A.size == 0:
        return []
    packed = np.packbits(A, axis=1, bitorder="little")
    return [int.from_bytes(r.tobytes(), "little") for r in packed]


def _complete_basis(sub, full):
    """Rows of `full` that extend the row space of `sub` (greedy in row order, GF(2)): a row is kept iff it is independent
    of `sub` and of the rows kept before it.  Incremental elimination on bit-packed rows: each vector is reduced against
    the pivots found so far (O(rank) XORs) instead of re-ranking the whole stack per candidate -- the hypergraph-product
    [[1600,64]] code takes 0.3 s instead of a minute (the reference's H_Prep, error_generate.py:145-248, is O(n^3) Python)."""
    full = np.array(full, dtype=np.uint8).reshape(-1, full.shape[1]) & 1
    pivots = {}                                   # lowest set bit -> reduced row

    def reduce(v):
        while v:
            low = v & -v
            b = pivots.get(low)
            if b is None:
                return v, low
            v ^= b
        return 0, 0

    for v in _row_ints(np.array(sub, dtype=np.uint8).reshape(-1, full.shape[1])):
        v, low = reduce(v)
        if v:
            pivots[low] = v
    keep = []
    for i, v in enumerate(_row_ints(full)):
        v, low = reduce(v)
        if v:
            pivots[low] = v
            keep.append(i)
    return full[keep].reshape(-1, full.shape[1])


def css_logicals(Hz, Hx):
    """Logical operators of the CSS code with Z-checks Hz [cz, n] and X-checks Hx [cx, n], laid out
    to act on an error vector r = [X-error slots | Z-error slots] exactly like the reference's
    `logical` (H_Prep.get_logical, error_generate.py:211-248): a decode fails logically iff
    logical @ r is odd in some row.  Rows: logical-Z representatives (in ker Hx, outside
    rowspace Hz) on the X slots, then logical-X representatives on the Z slots."""
    Hz = np.array(Hz, dtype=np.uint8) & 1
    Hx = np.array(Hx, dtype=np.uint8) & 1
    n = Hz.shape[1]
    assert not ((Hz.astype(np.int64) @ Hx.T.astype(np.int64)) % 2).any(), "checks do not commute"
    lz = _complete_basis(Hz, gf2_nullspace(Hx))   # Z-type logicals: commute with X checks
    lx = _complete_basis(Hx, gf2_nullspace(Hz))
    out = np.zeros((lz.shape[0] + lx.shape[0], 2 * n), dtype=np.uint8)
    out[:lz.shape[0], :n] = lz
    out[lz.shape[0]:, n:] = lx
    return out


def css_pcm(Hz, Hx):
    """blockdiag(Hz, Hx): [cz + cx, 2n]."""
    Hz = np.array(Hz, dtype=np.uint8)
    Hx = np.array(Hx, dtype=np.uint8)
    n = Hz.shape[1]
    pcm = np.zeros((Hz.shape[0] + Hx.shape[0], 2 * n), dtype=np.uint8)
    pcm[:Hz.shape[0], :n] = Hz
    pcm[Hz.shape[0]:, n:] = Hx
    return pcm


# ---- toric code (the reference's code family) ---------------------------------------------------
def toric_pcm(L):
    """[2L^2-2, 4L^2] parity-check matrix of the L x L toric code with the last plaquette and the
    last star generator dropped -- identical to generate_PCM(2*L*L-2, L)[0] of
    quantum/error_generate.py:39-132 (qubit j = 2L*a + b is the horizontal edge of cell (a, b),
    j + L its vertical edge)."""
    n, k = 2 * L * L, L * L - 1
    pcm = np.zeros((2 * k, 2 * n), dtype=np.uint8)
    for i in range(k):
        a, b = divmod(i, L)
        j = 2 * L * a + b
        # Z-type plaquette on the X slots: top, left, bottom, right edges
        for q in (j, j + L, (j + 2 * L) % n, j + 1 if b == L - 1 else j + L + 1):
            pcm[i, q] = 1
        # X-type star on the Z slots: right, down, up, left edges
        for q in (j, j + L, (j - L) % n, j - 1 if b != 0 else j - 1 + L):
            pcm[k + i, n + q] = 1
    return pcm


def toric_H(L):
    """The reference's `H` = generate_PCM(...)[0].T : [V = 4L^2, C = 2L^2 - 2]."""
    return np.ascontiguousarray(toric_pcm(L).T)


# ---- rotated surface code (BASELINE config 2/3/4; not in the reference) ----------------------------
def rotated_surface_checks(d):
    """(Hz, Hx) of the distance-d rotated surface code on a d x d data-qubit grid, (d^2-1)/2 checks
    each.  Plaquette (i, j), 0 <= i, j <= d, covers data qubits {i-1, i} x {j-1, j} inside the grid;
    it is X-type when i + j is even, Z-type otherwise; weight-2 boundary plaquettes are kept on the
    top/bottom rows when X-type and on the left/right columns when Z-type."""
    assert d >= 3 and d % 2 == 1
    hz, hx = [], []
    for i in range(d + 1):
        for j in range(d + 1):
            qs = [r * d + c for r in (i - 1, i) for c in (j - 1, j) if 0 <= r < d and 0 <= c < d]
            x_type = (i + j) % 2 == 0
            if len(qs) == 4:
                keep = True
            elif len(qs) == 2:
                keep = (x_type and i in (0, d)) or ((not x_type) and j in (0, d))
            else:
                keep = False
            if keep:
                row = np.zeros(d * d, dtype=np.uint8)
                row[qs] = 1
                (hx if x_type else hz).append(row)
    Hz, Hx = np.array(hz), np.array(hx)
    assert Hz.shape[0] == Hx.shape[0] == (d * d - 1) // 2
    return Hz, Hx


def rotated_surface_pcm(d):
    return css_pcm(*rotated_surface_checks(d))


# ---- hypergraph-product code (BASELINE config 5; not in the reference) -----------------------------
def regular_ldpc(n_checks, n_vars, col_w, row_w, seed=1234, max_tries=1000):
    """Random (col_w, row_w)-regular [n_checks, n_vars] matrix without double edges (configuration
    model with rejection), seeded."""
    assert n_vars * col_w == n_checks * row_w
    rng = np.random.RandomState(seed)
    for _ in range(max_tries):
        stubs_v = np.repeat(np.arange(n_vars), col_w)
        stubs_c = np.repeat(np.arange(n_checks), row_w)
        rng.shuffle(stubs_v)
        Hm = np.zeros((n_checks, n_vars), dtype=np.int64)
        np.add.at(Hm, (stubs_c, stubs_v), 1)
        if Hm.max() == 1:
            return Hm.astype(np.uint8)
    raise RuntimeError("could not sample a simple regular bipartite graph")


def hgp_checks(H1, H2=None):
    """Hypergraph product of classical codes H1 [r1, n1], H2 [r2, n2] -> (Hz, Hx) on
    n = n1*n2 + r1*r2 qubits:  Hx = [H1 (x) I_n2 | I_r1 (x) H2^T],  Hz = [I_n1 (x) H2 | H1^T (x) I_r2]."""
    H1 = np.array(H1, dtype=np.uint8)
    H2 = H1 if H2 is None else np.array(H2, dtype=np.uint8)
    r1, n1 = H1.shape
    r2, n2 = H2.shape
    Hx = np.hstack([np.kron(H1, np.eye(n2, dtype=np.uint8)), np.kron(np.eye(r1, dtype=np.uint8), H2.T)])
    Hz = np.hstack([np.kron(np.eye(n1, dtype=np.uint8), H2), np.kron(H1.T, np.eye(r2, dtype=np.uint8))])
    return Hz & 1, Hx & 1


def hgp_pcm(n_checks=24, n_vars=32, col_w=3, row_w=4, seed=1234):
    """[[n_vars^2 + n_checks^2, .]] HGP code of a seeded (col_w, row_w)-regular seed; the default
    24 x 32 seed gives the [[1600, 64]] code of BASELINE config 5 (V = 3200, C = 1536, E = 10752)."""
    seedH = regular_ldpc(n_checks, n_vars, col_w, row_w, seed)
    return css_pcm(*hgp_checks(seedH))


# ---- classical codes of classical/CGNNI.py ---------------------------------------------------------
def ldpc_toy_pcm():
    """The 4 x 8 toy LDPC hard-coded at classical/CGNNI.py:187-190 (the smallest bundled code)."""
    return np.array([[0, 1, 0, 1, 1, 0, 0, 1],
                     [1, 1, 1, 0, 0, 1, 0, 0],
                     [0, 0, 1, 0, 0, 1, 1, 1],
                     [1, 0, 0, 1, 1, 0, 1, 0]], dtype=np.uint8)


def _gf2_poly_divmod(num, den):
    num = list(num)
    dd = len(den) - 1
    q = [0] * (len(num) - dd)
    for i in range(len(num) - 1, dd - 1, -1):
        if num[i]:
            q[i - dd] = 1
            for j, c in enumerate(den):
                num[i - dd + j] ^= c
    return q, num[:dd]


def bch_63_45_pcm():
    """18 x 63 parity-check matrix of the cyclic BCH(63,45) code: row i is the check polynomial
    h(x) = (x^63 + 1) / g(x), coefficients highest degree first, shifted right by i -- the same
    matrix as the reference's classical/`BCH(63,45).txt` (tests pin it against a fixture).
    g(x) = 1 + x + x^2 + x^3 + x^6 + x^7 + x^9 + x^15 + x^16 + x^17 + x^18  (octal 1701317)."""
    g = [0] * 19
    for e in (0, 1, 2, 3, 6, 7, 9, 15, 16, 17, 18):
        g[e] = 1
    num = [0] * 64
    num[0] = num[63] = 1
    h, rem = _gf2_poly_divmod(num, g)
    assert not any(rem) and len(h) == 46
    first = np.zeros(63, dtype=np.uint8)
    first[:46] = h[::-1]
    return np.stack([np.roll(first, i) for i in range(18)])


def read_pcm_txt(path, transpose=False):
    """Read a parity-check matrix in the reference's text format (`classical/BCH(63,45).txt`: whitespace-separated
    0/1 rows, loaded there with np.loadtxt, CGNNI.py:181).  Returns uint8 [C, V] (transpose=True for files stored [V, C])."""
    H = np.loadtxt(path).astype(np.uint8)
    H = H.reshape(1, -1) if H.ndim == 1 else H
    return np.ascontiguousarray(H.T if transpose else H)


def write_pcm_txt(path, pcm):
    """Write a [C, V] 0/1 matrix in the same text format (one row per line, single spaces)."""
    np.savetxt(path, np.asarray(pcm, dtype=np.uint8), fmt="%d")


def edge_index_of(pcm):
    """Per-graph edge_index [2, E] int64 of H = pcm.T in `H.to_sparse()._indices()` order."""
    H = np.ascontiguousarray(np.array(pcm).T)
    v, c = np.nonzero(H)
    return np.stack([v, c]).astype(np.int64)
