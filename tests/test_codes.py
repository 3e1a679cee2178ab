"""CPU suite, part 2: code constructions (graph/index construction must be bit-exact)."""
import numpy as np
import pytest

from gnn_decode_b200 import codes


@pytest.mark.parametrize("L", [3, 4, 5, 6, 7])
def test_toric_pcm_equals_reference_generate_PCM(codes_npz, L):
    assert np.array_equal(codes.toric_pcm(L), codes_npz["toric_H_L%d" % L])


def test_bch_63_45_equals_reference_file(codes_npz):
    assert np.array_equal(codes.bch_63_45_pcm(), codes_npz["bch_63_45_H"])


@pytest.mark.parametrize("L", [4, 5])
def test_toric_logicals_span_same_space_as_reference(codes_npz, L):
    """The reference's only self-check is symplectic_product(H_prep, H).sum() == 0
    (error_generate.py:311).  Here: our CSS logicals and the reference's `logical` rows give the
    same failure verdict on every syndrome-free residual (same span modulo stabilizers)."""
    pcm = codes.toric_pcm(L)
    n, k = 2 * L * L, L * L - 1
    ours = codes.css_logicals(pcm[:k, :n], pcm[k:, n:])
    ref = codes_npz["toric_logical_L%d" % L]
    assert ours.shape == ref.shape == (4, 2 * n)
    assert not ((pcm.astype(int) @ np.zeros(2 * n, int)) % 2).any()
    ker = codes.gf2_nullspace(pcm)              # all residuals with zero syndrome
    rng = np.random.RandomState(0)
    for _ in range(200):
        r = (rng.randint(0, 2, ker.shape[0]) @ ker.astype(int)) % 2
        assert ((ours.astype(int) @ r) % 2).any() == ((ref.astype(int) @ r) % 2).any()


@pytest.mark.parametrize("d,E", [(3, 24), (5, 80), (7, 168), (11, 440)])
def test_rotated_surface(d, E):
    Hz, Hx = codes.rotated_surface_checks(d)
    assert not ((Hz.astype(int) @ Hx.T.astype(int)) % 2).any()
    pcm = codes.rotated_surface_pcm(d)
    assert pcm.shape == (d * d - 1, 2 * d * d) and pcm.sum() == E
    assert codes.gf2_rank(Hz) + codes.gf2_rank(Hx) == d * d - 1          # k = 1
    lg = codes.css_logicals(Hz, Hx)
    assert lg.shape == (2, 2 * d * d)
    assert min(lg[0].sum(), lg[1].sum()) >= d                              # weight >= distance


def test_hgp_1600_64():
    pcm = codes.hgp_pcm()
    assert pcm.shape == (1536, 3200) and pcm.sum() == 10752
    n = 1600
    Hz, Hx = pcm[:768, :n], pcm[768:, n:]
    assert not ((Hz.astype(int) @ Hx.T.astype(int)) % 2).any()
    assert n - codes.gf2_rank(Hz) - codes.gf2_rank(Hx) == 64
    assert np.array_equal(pcm, codes.hgp_pcm())                            # seeded: reproducible


def test_ldpc_toy_and_edge_index():
    pcm = codes.ldpc_toy_pcm()
    ei = codes.edge_index_of(pcm)
    assert ei.shape == (2, 16)
    assert np.all(np.diff(ei[0]) >= 0)       # sorted by variable


def test_pcm_text_round_trip(tmp_path, codes_npz):
    """The reference's .txt matrix format (classical/BCH(63,45).txt, np.loadtxt at CGNNI.py:181)."""
    pcm = codes.bch_63_45_pcm()
    f = tmp_path / "H.txt"
    codes.write_pcm_txt(str(f), pcm)
    assert np.array_equal(codes.read_pcm_txt(str(f)), pcm)
    assert np.array_equal(codes.read_pcm_txt(str(f)), codes_npz["bch_63_45_H"])
    assert np.array_equal(codes.read_pcm_txt(str(f), transpose=True), pcm.T)


def test_hgp_1600_64_logicals():
    """Logical operators of the [[1600,64]] hypergraph-product code (the input of gd_eval_failures for BASELINE config 5):
    2k = 128 rows that commute with every check, are independent of the stabilizers, and pair up symplectically with full
    rank.  The reference's H_Prep (error_generate.py:145-248) is O(n^3) Python; the bit-packed incremental elimination
    here takes about a second."""
    import time
    pcm = codes.hgp_pcm()
    n = 1600
    Hz, Hx = pcm[:768, :n], pcm[768:, n:]
    t0 = time.time()
    lg = codes.css_logicals(Hz, Hx)
    assert time.time() - t0 < 30
    assert lg.shape == (128, 2 * n)
    lz, lx = lg[:64, :n], lg[64:, n:]
    assert not lg[:64, n:].any() and not lg[64:, :n].any()
    assert not ((Hx.astype(int) @ lz.T.astype(int)) % 2).any()             # logical Z commutes with the X checks
    assert not ((Hz.astype(int) @ lx.T.astype(int)) % 2).any()
    assert codes.gf2_rank(np.vstack([Hz, lz])) == codes.gf2_rank(Hz) + 64  # outside the stabilizer group
    assert codes.gf2_rank(np.vstack([Hx, lx])) == codes.gf2_rank(Hx) + 64
    assert codes.gf2_rank((lz.astype(int) @ lx.T.astype(int)) % 2) == 64   # 64 anticommuting pairs


def test_complete_basis_matches_rank_definition():
    """_complete_basis keeps exactly the rows that raise the rank of [sub; rows kept so far], in row order."""
    rng = np.random.RandomState(3)
    for _ in range(25):
        c = rng.randint(2, 70)
        sub = (rng.rand(rng.randint(0, 6), c) < 0.4).astype(np.uint8)
        full = (rng.rand(rng.randint(1, 12), c) < 0.4).astype(np.uint8)
        got = codes._complete_basis(sub, full)
        cur, want = sub.reshape(-1, c), []
        for v in full:
            cand = np.vstack([cur, v[None]])
            if codes.gf2_rank(cand) > (codes.gf2_rank(cur) if cur.size else 0):
                want.append(v)
                cur = cand
        want = np.array(want, dtype=np.uint8).reshape(-1, c)
        assert got.shape == want.shape and (got == want).all()
