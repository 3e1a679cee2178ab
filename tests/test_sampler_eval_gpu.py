"""GPU suite: Philox sampler (bit-exact vs the NumPy oracle, distribution vs the reference's
gen_syn) and the failure counters (exact vs the oracle restatement of neural_BP.py:338-348)."""
import numpy as np
import pytest
import torch

from gnn_decode_b200 import codes
from gnn_decode_b200.evaluate import count_failures
from gnn_decode_b200.graph import TannerGraph
from gnn_decode_b200.sampler import sample_syndromes
from oracle import philox

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("code,noise", [(("toric", 4), 0), (("toric", 5), 0), (("rot", 5), 1), (("rot", 3), 0),
                                        (("hgp", 0), 1)])
def test_sampler_bit_exact_vs_oracle(code, noise):
    pcm = {"toric": codes.toric_pcm, "rot": codes.rotated_surface_pcm}.get(code[0], lambda _: codes.hgp_pcm(6, 8))(code[1])
    g = TannerGraph.from_pcm(pcm, DEV)
    P = [0.01, 0.05, 0.1, 0.2]
    B = 777
    x, err = sample_syndromes(g, B, P, noise=noise, seed=0x1234567890ABCDEF, first_sample=(1 << 33) + 5)
    xo, eo = philox.sample(pcm, B, P, noise=noise, seed=0x1234567890ABCDEF, first_sample=(1 << 33) + 5)
    assert np.array_equal(err.cpu().numpy(), eo)
    assert np.array_equal(x.cpu().numpy()[:, g.V:], xo[:, g.V:])                 # syndromes bit-exact
    assert np.allclose(x.cpu().numpy()[:, :g.V], xo[:, :g.V], rtol=1e-6)          # prior: log in double, cast
    # shards are slices of the full draw
    x2, e2 = sample_syndromes(g, 100, P, noise=noise, seed=0x1234567890ABCDEF, first_sample=(1 << 33) + 5 + 300)
    assert torch.equal(e2, err[300:400]) and torch.equal(x2, x[300:400])


@pytest.mark.parametrize("bit", [0, 1])
def test_awgn_sampler_matches_oracle_and_reference_statistics(bit):
    """Classical input (classical/CGNNI.py:125-147,159): x = [2y/sigma^2 | 0], y = BPSK + AWGN.  The Box-Muller
    draw is pinned to the float64 oracle; mean / variance of the LLR match the channel the reference builds."""
    pcm = codes.bch_63_45_pcm()
    g = TannerGraph.from_pcm(pcm, DEV)
    snr = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0]
    B = 4096
    x, cw = sample_syndromes(g, B, snr, noise=2 + bit, seed=77, first_sample=123)
    xo, cwo = philox.sample_awgn(g.V, g.C, B, snr, codeword_bit=bit, seed=77, first_sample=123)
    x = x.cpu().numpy()
    assert np.array_equal(cw.cpu().numpy(), cwo)
    assert np.array_equal(x[:, g.V:], np.zeros((B, g.C), np.float32))               # check slots zero-filled (:159)
    assert np.allclose(x[:, :g.V], xo[:, :g.V], rtol=2e-5, atol=2e-4)
    x1, _ = sample_syndromes(g, B, [4.0], noise=2 + bit, seed=5)                     # one SNR: LLR ~ N(2 tx / s^2, 4 / s^2)
    s2 = 1.0 / 10 ** 0.4
    llr = x1.cpu().numpy()[:, :g.V].astype(np.float64)
    tx = 1.0 - 2.0 * bit
    n = llr.size
    assert abs(llr.mean() - 2 * tx / s2) < 5 * (2 / np.sqrt(s2)) / np.sqrt(n)
    assert abs(llr.var() - 4 / s2) < 0.02 * 4 / s2
    x2, _ = sample_syndromes(g, 100, snr, noise=2 + bit, seed=77, first_sample=123 + 300)
    assert np.array_equal(x2.cpu().numpy(), x[300:400])                               # shards are slices


def test_sampler_distribution_matches_gen_syn_layout(codes_npz):
    """Layout of the reference's gen_syn output (fixture produced by the reference, seeded) and its
    flip statistics: every slot flips independently w.p. p, prior = log((1-p)/p), syndrome signs."""
    L = 4
    pcm = codes.toric_pcm(L)
    g = TannerGraph.from_pcm(pcm, DEV)
    xr, yr = codes_npz["gen_syn_L4_x"], codes_npz["gen_syn_L4_y"]
    assert xr.shape[1] == g.N and yr.shape[1] == g.V
    syn = (yr.astype(np.int64) @ pcm.T.astype(np.int64)) % 2                         # reference layout check
    assert np.array_equal(xr[:, g.V:], 1.0 - 2.0 * syn)
    B, p = 200000, 0.07
    x, err = sample_syndromes(g, B, [p], noise=0, seed=7)
    e = err.float()
    rate = e.mean().item()
    assert abs(rate - p) < 5 * np.sqrt(p * (1 - p) / (B * g.V))
    per_slot = e.mean(0)
    assert (per_slot - p).abs().max().item() < 6 * np.sqrt(p * (1 - p) / B)
    c = torch.corrcoef(e[:, :16].t())                                                # independence of slots
    assert (c - torch.eye(16, device=c.device)).abs().max().item() < 0.02
    assert np.allclose(x[:, :g.V].cpu().numpy(), np.log((1 - p) / p), rtol=1e-6)
    # depolarizing marginals: X, Y, Z each p/3
    gd5 = TannerGraph.from_pcm(codes.rotated_surface_pcm(5), DEV)
    _, err = sample_syndromes(gd5, B, [0.09], noise=1, seed=11)
    n = gd5.V // 2
    ex, ez = err[:, :n].bool(), err[:, n:].bool()
    for mask in (ex & ~ez, ex & ez, ~ex & ez):
        assert abs(mask.float().mean().item() - 0.03) < 5 * np.sqrt(0.03 * 0.97 / (B * n))


def test_failure_counters_exact():
    L = 4
    pcm = codes.toric_pcm(L)
    n, k = 2 * L * L, L * L - 1
    logical = codes.css_logicals(pcm[:k, :n], pcm[k:, n:])
    g = TannerGraph.from_pcm(pcm, DEV)
    B = 5000
    _, err = sample_syndromes(g, B, [0.03], seed=3)
    rng = np.random.RandomState(0)
    ker = codes.gf2_nullspace(pcm)
    hard = err.cpu().numpy().copy()
    # a third: perfect; a third: differ by a random syndrome-free residual; a third: random bits
    hard[B // 3:2 * B // 3] ^= ((rng.randint(0, 2, (2 * B // 3 - B // 3, ker.shape[0])) @ ker) % 2).astype(np.uint8)
    hard[2 * B // 3:] = rng.randint(0, 2, hard[2 * B // 3:].shape)
    want = philox.count_failures(pcm, logical, err.cpu().numpy(), hard)
    got = count_failures(g, err, torch.from_numpy(hard).to(DEV), logical)
    assert tuple(got.tolist()) == want and want[1] > 0 and want[0] > 0
    got2 = count_failures(g, err, err.clone(), logical)
    assert got2.tolist() == [0, 0, 0]


def test_logical_error_rate_matches_reference_statistics(codes_npz):
    """End-to-end sampler -> fused decoder -> failure counters, against (a) the oracle decoder on the
    SAME samples (failure counts must be equal: hard decisions are bit-exact) and (b) the loose
    known answer of the shimmed reference quantum/BP.py (L=4, Nc=10, p=0.01): 99/4096 = 2.4 %
    failures -- statistically indistinguishable within a binomial CI."""
    from gnn_decode_b200.quantum import BP
    from oracle import restate
    L = 4
    pcm = codes.toric_pcm(L)
    logical = codes_npz["toric_logical_L4"]                      # the reference's own `logical`
    g = TannerGraph.from_pcm(pcm, DEV)
    dec = BP.GNNI(10).to(DEV).eval().bind_graph(g)
    B = 40960
    x, err = sample_syndromes(g, B, [0.01], noise=0, seed=2024)
    prob, hard = dec.decode(x, return_hard=True)
    got = count_failures(g, err, hard, logical).tolist()
    rate = got[2] / B
    p_ref = 99 / 4096
    sigma = np.sqrt(p_ref * (1 - p_ref) * (1 / B + 1 / 4096))
    assert abs(rate - p_ref) < 4 * sigma, (rate, p_ref)
    # same samples through the fp64 oracle decoder
    n = 4096
    ei = torch.from_numpy(codes.edge_index_of(pcm))
    ref = restate.decode("bp_quantum", ei, g.V, g.C, x[:n].cpu().double(), {}, T=10)
    want = philox.count_failures(pcm, logical, err[:n].cpu().numpy(), (ref["prob"] > 0.5).numpy().astype(np.uint8))
    got_n = count_failures(g, err[:n], hard[:n], logical).tolist()
    assert abs(got_n[2] - want[2]) <= 1, (got_n, want)            # a tie-break flip at most
