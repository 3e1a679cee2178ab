"""GPU suite at BASELINE.json's FULL sizes, where the oracle is too slow to run on everything:
size-independent properties (the batch is block-diagonal, so a syndrome's result may not depend
on what it is batched with; decoding is deterministic; an error-free syndrome decodes to "no
flip"; a sample -> decode -> count loop is consistent) plus an oracle check on a random subset."""
import numpy as np
import pytest
import torch

from gnn_decode_b200 import codes
from gnn_decode_b200.evaluate import count_failures
from gnn_decode_b200.graph import TannerGraph
from gnn_decode_b200.quantum import BP, QGNNI, decoder_v2_4
from gnn_decode_b200.sampler import sample_syndromes
from conftest import logit_worst
from oracle import restate

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-4          # BASELINE.json north_star: soft logits within 1e-4 relative in fp32


def _v2_4(T):
    z = np.load(__file__.rsplit("/", 1)[0] + "/golden/v2_4_toricL5_epoch3.npz")
    w = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}
    dec = decoder_v2_4.GNNI(T)
    dec.load_state_dict(w)
    return dec.to(DEV).eval(), w


def _subset_vs_oracle(program, pcm, x, logit, weights, T, idx, rtol):
    ei = torch.from_numpy(codes.edge_index_of(pcm))
    Cn, V = pcm.shape
    ref = restate.decode(program, ei, V, Cn, x[idx].cpu().double(), weights, T=T, dtype=torch.float64)["logit"]
    worst, max_err = logit_worst(logit[idx], ref, rtol)
    assert worst <= 1.0, "logits: %.3g x the bar (max abs err %.3g)" % (worst, max_err)


@pytest.mark.parametrize("code,B", [(("rotated", 5), 65536), (("toric", 11), 65536), (("rotated", 11), 65536)])
def test_v2_4_full_batch_properties(code, B):
    """configs[1] (rotated d=5, depolarizing, B = 65536) and configs[2] (d = 11, B = 65536 per GPU)."""
    pcm = codes.rotated_surface_pcm(code[1]) if code[0] == "rotated" else codes.toric_pcm(code[1])
    g = TannerGraph.from_pcm(pcm, DEV)
    dec, w = _v2_4(15)
    x, err = sample_syndromes(g, B, [0.01, 0.03, 0.05, 0.08], noise=1 if code[0] == "rotated" else 0, seed=11)
    prob, logit, hard = dec.decode(x, graph=g, return_logits=True, return_hard=True)
    prob2 = dec.decode(x, graph=g)
    assert torch.equal(prob, prob2)                                          # deterministic
    assert torch.equal(hard.bool(), prob > 0.5)
    # batch-composition invariance: any slice decoded alone is bit-identical (different tile geometry)
    for lo, n in ((0, 8), (12345, 777), (B - 1001, 1001)):
        assert torch.equal(dec.decode(x[lo:lo + n].contiguous(), graph=g), prob[lo:lo + n])
    # rank shards are slices of the single-GPU draw and decode (no collective)
    xs, _ = sample_syndromes(g, 4096, [0.01, 0.03, 0.05, 0.08], noise=1 if code[0] == "rotated" else 0, seed=11,
                             first_sample=B // 2)
    assert torch.equal(xs, x[B // 2:B // 2 + 4096])
    assert torch.equal(dec.decode(xs, graph=g), prob[B // 2:B // 2 + 4096])
    idx = torch.from_numpy(np.random.RandomState(0).choice(B, 48, replace=False))
    _subset_vs_oracle("v2_4", pcm, x, logit, w, 15, idx.to(DEV), RTOL)


@pytest.mark.parametrize("code", ["ldpc", "bch"])
def test_cgnni_config0_full_batch_vs_oracle(code):
    """configs[0]: classical/CGNNI.GNNI(25) with the shipped checkpoint on the smallest bundled code (H_LDPC 4x8) and on
    BCH(63,45), B = 1024 channel LLRs from the AWGN mode of the Philox sampler (classical/CGNNI.py:125-159): the WHOLE batch
    against the oracle, hard decisions bit-exact away from ties, slices bit-identical."""
    from gnn_decode_b200.classical import CGNNI
    pcm = codes.ldpc_toy_pcm() if code == "ldpc" else codes.bch_63_45_pcm()
    g = TannerGraph.from_pcm(pcm, DEV)
    z = np.load(__file__.rsplit("/", 1)[0] + "/golden/cgnni_ldpc_epoch18.npz")
    w = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}
    dec = CGNNI.GNNI(25)
    dec.load_state_dict(w)
    dec = dec.to(DEV).eval()
    B = 1024
    x, cw = sample_syndromes(g, B, [1.0, 2.0, 3.0, 4.0, 5.0, 6.0], noise=2 if code == "ldpc" else 3, seed=1234)
    assert bool((x[:, g.V:] == 0).all()) and bool((cw == (0 if code == "ldpc" else 1)).all())
    prob, logit, hard = dec.decode(x, graph=g, return_logits=True, return_hard=True)
    assert torch.equal(prob, dec.decode(x, graph=g))
    for lo, n in ((0, 8), (100, 333)):
        assert torch.equal(dec.decode(x[lo:lo + n].contiguous(), graph=g), prob[lo:lo + n])
    ei = torch.from_numpy(codes.edge_index_of(pcm))
    Cn, V = pcm.shape
    ref = restate.decode("cgnni", ei, V, Cn, x.cpu(), w, T=25, dtype=torch.float64)
    worst, max_err = logit_worst(logit, ref["logit"], RTOL)
    assert worst <= 1.0, "logits: %.3g x the bar (max abs err %.3g)" % (worst, max_err)
    decided = ref["logit"].abs() > 1e-3
    assert torch.equal(hard.cpu().bool()[decided], (ref["logit"] < 0)[decided])
    assert (prob.cpu().double() - ref["prob"]).abs().max().item() < 1e-5      # includes the reference's clamp to [1e-7, 1 - 1e-7]


def test_ler_matches_the_oracle_on_20k_syndromes():
    """Hard decisions and failure counters of the GPU decoder vs the fp64 oracle on 20 480 Philox syndromes of the headline
    workload (the 2^18-sample run of the same script is recorded under profiles/)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import ler_vs_oracle
    r = ler_vs_oracle.run(20480)
    assert r["worst_over_bar"] <= 1.0
    assert r["hard_mismatches"] == 0 or r["max_abs_ref_logit_at_mismatch"] < 1e-3        # ties only
    assert all(abs(a - b) <= r["hard_mismatches"] for a, b in zip(r["failures_gpu"], r["failures_oracle"]))


def test_hgp_streamed_full_batch_properties():
    """configs[4]: hypergraph-product [[1600,64]] (V = 3200, E = 10752), B = 16384, many iterations,
    streamed global-memory path: determinism, batch-composition invariance, oracle subset, and the
    sum-product decoder leaves error-free syndromes alone and corrects all weight-1 errors."""
    pcm = codes.hgp_pcm()
    g = TannerGraph.from_pcm(pcm, DEV)
    B = 16384
    x, err = sample_syndromes(g, B, [0.005, 0.01, 0.02], noise=1, seed=3)
    torch.manual_seed(0)
    q = QGNNI.GNNI(20).to(DEV).eval()
    assert g.launch_info(q.gd_model(), B)["resident"] == 0
    prob, logit = q.decode(x, graph=g, return_logits=True)
    assert torch.equal(prob, q.decode(x, graph=g))
    for lo, n in ((0, 40), (5000, 333)):                                     # other tile sizes / ragged tiles
        assert torch.equal(q.decode(x[lo:lo + n].contiguous(), graph=g), prob[lo:lo + n])
    idx = torch.tensor([0, 1, 4097, 9999, B - 1], device=DEV)
    _subset_vs_oracle("qgnni", pcm, x, logit, {k: v.cpu() for k, v in q.state_dict().items()}, 20, idx, RTOL)

    bp = BP.GNNI(30).to(DEV).eval()
    prob, logit, hard = bp.decode(x, graph=g, return_logits=True, return_hard=True)
    assert torch.equal(hard, bp.decode(x, graph=g, return_hard=True)[1])
    _subset_vs_oracle("bp_quantum", pcm, x, logit, {}, 30, idx, 2e-3)         # BP bar: tests/test_parity_gpu.py
    clean = (err.sum(1) == 0)
    assert int(clean.sum()) > 0 and int(hard[clean].sum()) == 0              # no error -> no correction
    w1 = (err.sum(1) == 1)
    cnt = count_failures(g, err[w1].contiguous(), hard[w1].contiguous())
    assert int(w1.sum()) > 0 and int(cnt[0]) == 0                            # single errors: syndrome always cleared
    total = count_failures(g, err, hard)
    assert int(total[0]) < 0.10 * B                                          # sanity: plain BP well below threshold


def test_variable_phase_tables_and_their_fallbacks(gd_opt):
    """decoder_v2_4's variable-phase MLP is read from per-prior cubic tables when every variable of a syndrome carries the
    same prior.  Every way out of that fast path must give the reference's numbers too:
    (a) uniform priors, few distinct values: tables (vs the oracle, and vs the direct kernel GD_NO_VTAB=1);
    (b) a different prior on every variable: nothing is tabulated -> bit-identical to the direct kernel;
    (c) more distinct per-syndrome priors than table slots: some syndromes tabulated, the rest direct;
    (d) a mix of uniform and non-uniform syndromes in the same tiles;
    and the table look-up may not make a result depend on which syndromes share a tile (batch slices bit-identical)."""
    pcm = codes.rotated_surface_pcm(5)
    g = TannerGraph.from_pcm(pcm, DEV)
    V = g.V
    dec, w = _v2_4(15)
    B = 20000
    info = g.tables_info(dec.gd_model(), B)
    assert info[2] > 0 and info[3] >= 10, info            # the reference's list of 10 error rates fits the table slots
    x, _ = sample_syndromes(g, B, [0.01 * k for k in range(1, 11)], noise=1, seed=5)
    gen = torch.Generator(device=DEV).manual_seed(3)
    cases = {"uniform10": x.clone()}
    xb = x.clone()
    xb[:, :V] = 2.0 + 3.0 * torch.rand(B, V, device=DEV, generator=gen)
    cases["per_variable"] = xb
    xc = x.clone()
    xc[:, :V] = (2.0 + 3.0 * torch.rand(B, 1, device=DEV, generator=gen)).expand(B, V)      # B distinct per-syndrome priors
    cases["many_distinct"] = xc
    xd = x.clone()
    odd = torch.arange(B, device=DEV) % 3 == 1
    xd[odd, 7] += 0.25                                                                       # one variable off: not uniform
    cases["mixed"] = xd
    idx = torch.from_numpy(np.random.RandomState(1).choice(B, 40, replace=False)).to(DEV)
    out = {}
    for name, xx in cases.items():
        prob, logit, hard = dec.decode(xx, graph=g, return_logits=True, return_hard=True)
        out[name] = (prob, logit, hard)
        _subset_vs_oracle("v2_4", pcm, xx, logit, w, 15, idx, RTOL)
        if name in ("uniform10", "mixed"):              # few distinct priors: slices are bit-identical (table content depends on the prior only)
            for lo, n in ((0, 8), (777, 1234), (B - 501, 501)):
                assert torch.equal(dec.decode(xx[lo:lo + n].contiguous(), graph=g), prob[lo:lo + n]), (name, lo)
    gd_opt.set("GD_NO_VTAB")
    assert g.tables_info(dec.gd_model(), B)[2] == 0
    for name, xx in cases.items():
        prob, logit, hard = out[name]
        p2, l2, h2 = dec.decode(xx, graph=g, return_logits=True, return_hard=True)
        if name == "per_variable":
            assert torch.equal(prob, p2) and torch.equal(hard, h2)
        else:
            scale = 1.0 + l2.pow(2).mean().sqrt().item()
            assert (logit - l2).abs().max().item() <= 0.05 * RTOL * scale, name      # tables vs direct: 20x inside the parity bar
            decided = l2.abs() > 1e-3
            assert torch.equal(hard[decided], h2[decided]), name
