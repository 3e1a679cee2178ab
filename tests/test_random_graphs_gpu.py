"""GPU suite: random irregular Tanner graphs (isolated variables, degree-1 nodes, check degrees beyond the
unrolled 1..8 range) through every program, on the resident and on the streamed kernels, against the oracle.
The reference only ever runs regular codes; the kernels must not depend on that."""
import numpy as np
import pytest
import torch

from gnn_decode_b200 import codes
from gnn_decode_b200.graph import TannerGraph
from oracle import restate

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _random_pcm(rng, C, V, max_row):
    pcm = np.zeros((C, V), dtype=np.uint8)
    for c in range(C):
        d = rng.randint(1, max_row + 1)
        pcm[c, rng.choice(V, size=min(d, V), replace=False)] = 1
    pcm[:, rng.randint(0, V)] = 0            # at least one isolated variable
    if pcm.sum() == 0:
        pcm[0, 0] = 1
    return pcm


def _decoder(program, T, E):
    from gnn_decode_b200.quantum import decoder_v2_4, QGNNI, BP, neural_BP, QGNNNI_ca, decoder_v3_0, decoder_v1_2_2
    from gnn_decode_b200.classical import CGNNI, BP as CBP
    if program == "neural_bp":
        dec = neural_BP.GNNI(T, n_edges=E)
        with torch.no_grad():
            for n_, p_ in dec.named_parameters():
                p_.copy_(torch.full_like(p_, 0.25) if n_ == "alpha" else torch.rand_like(p_) * 0.6 + 0.6)
        return dec
    return {"v2_4": decoder_v2_4.GNNI, "qgnni": QGNNI.GNNI, "bp_quantum": BP.GNNI, "cgnni": CGNNI.GNNI,
            "bp_classical": CBP.GNNI, "gru_ca": QGNNNI_ca.GNNI, "v3_0": decoder_v3_0.GNNI, "v1_2_2": decoder_v1_2_2.GNNI}[program](T)


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("program", ["v2_4", "qgnni", "cgnni", "bp_quantum", "bp_classical", "neural_bp", "gru_ca", "v3_0", "v1_2_2"])
def test_random_graph_matches_oracle(program, seed, gd_opt):
    rng = np.random.RandomState(100 * seed + 7)
    C, V = rng.randint(3, 14), rng.randint(6, 40)
    pcm = _random_pcm(rng, C, V, max_row=(12 if seed == 2 else 6))
    ei = torch.from_numpy(codes.edge_index_of(pcm))
    E = ei.size(1)
    g = TannerGraph.from_pcm(pcm, DEV)
    B, T = 70 + 13 * seed, 4
    torch.manual_seed(seed)
    x = torch.randn(B, V + C, dtype=torch.float64) * 1.5 + 1.0
    if program not in ("cgnni", "bp_classical"):
        x[:, V:] = torch.sign(torch.randn(B, C, dtype=torch.float64))
    else:
        x[:, V:] = 0.0
    dec = _decoder(program, T, E).to(DEV).eval()
    if program == "v3_0":                                      # default GRUCell(1,1) / 10-unit MLPs barely move anything: make the messages matter
        with torch.no_grad():
            for n_, p_ in dec.named_parameters():
                if "rnn" in n_ or n_.endswith(".2.weight"):
                    p_.mul_(3.0)
    w = {k: v.detach().cpu() for k, v in dec.state_dict().items()}
    ref_all = restate.decode(program, ei, V, C, x, w, T=T, dtype=torch.float64)
    ref = ref_all["logit"]
    bp = "bp" in program or program == "v1_2_2"
    rtol = 2e-3 if bp else 1e-4
    modes = ["resident"] + (["streamed"] if program not in ("neural_bp", "gru_ca", "v3_0", "v1_2_2") else [])
    if program == "v3_0":                                      # the second read-out, at the check nodes
        _, _, _, logit_c = dec.decode_aux(x.to(DEV), graph=g, return_logits=True)
        rc = ref_all["logit_chk"]
        assert bool(((logit_c.double().cpu() - rc).abs() <= 1e-4 * rc.abs().clamp_min(1.0)).all())
    if program == "v1_2_2":                                    # every iteration's read-out
        _, la = dec.decode_all(x.to(DEV), graph=g, return_logits=True)
        ra = ref_all["all_logit"]
        ok_all = (la.double().cpu() - ra).abs() <= 2e-3 * ra.abs().clamp_min(1.0)
        assert bool((ok_all | (ra.abs() > 30)).all()), (la.double().cpu() - ra).abs().max().item()
    for mode in modes:
        if mode == "streamed":
            gd_opt.set("GD_FORCE_STREAMED")
        _, logit, hard = dec.decode(x.to(DEV), graph=g, return_logits=True, return_hard=True)
        gd_opt.unset("GD_FORCE_STREAMED")
        got = logit.double().cpu()
        ok = (got - ref).abs() <= rtol * ref.abs().clamp_min(1.0)          # conftest.logit_bound
        if bp:   # saturated sum-product messages: compare where the reference itself is not at its clamp
            ok = ok | (ref.abs() > 30)
        assert bool(ok.all()), "%s %s: max abs err %.3g" % (program, mode, (got - ref).abs().max().item())
        decided = ref.abs() > 1e-3
        assert torch.equal(hard.cpu().bool()[decided], (ref < 0)[decided])


@pytest.mark.parametrize("mode", ["v2_4-resident", "qgnni-light", "bp-light", "qgnni-streamed", "bp-streamed", "v2_4-streamed"])
def test_repeated_launches_are_bit_identical(mode, gd_opt):
    """compute-sanitizer is not available on the GPU pool, so data races are hunted the blunt way: 25 launches of
    every kernel family on a ragged batch, concurrently on two streams, must all be bit-identical."""
    from gnn_decode_b200.quantum import decoder_v2_4, QGNNI, BP
    from gnn_decode_b200.sampler import sample_syndromes
    prog, kern = mode.split("-")
    pcm = codes.rotated_surface_pcm(5) if prog == "v2_4" else codes.toric_pcm(4)
    g = TannerGraph.from_pcm(pcm, DEV)
    torch.manual_seed(3)
    dec = {"v2_4": decoder_v2_4.GNNI(6), "qgnni": QGNNI.GNNI(7), "bp": BP.GNNI(7)}[prog].to(DEV).eval()
    B = 5003 if kern != "streamed" else 1203
    x, _ = sample_syndromes(g, B, [0.03, 0.08], noise=1 if prog == "v2_4" else 0, seed=4)
    if kern == "light":
        gd_opt.set("GD_FORCE_LIGHT")
    if kern == "streamed":
        gd_opt.set("GD_FORCE_STREAMED")
    first = dec.decode(x, graph=g, return_logits=True)[1].clone()
    streams = [torch.cuda.Stream(DEV), torch.cuda.Stream(DEV)]
    torch.cuda.synchronize()
    outs = []
    for i in range(25):
        if kern == "streamed":                      # the streamed slab is per graph: one stream at a time (include/gnn_decode.h)
            outs.append(dec.decode(x, graph=g, return_logits=True)[1])
        else:
            with torch.cuda.stream(streams[i & 1]):
                outs.append(dec.decode(x, graph=g, return_logits=True)[1])
    torch.cuda.synchronize()
    for o in outs:
        assert torch.equal(o, first)
