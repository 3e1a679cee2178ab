"""Optimizer step: oracle/adam.py pinned against torch.optim.Adam itself on the CPU; gd_adam_step and
FusedTrainer against the oracle / against train_step_grads + torch.optim.Adam on the GPU."""
import ctypes as ct

import numpy as np
import pytest
import torch

from oracle.adam import AdamOracle


def _torch_adam_run(w0, grads, **kw):
    p = torch.nn.Parameter(torch.from_numpy(w0.copy()))
    opt = torch.optim.Adam([p], **kw)
    out = []
    for g in grads:
        p.grad = torch.from_numpy(g.copy())
        opt.step()
        out.append(p.detach().numpy().copy())
    return out


@pytest.mark.parametrize("wd", [0.0, 1e-9, 1e-2])
def test_oracle_matches_torch_adam(wd):
    rng = np.random.default_rng(5)
    n = 1283
    w0 = rng.standard_normal(n).astype(np.float32)
    grads = [(rng.standard_normal(n) * 10.0 ** rng.integers(-3, 3)).astype(np.float32) for _ in range(25)]
    ref = _torch_adam_run(w0, grads, lr=3e-4, weight_decay=wd)
    o = AdamOracle(n, lr=3e-4, weight_decay=wd)
    w = w0
    for k, g in enumerate(grads):
        w = o.step(w, g)
        np.testing.assert_allclose(w, ref[k], rtol=2e-6, atol=2e-7)


def test_oracle_fp64_matches_torch_adam_fp64():
    rng = np.random.default_rng(6)
    w0 = rng.standard_normal(100)
    grads = [rng.standard_normal(100) for _ in range(10)]
    ref = _torch_adam_run(w0, grads, lr=3e-4, weight_decay=1e-9)
    o = AdamOracle(100, dtype=np.float64)
    w = w0
    for k, g in enumerate(grads):
        w = o.step(w, g)
        np.testing.assert_allclose(w, ref[k], rtol=1e-12, atol=1e-14)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 179, 1283, 100000])
def test_gd_adam_step_matches_oracle(n):
    from gnn_decode_b200 import _cabi
    lib = _cabi.lib()
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(n)
    w0 = rng.standard_normal(n).astype(np.float32)
    w = torch.from_numpy(w0).to(dev)
    m = torch.zeros(n, device=dev)
    v = torch.zeros(n, device=dev)
    o = AdamOracle(n)
    adam = _cabi.GdAdam(3e-4, 0.9, 0.999, 1e-8, 1e-9, 0)
    wo = w0
    st = ct.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for k in range(12):
        g = (rng.standard_normal(n) * 10.0 ** rng.integers(-3, 3)).astype(np.float32)
        gd = torch.from_numpy(g).to(dev)
        adam.step += 1
        _cabi.check(lib.gd_adam_step(ct.byref(adam), ct.c_void_p(w.data_ptr()), ct.c_void_p(gd.data_ptr()),
                                     ct.c_void_p(m.data_ptr()), ct.c_void_p(v.data_ptr()), n, 1.0, st))
        wo = o.step(wo, g)
        np.testing.assert_allclose(w.cpu().numpy(), wo, rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(m.cpu().numpy(), o.m, rtol=1e-5, atol=1e-6 * float(np.abs(o.m).max()))   # cancellation near zero
    np.testing.assert_allclose(v.cpu().numpy(), o.v, rtol=1e-5, atol=1e-12)


@pytest.mark.gpu
def test_gd_adam_step_rejects_bad_arguments():
    from gnn_decode_b200 import _cabi
    lib = _cabi.lib()
    t = torch.zeros(4, device="cuda")
    p = ct.c_void_p(t.data_ptr())
    bad = _cabi.GdAdam(3e-4, 0.9, 0.999, 1e-8, 0.0, 0)              # step 0
    assert lib.gd_adam_step(ct.byref(bad), p, p, p, p, 4, 1.0, None) == _cabi.GD_ERR_INVALID
    bad = _cabi.GdAdam(3e-4, 1.0, 0.999, 1e-8, 0.0, 1)              # beta1 = 1
    assert lib.gd_adam_step(ct.byref(bad), p, p, p, p, 4, 1.0, None) == _cabi.GD_ERR_INVALID
    ok = _cabi.GdAdam(3e-4, 0.9, 0.999, 1e-8, 0.0, 1)
    assert lib.gd_adam_step(ct.byref(ok), p, p, None, p, 4, 1.0, None) == _cabi.GD_ERR_INVALID
    assert lib.gd_adam_step(ct.byref(ok), p, p, p, p, 0, 1.0, None) == _cabi.GD_ERR_INVALID


@pytest.mark.gpu
def test_fused_trainer_matches_grads_plus_torch_adam():
    """FusedTrainer.step == train_step_grads + torch.optim.Adam on the same batches (the reference's loop body,
    quantum/decoder_v2_4.py:325-341), step for step."""
    from gnn_decode_b200 import codes
    from gnn_decode_b200.graph import TannerGraph
    from gnn_decode_b200.quantum import decoder_v2_4
    from gnn_decode_b200.sampler import sample_syndromes
    from gnn_decode_b200.train import FusedTrainer, train_step_grads
    dev = torch.device("cuda", 0)
    Hz, Hx = codes.rotated_surface_checks(5)
    pcm = codes.css_pcm(Hz, Hx)
    logical = codes.css_logicals(Hz, Hx)
    g = TannerGraph.from_pcm(pcm, dev)
    torch.manual_seed(3)
    dec_a = decoder_v2_4.GNNI(6).to(dev).train().bind_graph(g)
    dec_b = decoder_v2_4.GNNI(6).to(dev).train().bind_graph(g)
    dec_b.load_state_dict(dec_a.state_dict())
    opt = torch.optim.Adam(dec_a.parameters(), 3e-4, weight_decay=1e-9)
    tr = FusedTrainer(dec_b, g, logical)
    w_start = tr.w.clone()
    for k in range(4):
        x, err = sample_syndromes(g, 512, [0.02, 0.05, 0.08], noise=1, seed=11 + k)
        loss_a, _ = train_step_grads(dec_a, g, x, err, logical)
        opt.step()
        loss_b = tr.step(x, err)
        assert abs(loss_a.item() - loss_b.item()) <= 1e-5 * abs(loss_a.item())
    assert tr.step_count == 4
    tr.sync_module()
    for (ka, pa), (kb, pb) in zip(dec_a.state_dict().items(), dec_b.state_dict().items()):
        assert ka == kb and pa.dtype == pb.dtype
        # dec_a steps fp64 parameters with fp32 gradients, the trainer fp32 master weights: 4 steps of lr 3e-4
        np.testing.assert_allclose(pb.cpu().numpy(), pa.cpu().numpy(), rtol=0, atol=2e-6)
    assert (tr.w - w_start).abs().max().item() > 5e-4            # and the weights really moved (4 steps of ~lr each)
