"""Developer script: prints the variable-phase table diagnostics of one CTA (library built with
GD_EXTRA_NVCC=-DGD_VTAB_DEBUG; the kernel prints when B == 4242)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_decode_b200 import codes
from gnn_decode_b200.graph import TannerGraph
from gnn_decode_b200.sampler import sample_syndromes
from bench import load_weights, make_decoder

P10 = [0.01 * k for k in range(1, 11)]
dev = torch.device("cuda", 0)
pcm = codes.rotated_surface_pcm(5)
g = TannerGraph.from_pcm(pcm, dev)
w, _ = load_weights("v2_4")
dec = make_decoder("v2_4", 15, w).to(dev).eval().bind_graph(g)
x, err = sample_syndromes(g, 4242, P10, noise=1, seed=1234)
print("priors of syndrome 0..5:", x[:6, :3].tolist())
print("distinct priors:", torch.unique(x[:, :g.V]).tolist()[:12])
_, logit, _ = dec.decode(x, return_logits=True, return_hard=True)
torch.cuda.synchronize()
m = logit.abs().max().item()
print("max |logit|", m)

def timed(B, reps=10):
    xx, _ = sample_syndromes(g, B, P10, noise=1, seed=1234)
    for _ in range(3):
        dec.decode(xx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dec.decode(xx)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, g.launch_info(dec.gd_model(), B)

for B in (65536, 9472):
    for env in ({}, {"GD_NO_VTAB": "1"}, {"GD_VTAB_K": "12", "GD_VTAB_N": "512"}, {"GD_VTAB_K": "10", "GD_VTAB_N": "512"}, {"GD_VTAB_K": "10", "GD_VTAB_N": "640"}, {"GD_VTAB_K": "12", "GD_VTAB_N": "384"}):
        for k in ("GD_NO_VTAB", "GD_VTAB_K", "GD_VTAB_N"):
            os.environ.pop(k, None)
        os.environ.update(env)
        ms, info = timed(B)
        print("B=%d %-22s %.3f ms  %.2f M syn/s  %s" % (B, env, ms, B / ms / 1e3, info))
