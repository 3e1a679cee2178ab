"""One fused training step (for ncu captures): rotated d=7, B=4096, T=15."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_decode_b200 import codes
from gnn_decode_b200.graph import TannerGraph
from gnn_decode_b200.quantum import decoder_v2_4
from gnn_decode_b200.sampler import sample_syndromes
from gnn_decode_b200.train import train_step_grads
dev = torch.device("cuda", 0)
Hz, Hx = codes.rotated_surface_checks(7)
g = TannerGraph.from_pcm(codes.css_pcm(Hz, Hx), dev)
torch.manual_seed(0)
dec = decoder_v2_4.GNNI(15).to(dev).train().bind_graph(g)
x, err = sample_syndromes(g, int(sys.argv[1]) if len(sys.argv) > 1 else 4096, [0.01, 0.03, 0.05, 0.08], noise=1, seed=1)
for _ in range(3):
    train_step_grads(dec, g, x, err, codes.css_logicals(Hz, Hx))
torch.cuda.synchronize()
print("done")
