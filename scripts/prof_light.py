"""One decode of a light (ReLU / sum-product) program on toric L=5, B=65536 (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_decode_b200 import codes
from gnn_decode_b200.graph import TannerGraph
from gnn_decode_b200.quantum import QGNNI, BP
from gnn_decode_b200.sampler import sample_syndromes
dev = torch.device("cuda", 0)
g = TannerGraph.from_pcm(codes.toric_pcm(5), dev)
which = sys.argv[1] if len(sys.argv) > 1 else "qgnni"
torch.manual_seed(0)
dec = (QGNNI.GNNI(25) if which == "qgnni" else BP.GNNI(10)).to(dev).eval()
x, _ = sample_syndromes(g, 65536, [0.01, 0.03, 0.05, 0.08], noise=0, seed=1)
for _ in range(3):
    dec.decode(x, graph=g, return_hard=True)
torch.cuda.synchronize()
print("done")
