import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_decode_b200 import codes
from gnn_decode_b200.graph import TannerGraph
from gnn_decode_b200.quantum import QGNNI, BP
dev = torch.device("cuda", 0)
g = TannerGraph.from_pcm(codes.hgp_pcm(), dev)
which = sys.argv[1] if len(sys.argv) > 1 else "qgnni"
dec = (QGNNI.GNNI(20) if which == "qgnni" else BP.GNNI(20)).to(dev).eval()
B = 16384
x = torch.randn(B, g.N, device=dev); x[:, g.V:] = torch.sign(x[:, g.V:]); x[:, :g.V] = 2.0 + 0.5 * x[:, :g.V]
for _ in range(3):
    dec.decode(x, graph=g)
torch.cuda.synchronize()
print("done")
