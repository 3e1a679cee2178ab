"""One decode of the headline workload (for ncu captures): rotated d=5, depolarizing, B=65536, checkpoint weights."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_decode_b200 import codes
from gnn_decode_b200.graph import TannerGraph
from gnn_decode_b200.quantum import decoder_v2_4
from gnn_decode_b200.sampler import sample_syndromes
dev = torch.device("cuda", 0)
kind = sys.argv[1] if len(sys.argv) > 1 else "rot5"
pcm = {"rot5": lambda: codes.rotated_surface_pcm(5), "rot11": lambda: codes.rotated_surface_pcm(11), "rot7": lambda: codes.rotated_surface_pcm(7), "toric5": lambda: codes.toric_pcm(5), "toric11": lambda: codes.toric_pcm(11)}[kind]()
g = TannerGraph.from_pcm(pcm, dev)
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "v2_4_toricL5_epoch3.npz"))
dec = decoder_v2_4.GNNI(15)
dec.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")})
dec = dec.to(dev).eval()
x, _ = sample_syndromes(g, 65536, [0.01, 0.02, 0.03, 0.04, 0.05, 0.06, 0.07, 0.08, 0.09, 0.1], noise=1 if kind.startswith("rot") else 0, seed=1)
for _ in range(3):
    dec.decode(x, graph=g, return_hard=True)
torch.cuda.synchronize()
print("done")
