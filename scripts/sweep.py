"""Developer sweep of the (tile, R) launch geometry via the GD_TILE / GD_R overrides."""
import os, sys, itertools
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_decode_b200 import codes
from gnn_decode_b200.graph import TannerGraph
from gnn_decode_b200.quantum import decoder_v2_4
from scripts.dev_bench import time_decode

dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "rot5"
pcm, B = {"rot5": (codes.rotated_surface_pcm(5), 65536), "tor5": (codes.toric_pcm(5), 65536),
          "rot11": (codes.rotated_surface_pcm(11), 16384), "tor11": (codes.toric_pcm(11), 8192)}[which]
g = TannerGraph.from_pcm(pcm, dev)
dec = decoder_v2_4.GNNI(15).to(dev).eval()
units = g.E * (15 * 256 + 128)
os.environ.pop("GD_TILE", None); os.environ.pop("GD_R", None); os.environ.pop("GD_EB", None)
ms = time_decode(dec, g, B, iters=3, warm=1)
print("auto   %s  %.3f ms  %.3f Msyn/s  %.3f Tunits/s" % (g.launch_info(dec.gd_model(), B), ms, B / ms / 1e3, B * units / ms / 1e9), flush=True)
tiles = [int(t) for t in sys.argv[2].split(",")] if len(sys.argv) > 2 else [32, 56, 64, 80, 96, 112, 128, 160, 192]
Rs = [int(t) for t in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 4, 5, 8, 10, 16, 20]
CPS = [int(t) for t in sys.argv[4].split(",")] if len(sys.argv) > 4 else [4, 2]
for t, r, c in itertools.product(tiles, Rs, CPS):
    if t * r > 1024 or (t * r) % 32:
        continue
    os.environ["GD_TILE"], os.environ["GD_R"], os.environ["GD_EB"] = str(t), str(r), str(c)
    try:
        info = g.launch_info(dec.gd_model(), B)
        if info["tile"] != t or info["threads"] != t * r :
            continue
        ms = time_decode(dec, g, B, iters=3, warm=1)
    except Exception as e:
        print("tile", t, "R", r, "failed", e); continue
    print("EB %d tile %4d R %3d thr %4d  %.3f ms  %.3f Msyn/s  %.3f Tunits/s" % (c, t, r, t * r, ms, B / ms / 1e3, B * units / ms / 1e9), flush=True)
