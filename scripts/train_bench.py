"""Training-step timing (BASELINE config 4): decoder_v2_4 program on the rotated surface code d=7,
B syndromes per GPU, fp32: forward(+stash) + LossFunc + hand-written backward + gradient all-reduce
(NCCL when launched under torchrun) + Adam.  Developer script, not the driver's bench.py."""
import math, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
from gnn_decode_b200 import codes
from gnn_decode_b200.dist import allreduce_flat_grads
from gnn_decode_b200.graph import TannerGraph
from gnn_decode_b200.quantum import decoder_v2_4
from gnn_decode_b200.sampler import sample_syndromes
from gnn_decode_b200.train import FusedTrainer, train_step_grads

rank, world, lr_ = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr_)
dev = torch.device("cuda", lr_)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
d = int(sys.argv[1]) if len(sys.argv) > 1 else 7
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
T = 15
Hz, Hx = codes.rotated_surface_checks(d)
pcm = codes.css_pcm(Hz, Hx)
logical = torch.from_numpy(codes.css_logicals(Hz, Hx)).float().to(dev)
Ht = torch.from_numpy(pcm.T.copy()).float().to(dev)          # the reference's H [V, C]
g = TannerGraph.from_pcm(pcm, dev)
torch.manual_seed(0)
dec = decoder_v2_4.GNNI(T).to(dev).train().bind_graph(g)
opt = torch.optim.Adam(dec.parameters(), 3e-4, weight_decay=1e-9)
x, err = sample_syndromes(g, B, [0.01, 0.03, 0.05, 0.08], noise=1, seed=1, first_sample=rank * B)
y = err.float()

def loss_fn(prob):
    z = (y + prob).t()
    return torch.sin(Ht.t() @ z * (math.pi / 2)).abs().sum() + torch.sin(logical @ z * (math.pi / 2)).abs().sum()

FUSED = os.environ.get("GD_TRAIN_AUTOGRAD") is None     # default: fused forward -> sparse loss kernel -> backward
logical_u8 = codes.css_logicals(Hz, Hx)


P2P = None
if world > 1 and os.environ.get("GD_NCCL_ALLREDUCE") is None:
    try:
        from gnn_decode_b200.dist import P2PAllReduce
        P2P = P2PAllReduce(sum(p.numel() for p in dec._gd_params()), dev)   # peer-memory kernel (csrc/gd_p2p.cu)
    except Exception as e:                                                    # no symmetric memory: NCCL
        if rank == 0:
            print("P2PAllReduce unavailable (%s): using NCCL" % e)


TRAINER = None
if FUSED and os.environ.get("GD_TORCH_ADAM") is None and (world == 1 or P2P is not None):
    TRAINER = FusedTrainer(dec, g, logical_u8, lr=3e-4, weight_decay=1e-9, p2p=P2P)   # Adam as a kernel too (gd_adam_step / fused into the all-reduce)


def step_trainer(timers=None):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    loss = TRAINER.step(x, err)
    ev[1].record()
    if timers is not None:
        torch.cuda.synchronize()
        timers["fwd+loss+bwd+allreduce+adam"] = timers.get("fwd+loss+bwd+allreduce+adam", 0.0) + ev[0].elapsed_time(ev[1])
    return loss


def step_fused(timers=None):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    loss, _ = train_step_grads(dec, g, x, err, logical_u8, p2p=P2P)
    ev[1].record()
    if P2P is None:
        allreduce_flat_grads(dec.parameters())
    opt.step()
    ev[2].record()
    if timers is not None:
        torch.cuda.synchronize()
        for i, k in enumerate(("fwd+loss+bwd", "allreduce+adam")):
            timers[k] = timers.get(k, 0.0) + ev[i].elapsed_time(ev[i + 1])
    return loss


def step(timers=None):
    if TRAINER is not None:
        return step_trainer(timers)
    if FUSED:
        return step_fused(timers)
    opt.zero_grad(set_to_none=False)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    prob = dec.decode(x)
    ev[1].record()
    loss = loss_fn(prob)
    ev[2].record()
    loss.backward()
    ev[3].record()
    allreduce_flat_grads(dec.parameters())
    opt.step()
    ev[4].record()
    if timers is not None:
        torch.cuda.synchronize()
        for i, k in enumerate(("fwd", "loss", "bwd", "allreduce+adam")):
            timers[k] = timers.get(k, 0.0) + ev[i].elapsed_time(ev[i + 1])
    return loss

for _ in range(3):
    step()
torch.cuda.synchronize()
n = 10
timers = {}
t0 = time.perf_counter()
for _ in range(n):
    l = step(timers)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
if P2P is not None:
    P2P.check()
if rank == 0:
    print("gradient all-reduce:", "peer-memory kernel (gd_p2p_allreduce)" if P2P is not None else ("NCCL" if world > 1 else "none (1 GPU)"),
          "| optimizer:", "Adam kernel (FusedTrainer)" if TRAINER is not None else "torch.optim.Adam")
    print("rotated d=%d V=%d C=%d E=%d  B/GPU=%d x %d GPU  T=%d: %.3f ms/step  %.1f steps/s  %.3f M syndromes/s  loss %.2f" %
          (d, g.V, g.C, g.E, B, world, T, dt * 1e3, 1 / dt, world * B / dt / 1e6, l.item()))
    print("  per-step ms:", {k: round(v / n, 3) for k, v in timers.items()})
if world > 1:
    dist.destroy_process_group()
