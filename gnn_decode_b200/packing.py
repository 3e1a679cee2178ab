"""Packed syndrome format of gd_decode_packed_* (include/gnn_decode.h): per syndrome ONE prior value + C check-sign bits in,
V hard-decision bits out -- what the reference's x = [log((1-p)/p)] * V | (-1)^syndrome (gen_syn, quantum/error_generate.py:
258, 270-276) actually carries.  Pure torch helpers (any device); the kernels never need them."""
import torch


def pack_x(x, V):
    """x [B, V+C] with one prior per row and +-1 check inputs -> (prior [B] fp32, synd_bits [B, ceil(C/32)] int32)."""
    B, N = x.shape
    Cn = N - V
    nw = (Cn + 31) // 32
    fired = (x[:, V:] < 0).to(torch.int64)
    pad = torch.zeros((B, nw * 32 - Cn), dtype=torch.int64, device=x.device)
    f = torch.cat([fired, pad], 1).view(B, nw, 32)
    words = (f << torch.arange(32, device=x.device, dtype=torch.int64)).sum(-1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
    return x[:, 0].to(torch.float32).contiguous(), words.contiguous()


def unpack_bits(bits, n):
    """bits [B, ceil(n/32)] int32 -> [B, n] uint8."""
    B, nw = bits.shape
    w = bits.to(torch.int64) & 0xFFFFFFFF
    out = (w.unsqueeze(-1) >> torch.arange(32, device=bits.device, dtype=torch.int64)) & 1
    return out.view(B, nw * 32)[:, :n].to(torch.uint8)
