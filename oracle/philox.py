"""Oracle (TEST INFRASTRUCTURE ONLY) for the synthetic-input sampler: a NumPy restatement of
Philox4x32-10 (Salmon et al., SC'11; the published Random123 algorithm) and of the sampling rule
csrc/gd_sampler.cu implements on top of it, whose layout/distribution is that of the reference's
gen_syn (quantum/error_generate.py:252-278: p uniformly from the list P per sample, prior
log((1-p)/p) on every slot, independent flips, x = [prior | (-1)^(H^T e mod 2)], y = e).
The reference draws from an unseeded Mersenne Twister, so there is no sample-level golden vector:
this oracle pins the CUDA sampler bit-for-bit, and the reference pins the distribution."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over counters (uint32 arrays); key scalars. Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(a, dtype=np.uint32).copy() for a in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return c0, c1, c2, c3


def u01(w):
    return (w >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def sample(pcm, B, p_list, noise=0, seed=1234, first_sample=0):
    """-> (x [B, V+C] float32, err [B, V] uint8); pcm [C, V]."""
    pcm = np.asarray(pcm, dtype=np.uint8)
    Cn, V = pcm.shape
    p_list = np.asarray(p_list, dtype=np.float32)
    sid = np.uint64(first_sample) + np.arange(B, dtype=np.uint64)
    s_lo, s_hi = (sid & MASK).astype(np.uint32), (sid >> np.uint64(32)).astype(np.uint32)
    k0, k1 = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
    sel = philox4x32_10(s_lo, s_hi, np.uint32(0), np.uint32(0), k0, k1)[0]
    pi = (sel % np.uint32(len(p_list))).astype(np.int64)
    p = p_list[pi]                                               # float32
    n_draw = V if noise == 0 else V // 2
    nblk = (n_draw + 3) // 4
    blk = np.arange(nblk, dtype=np.uint32)
    r = philox4x32_10(s_lo[:, None], s_hi[:, None], blk[None, :], np.uint32(1), k0, k1)
    u = u01(np.stack(r, axis=-1).reshape(B, nblk * 4)[:, :n_draw])   # word q of block j0/4 -> draw j0+q
    if noise == 0:
        err = (u < p[:, None]).astype(np.uint8)
        pm = p.astype(np.float64)
    else:
        t1 = (p / np.float32(3.0)).astype(np.float32)
        t2 = (np.float32(2.0) * p / np.float32(3.0)).astype(np.float32)
        ex = (u < t2[:, None])
        ez = (u >= t1[:, None]) & (u < p[:, None])
        err = np.concatenate([ex, ez], 1).astype(np.uint8)
        pm = 2.0 * p.astype(np.float64) / 3.0
    prior = np.log((1.0 - pm) / pm).astype(np.float32)
    syn = (err.astype(np.int64) @ pcm.T.astype(np.int64)) % 2
    x = np.concatenate([np.repeat(prior[:, None], V, 1), (1.0 - 2.0 * syn).astype(np.float32)], 1)
    return x, err


def sample_awgn(V, Cn, B, snr_db_list, codeword_bit=0, seed=1234, first_sample=0):
    """The classical input (classical/CGNNI.py:125-147,159) as csrc/gd_sampler.cu draws it (noise 2/3):
    sigma = (1/10^(SNR/10))^0.5, y = (1 - 2c) + sigma * N(0,1) (Box-Muller on Philox words),
    x = [2 y / sigma^2 | 0].  float64 arithmetic: pins the CUDA kernel to ~1e-5 relative, not bitwise."""
    snr = np.asarray(snr_db_list, dtype=np.float32)
    sid = np.uint64(first_sample) + np.arange(B, dtype=np.uint64)
    s_lo, s_hi = (sid & MASK).astype(np.uint32), (sid >> np.uint64(32)).astype(np.uint32)
    k0, k1 = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
    sel = philox4x32_10(s_lo, s_hi, np.uint32(0), np.uint32(0), k0, k1)[0]
    pi = (sel % np.uint32(len(snr))).astype(np.int64)
    sigma = np.sqrt(1.0 / 10.0 ** (snr.astype(np.float64) / 10.0)).astype(np.float32).astype(np.float64)[pi]
    inv2 = (2.0 / np.sqrt(1.0 / 10.0 ** (snr.astype(np.float64) / 10.0)) ** 2).astype(np.float32).astype(np.float64)[pi]
    nblk = (V + 3) // 4
    blk = np.arange(nblk, dtype=np.uint32)
    r = philox4x32_10(s_lo[:, None], s_hi[:, None], blk[None, :], np.uint32(1), k0, k1)
    ua = lambda w: ((w >> np.uint32(8)).astype(np.float64) + 1.0) / 16777216.0
    ub = lambda w: (w >> np.uint32(8)).astype(np.float64) / 16777216.0
    r0, r1 = np.sqrt(-2.0 * np.log(ua(r[0]))), np.sqrt(-2.0 * np.log(ua(r[2])))
    a0, a1 = 2.0 * np.pi * ub(r[1]), 2.0 * np.pi * ub(r[3])
    nz = np.stack([r0 * np.cos(a0), r0 * np.sin(a0), r1 * np.cos(a1), r1 * np.sin(a1)], axis=-1).reshape(B, nblk * 4)[:, :V]
    tx = 1.0 - 2.0 * codeword_bit
    llr = (tx + sigma[:, None] * nz) * inv2[:, None]
    x = np.concatenate([llr, np.zeros((B, Cn))], 1).astype(np.float32)
    return x, np.full((B, V), codeword_bit, dtype=np.uint8)


def count_failures(pcm, logical, err, hard):
    """LossFunc.forward(train=0) of quantum/neural_BP.py:338-348 on 0/1 arrays:
    (syndrome failures, logical failures among syndrome-ok, total)."""
    r = (np.asarray(err, np.int64) ^ np.asarray(hard, np.int64))
    syn_bad = ((r @ np.asarray(pcm, np.int64).T) % 2).any(1)
    log_bad = ((r @ np.asarray(logical, np.int64).T) % 2).any(1) if logical is not None and len(logical) else np.zeros(len(r), bool)
    a = int(syn_bad.sum())
    b = int((~syn_bad & log_bad).sum())
    return a, b, a + b
