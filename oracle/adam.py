"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the optimizer the reference's training loop
uses -- torch.optim.Adam(decoder.parameters(), lr=3e-4, weight_decay=1e-9) (quantum/decoder_v2_4.py:323, stepped at
:338).  The algorithm lives in PyTorch (third party, present in this image): this follows its default single-tensor
path (torch/optim/adam.py `_single_tensor_adam`, amsgrad=False, maximize=False, L2 weight decay folded into the
gradient).  Pinned against torch.optim.Adam itself in tests/test_adam.py (CPU)."""
import numpy as np


class AdamOracle(object):
    def __init__(self, n, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-9, dtype=np.float32):
        self.lr, self.b1, self.b2, self.eps, self.wd = lr, betas[0], betas[1], eps, weight_decay
        self.m = np.zeros(n, dtype)
        self.v = np.zeros(n, dtype)
        self.t = 0
        self.dtype = dtype

    def step(self, w, g):
        """w, g: 1-D arrays; returns the updated w (a new array)."""
        d = self.dtype
        w = w.astype(d)
        g = g.astype(d)
        self.t += 1
        if self.wd != 0:
            g = g + d(self.wd) * w
        self.m = self.m + d(1 - self.b1) * (g - self.m)                        # exp_avg.lerp_(grad, 1 - beta1)
        self.v = self.v * d(self.b2) + d(1 - self.b2) * g * g                  # mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        bc1 = 1 - self.b1 ** self.t
        bc2 = 1 - self.b2 ** self.t
        step_size = d(self.lr / bc1)
        denom = np.sqrt(self.v) / d(bc2 ** 0.5) + d(self.eps)
        return (w - step_size * (self.m / denom)).astype(d)
