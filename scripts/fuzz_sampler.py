"""Randomised parity sweep of the sampler (forward + both gradients) against the CPU oracle (GPU box).
Shapes are drawn around the edges of the register-blocked kernels: channel tails, ragged row parity classes,
images smaller than the patch radius, W % 8 != 0.   python scripts/fuzz_sampler.py [n] [seed]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import sampler_oracle
from understanding_flow_robustness_b200 import _lib, spatial_correlation_sample

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst = 0.0
fast = 0
for i in range(n):
    P, dpw = ((21, 2), (9, 1))[int(rng.integers(2))]
    dph = int(rng.choice([dpw, dpw, 1, 2, 3]))
    B, C = int(rng.integers(1, 4)), int(rng.choice([1, 3, 7, 8, 12, 31, 32, 33, 40, 64, 70, 129, 196]))
    H, W = int(rng.integers(1, 30)), int(rng.choice([4, 8, 12, 16, 20, 28, 36, 44, 64, 68]))
    in1 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    in2 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    gout = rng.standard_normal((B, P, P, H, W)).astype(np.float32)
    a = torch.from_numpy(in1).cuda().requires_grad_()
    b = torch.from_numpy(in2).cuda().requires_grad_()
    out = spatial_correlation_sample(a, b, kernel_size=1, patch_size=P, dilation_patch=(dph, dpw))
    out.backward(torch.from_numpy(gout).cuda())
    ref = sampler_oracle.forward(in1, in2, 1, P, 1, 0, 1, (dph, dpw))
    r1, r2 = sampler_oracle.backward(in1, in2, gout, 1, P, 1, 0, 1, (dph, dpw))

    def rel(x, y):
        return float(np.abs(x - y).max() / max(np.abs(y).max(), 1e-30))

    e = max(rel(out.detach().cpu().numpy(), ref), rel(a.grad.cpu().numpy(), r1), rel(b.grad.cpu().numpy(), r2))
    f = _lib.lib().b200corr_sampler_uses_fast_path(B, C, H, W, 1, 1, P, P, 0, 0, 1, 1, dph, dpw, 1, 1, 0, 1)
    fast += f
    worst = max(worst, e)
    if e > 1e-5:
        print("MISMATCH", (B, C, H, W, P, dph, dpw), e, "fast" if f else "generic")
print(f"{n} cases ({fast} on the register-blocked kernels), worst rel err {worst:.2e}")
sys.exit(0 if worst <= 1e-5 else 1)
