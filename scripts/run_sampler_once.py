"""Runs the FlowNetC-shaped sampler forward+backward a few times (ncu target)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from understanding_flow_robustness_b200 import backend

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
a = torch.randn(B, 256, 48, 160, device="cuda")
b = torch.randn(B, 256, 48, 160, device="cuda")
g = torch.randn(B, 21, 21, 48, 160, device="cuda")
q = (1, 1, 21, 21, 0, 0, 1, 1, 2, 2, 1, 1)
for _ in range(reps):
    out = backend.forward(a, b, *q)
    g1, g2 = backend.backward(a, b, g, *q)
torch.cuda.synchronize()
print("ok", float(out[0, 10, 10, 5, 5]), float(g1[0, 0, 0, 0]))
