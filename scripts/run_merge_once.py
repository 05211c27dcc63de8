"""Runs the fused FlowNetC merge block forward + backward twice at (8,256,48,160) (ncu target)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from understanding_flow_robustness_b200 import correlate_merge

a = torch.randn(8, 256, 48, 160, device="cuda", requires_grad=True)
b = torch.randn(8, 256, 48, 160, device="cuda", requires_grad=True)
r = torch.randn(8, 32, 48, 160, device="cuda")
g = torch.randn(8, 473, 48, 160, device="cuda")
for _ in range(2):
    correlate_merge(a, b, r).backward(g)
torch.cuda.synchronize()
print("ok", float(a.grad[0, 0, 0, 0]))
