"""Runs the RAFT lookup backward and the alt_cuda_corr forward a few times (ncu target)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from understanding_flow_robustness_b200 import AlternateCorrBlock, coords_grid, raft_corr

B = 4
f1 = torch.randn(B, 256, 48, 160, device="cuda")
f2 = torch.randn(B, 256, 48, 160, device="cuda")
c = coords_grid(B, 48, 160, "cuda") + 3.0 * torch.randn(B, 2, 48, 160, device="cuda")
with torch.no_grad():
    pyr = raft_corr.allpairs_pyramid(f1, f2, 4, "tf32")
    glv = [torch.zeros_like(v) for v in pyr]
    g = torch.randn(B, 324, 48, 160, device="cuda")
    for _ in range(3):
        raft_corr.lookup_backward(glv, c, g, 4, 48, 160, "grid_sample")
    alt = AlternateCorrBlock(f1, f2, 4, 4)(c)
torch.cuda.synchronize()
print("ok", float(glv[0].abs().max()), float(alt[0, 40, 5, 5]))
