"""Runs the RAFT lookup backward, the volume backward and the alt_cuda_corr forward / backward a few times (ncu target)."""
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
    raft_corr.volume_backward(glv, f1, f2, 1.0 / 16.0, "tf32")
    f1n, f2n = f1.permute(0, 2, 3, 1).contiguous(), f2.permute(0, 2, 3, 1).contiguous()
    cn = c.permute(0, 2, 3, 1).reshape(B, 1, 48, 160, 2).contiguous()
    raft_corr.alt_cuda_corr.backward(f1n, f2n, cn, torch.randn(B, 1, 81, 48, 160, device="cuda"), 4)
torch.cuda.synchronize()
print("ok", float(glv[0].abs().max()), float(alt[0, 40, 5, 5]))
