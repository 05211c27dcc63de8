import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from understanding_flow_robustness_b200 import coords_grid, raft_corr
B = 4
f1 = torch.randn(B, 256, 48, 160, device="cuda"); f2 = torch.randn(B, 256, 48, 160, device="cuda")
c = coords_grid(B, 48, 160, "cuda") + 3.0 * torch.randn(B, 2, 48, 160, device="cuda")
f1n, f2n = f1.permute(0, 2, 3, 1).contiguous(), f2.permute(0, 2, 3, 1).contiguous()
cn = c.permute(0, 2, 3, 1).reshape(B, 1, 48, 160, 2).contiguous()
for _ in range(3):
    raft_corr.alt_cuda_corr.forward(f1n, f2n, cn, 4)
torch.cuda.synchronize(); print("ok")
