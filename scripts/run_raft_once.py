"""Runs the RAFT CorrBlock build + lookups + alt path a few times (ncu target)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from understanding_flow_robustness_b200 import AlternateCorrBlock, CorrBlock, coords_grid

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
f1 = torch.randn(B, 256, 48, 160, device="cuda")
f2 = torch.randn(B, 256, 48, 160, device="cuda")
c = coords_grid(B, 48, 160, "cuda") + 3.0 * torch.randn(B, 2, 48, 160, device="cuda")
with torch.no_grad():
    for _ in range(reps):
        blk = CorrBlock(f1, f2, 4, 4, precision="tf32")
        o = blk(c)
        o2 = blk(c + 1.0)
    alt = AlternateCorrBlock(f1, f2, 4, 4)(c)
torch.cuda.synchronize()
print("ok", float(o[0, 40, 5, 5]), float(alt[0, 40, 5, 5]))
