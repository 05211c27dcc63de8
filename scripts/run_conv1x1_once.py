"""A few pipelined lookup -> conv1x1 calls at BASELINE config 3 (ncu target)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from understanding_flow_robustness_b200 import CorrBlock, coords_grid  # noqa: E402

torch.manual_seed(0)
B, C, H, W = 4, 256, 48, 160
f1 = torch.randn(B, C, H, W, device="cuda")
f2 = torch.randn(B, C, H, W, device="cuda")
conv = torch.nn.Conv2d(324, 256, 1).cuda()
with torch.no_grad():
    blk = CorrBlock(f1, f2, 4, 4, precision="tf32")
    for i in range(3):
        c = coords_grid(B, H, W, "cuda") + 3.0 * torch.randn(B, 2, H, W, device="cuda")
        out = blk.lookup_convc1(c, conv.weight, conv.bias, impl="pipelined")
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
