"""Small invocation of every kernel (compute-sanitizer target)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from understanding_flow_robustness_b200 import AlternateCorrBlock, CorrBlock, coords_grid, spatial_correlation_sample

torch.manual_seed(0)
for (B, C, H, W, P, dp) in [(1, 128, 11, 20, 21, 2), (1, 128, 6, 12, 9, 1), (1, 6, 7, 9, 5, 2)]:
    a = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
    b = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
    out = spatial_correlation_sample(a, b, 1, P, 1, 0, 1, dp)
    out.square().sum().backward()
f1 = torch.randn(1, 40, 11, 20, device="cuda", requires_grad=True)
f2 = torch.randn(1, 40, 11, 20, device="cuda", requires_grad=True)
c = coords_grid(1, 11, 20, "cuda") + 4.0 * torch.randn(1, 2, 11, 20, device="cuda")
blk = CorrBlock(f1, f2, 3, 3)
(blk(c).sum() + blk(c + 30.0).sum()).backward()
alt = AlternateCorrBlock(f1, f2, 3, 3)(c)
alt.sum().backward()
torch.cuda.synchronize()
print("sanitize_small ok", float(out.sum()), float(alt.sum()))
