"""Times one RAFT lookup (B=4, 48x160, 4 levels, radius 4) with fresh coordinates per launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from understanding_flow_robustness_b200 import coords_grid, raft_corr

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
f1 = torch.randn(B, 256, 48, 160, device="cuda")
f2 = torch.randn(B, 256, 48, 160, device="cuda")
pyr = raft_corr.allpairs_pyramid(f1, f2, 4, "tf32")
cs = [coords_grid(B, 48, 160, "cuda") + 3.0 * torch.randn(B, 2, 48, 160, device="cuda") for _ in range(12)]
for mode in ("grid_sample", "direct"):
    for _ in range(3):
        for c in cs:
            o = raft_corr.lookup_forward(pyr, c, 4, 48, 160, mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        for c in cs:
            o = raft_corr.lookup_forward(pyr, c, 4, 48, 160, mode)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 60
    gb = B * 7680 * (324 * 4 + 4 * 100 * 4) / 1e9
    print(f"lookup {mode}: {ms * 1e3:.1f} us  ({gb / ms * 1e3:.0f} GB/s algorithmic)")

# ---- backward: 12 lookups accumulate into one gradient pyramid, then pyramid fold + two bmm
from understanding_flow_robustness_b200 import CorrBlock

f1g = f1.clone().requires_grad_()
f2g = f2.clone().requires_grad_()
def fwd_bwd(parts):
    blk = CorrBlock(f1g, f2g, 4, 4)
    outs = [blk(c) for c in cs]
    loss = sum(o.sum() for o in outs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss.backward()
    e1.record()
    torch.cuda.synchronize()
    f1g.grad = f2g.grad = None
    return e0.elapsed_time(e1)
fwd_bwd(0)
ts = [fwd_bwd(0) for _ in range(3)]
print(f"CorrBlock backward (12 lookups + pyramid fold + 2 bmm, B={B}): {min(ts):.3f} ms")
g = torch.randn_like(o)
glv = [torch.zeros_like(v) for v in pyr]
for _ in range(2):
    raft_corr.lookup_backward(glv, cs[0], g, 4, 48, 160, "grid_sample")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for c in cs:
    raft_corr.lookup_backward(glv, c, g, 4, 48, 160, "grid_sample")
e1.record()
torch.cuda.synchronize()
print(f"lookup backward: {e0.elapsed_time(e1) / 12 * 1e3:.1f} us per lookup")
