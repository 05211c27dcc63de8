"""CorrBlock backward (B=4, 256x48x160, 12 lookups), five runs per layout with the number of cudaMallocs made inside the
backward: 2.5 ms whenever the caching allocator serves the 1.25 GB gradient pyramid from its cache, 4-17 ms when it has to
grow (GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from understanding_flow_robustness_b200 import CorrBlock, coords_grid
B, C, H, W = 4, 256, 48, 160
f1g = torch.randn(B, C, H, W, device="cuda").requires_grad_()
f2g = torch.randn(B, C, H, W, device="cuda").requires_grad_()
cs = [coords_grid(B, H, W, "cuda") + 3.0 * torch.randn(B, 2, H, W, device="cuda") for _ in range(12)]
def whole(layout):
    blk = CorrBlock(f1g, f2g, 4, 4, layout=layout)
    loss = sum(blk(c).sum() for c in cs)
    torch.cuda.synchronize()
    s0 = torch.cuda.memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); loss.backward(); e1.record(); torch.cuda.synchronize()
    s1 = torch.cuda.memory_stats()
    f1g.grad = f2g.grad = None
    return round(e0.elapsed_time(e1), 3), s1["num_device_alloc"] - s0["num_device_alloc"], s1["num_alloc_retries"] - s0["num_alloc_retries"]
for layout in ("auto", "rowmajor", "auto"):
    print(layout, [whole(layout) for _ in range(5)], round(torch.cuda.memory_reserved() / 1e9, 2), "GB reserved")
