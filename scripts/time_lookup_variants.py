"""RAFT build + 12 lookups (blocked layout, graph replay) for the library selected by B200CORR_LIB."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from understanding_flow_robustness_b200 import CorrBlock, coords_grid  # noqa: E402

B, C, H, W = 4, 256, 48, 160
torch.manual_seed(0)
f1 = torch.randn(B, C, H, W, device="cuda")
f2 = torch.randn(B, C, H, W, device="cuda")
cs = [coords_grid(B, H, W, "cuda") + 3.0 * torch.randn(B, 2, H, W, device="cuda") for _ in range(12)]
with torch.no_grad():
    blk = CorrBlock(f1, f2, 4, 4, precision="tf32")
    ref = [blk(c).clone() for c in cs[:2]]

    def lookups():
        for c in cs:
            blk(c)
    lookups()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        lookups()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(os.environ.get("B200CORR_LIB", "default").split("/")[-1], "lookup us:", round(e0.elapsed_time(e1) / 120 * 1e3, 2),
          "checksum", float(sum(r.double().sum() for r in ref)))
