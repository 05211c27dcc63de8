"""Times the RAFT lookup (B=4, 48x160, 4 levels, radius 4; CUDA-graph replay of 12 lookups with fresh coordinates)
on a row-major and on a blocked pyramid (GPU box).  python scripts/time_lookup_layouts.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from understanding_flow_robustness_b200 import coords_grid, raft_corr

B = 4
f1 = torch.randn(B, 256, 48, 160, device="cuda")
f2 = torch.randn(B, 256, 48, 160, device="cuda")
cs = [coords_grid(B, 48, 160, "cuda") + 3.0 * torch.randn(B, 2, 48, 160, device="cuda") for _ in range(12)]
res = {}
for name in ("rowmajor", "blocked"):
    if name == "blocked":
        pyr, mask = raft_corr.allpairs_pyramid(f1, f2, 4, "tf32", blocked=True)
    else:
        pyr, mask = raft_corr.allpairs_pyramid(f1, f2, 4, "tf32"), 0

    def run():
        for c in cs:
            raft_corr.lookup_forward(pyr, c, 4, 48, 160, blocked_levels=mask)
    run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    res[name + "_us"] = e0.elapsed_time(e1) / 120 * 1e3
    res[name + "_mask"] = mask
    del pyr, g
print(json.dumps(res))
