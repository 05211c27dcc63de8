"""RAFT lookup backward (12 launches into one gradient pyramid, graph replay) for the library in B200CORR_LIB."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from understanding_flow_robustness_b200 import coords_grid, raft_corr  # noqa: E402

B, H, W = 4, 48, 160
torch.manual_seed(0)
cs = [coords_grid(B, H, W, "cuda") + 3.0 * torch.randn(B, 2, H, W, device="cuda") for _ in range(12)]
g = torch.randn(B, 324, H, W, device="cuda")
glv = [torch.zeros(B * H * W, 1, H >> l, W >> l, device="cuda") for l in range(4)]


def run():
    for c in cs:
        raft_corr.lookup_backward(glv, c, g, 4, H, W)


run()
torch.cuda.synchronize()
chk = float(sum(v.double().abs().sum() for v in glv))
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    run()
for _ in range(2):
    gr.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    gr.replay()
e1.record()
torch.cuda.synchronize()
print(os.environ.get("B200CORR_LIB", "default").split("/")[-1], "lookup backward us:", round(e0.elapsed_time(e1) / 120 * 1e3, 2), "checksum", chk)
