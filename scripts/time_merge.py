"""Times the plain and the fused-merge forward kernels and the merge block (GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json

import torch

import bench
from understanding_flow_robustness_b200 import backend, merge_block

dev = torch.device("cuda", 0)
B, C, H, W = 8, 256, 48, 160
a = torch.randn(B, C, H, W, device=dev)
b = torch.randn(B, C, H, W, device=dev)
out = torch.empty(B, 21, 21, H, W, device=dev)
merged = torch.empty(B, 473, H, W, device=dev)
Q = (1, 1, 21, 21, 0, 0, 1, 1, 2, 2, 1, 1)


def t(fn, n=50):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for _ in range(200):
    backend.forward(a, b, *Q, out=out)
res = {"plain_fwd_ms": t(lambda: backend.forward(a, b, *Q, out=out)),
       "merge_fwd_ms": t(lambda: merge_block.merge_forward(a, b, merged, 32)),
       "plain_fwd_ms_again": t(lambda: backend.forward(a, b, *Q, out=out))}
if "fwdonly" not in sys.argv:
    res["merge_block"] = bench.merge_bench(dev)
print(json.dumps(res))
