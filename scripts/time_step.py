"""Times sampler forward / backward (graph-free, CUDA events) at the bench shape -- variant comparisons."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from understanding_flow_robustness_b200 import backend

dev = torch.device("cuda", 0)
B, C, H, W = 8, 256, 48, 160
a = torch.randn(B, C, H, W, device=dev)
b = torch.randn(B, C, H, W, device=dev)
g = torch.randn(B, 21, 21, H, W, device=dev)
out = torch.empty(B, 21, 21, H, W, device=dev)
g1, g2 = torch.empty_like(a), torch.empty_like(b)
Q = (1, 1, 21, 21, 0, 0, 1, 1, 2, 2, 1, 1)


def t(fn, n=100):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for _ in range(300):
    backend.forward(a, b, *Q, out=out)
r = {"fwd_ms": t(lambda: backend.forward(a, b, *Q, out=out)),
     "bwd_ms": t(lambda: backend.backward(a, b, g, *Q, out=(g1, g2))),
     "fwd_ms_2": t(lambda: backend.forward(a, b, *Q, out=out)),
     "bwd_ms_2": t(lambda: backend.backward(a, b, g, *Q, out=(g1, g2)))}
print(json.dumps(r))
