"""the PWC-Net correlation call sites (models/PWCNet.py:42-50: patch 9, dilation_patch 1)
at the five pyramid levels of a 384x1280 input, batch 8 -- this library vs the reference's CUDA kernels
compiled for sm_100a (oracle/_ref); plus warp() (PWCNet.py:164-204) as one kernel vs the reference's torch ops
(restated in oracle/warp_oracle.py) on the same GPU.  Writes gpurun_out/r1_pwc_levels.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import build_ref_cuda, warp_oracle
from understanding_flow_robustness_b200 import _lib, backend, warp

q = (1, 1, 9, 9, 0, 0, 1, 1, 1, 1, 1, 1)
B = 8
LEVELS = [(6, 196, 6, 20), (5, 128, 12, 40), (4, 96, 24, 80), (3, 64, 48, 160), (2, 32, 96, 320)]


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


try:
    ref = build_ref_cuda.load_module("ref_sampler_cuda")
except Exception as e:
    ref = None
    print("reference CUDA extension unavailable:", e)
rows = []
for lvl, C, H, W in LEVELS:
    a = torch.randn(B, C, H, W, device="cuda")
    b = torch.randn(B, C, H, W, device="cuda")
    g = torch.randn(B, 9, 9, H, W, device="cuda")
    L = _lib.lib()
    row = {"level": lvl, "shape": [B, C, H, W],
           "fast_fwd": bool(L.b200corr_sampler_uses_fast_path(B, C, H, W, *q, 0, 0)),
           "fast_bwd": bool(L.b200corr_sampler_uses_fast_path(B, C, H, W, *q, 0, 1)),
           "ours_fwd_ms": timeit(lambda: backend.forward(a, b, *q)),
           "ours_bwd_ms": timeit(lambda: backend.backward(a, b, g, *q))}
    if ref is not None:
        row["reference_cuda_fwd_ms"] = timeit(lambda: ref.forward(a, b, *q), n=5, warm=1)
        row["reference_cuda_bwd_ms"] = timeit(lambda: ref.backward(a, b, g, *q), n=3, warm=1)
        o1, o2 = backend.forward(a, b, *q), ref.forward(a, b, *q)
        row["max_rel_diff_fwd"] = float((o1 - o2).abs().max() / o2.abs().max())
    if lvl < 6:   # warp() runs in front of every correlation below the top level (PWCNet.py:293-339)
        xw = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
        fw = (2.0 * torch.randn(B, 2, H, W, device="cuda")).requires_grad_()

        def fb(fn):
            def run():
                o = fn(xw, fw)
                torch.autograd.grad(o, (xw, fw), g[:, 0, 0, None].expand_as(o).contiguous())
            return run
        with torch.no_grad():
            row["warp_ours_fwd_ms"] = timeit(lambda: warp(xw, fw))
            row["warp_reference_fwd_ms"] = timeit(lambda: warp_oracle.warp(xw, fw))
            row["warp_max_rel_diff"] = float((warp(xw, fw) - warp_oracle.warp(xw, fw)).abs().max() / warp_oracle.warp(xw, fw).abs().max())
        row["warp_ours_fwd_bwd_ms"] = timeit(fb(warp))
        row["warp_reference_fwd_bwd_ms"] = timeit(fb(warp_oracle.warp))
    rows.append(row)
    print(json.dumps(row))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/r1_pwc_levels.json", "w"), indent=1)
