"""alt_cuda_corr forward / backward per pyramid level, BASELINE config 3 (B=4).  Run once per library variant
(B200CORR_LIB=... selects the .so)."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from understanding_flow_robustness_b200 import alt_cuda_corr, coords_grid  # noqa: E402

B, C, H, W, r = 4, 256, 48, 160, 4
torch.manual_seed(0)
f1 = torch.randn(B, C, H, W, device="cuda")
f2 = torch.randn(B, C, H, W, device="cuda")


def timeit(fn, n=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {"lib": os.environ.get("B200CORR_LIB", "default")}
f1n = f1.permute(0, 2, 3, 1).contiguous()
pyr = [f2]
for _ in range(3):
    pyr.append(F.avg_pool2d(pyr[-1], 2, stride=2))
f2n = [p.permute(0, 2, 3, 1).contiguous() for p in pyr]
for sigma in (0.0, 3.0, 10.0):
    coords = coords_grid(B, H, W, "cuda") + sigma * torch.randn(B, 2, H, W, device="cuda")
    cn = coords.permute(0, 2, 3, 1)
    row = {}
    for lvl in range(4):
        ci = (cn / 2 ** lvl).reshape(B, 1, H, W, 2).contiguous()
        g = torch.randn(B, 1, 81, H, W, device="cuda")
        t_f = timeit(lambda: alt_cuda_corr.forward(f1n, f2n[lvl], ci, r))
        t_b = timeit(lambda: alt_cuda_corr.backward(f1n, f2n[lvl], ci, g, r), 5)
        flop = 2.0 * B * H * W * 100 * C
        row[f"level{lvl}"] = {"fwd_ms": round(t_f, 4), "fwd_tflops": round(flop / t_f / 1e9, 2), "bwd_ms": round(t_b, 4),
                              "bwd_tflops": round(2 * flop / t_b / 1e9, 2)}
    row["sum_fwd_ms"] = round(sum(row[f"level{i}"]["fwd_ms"] for i in range(4)), 4)
    res[f"sigma_{sigma}"] = row
print(json.dumps(res))
