"""Same-GPU comparison against the reference's own CUDA code (oracle/_ref, compiled unmodified for
sm_100a) and against torch's matmul / avg_pool2d / grid_sample CorrBlock path.  Test/bench
infrastructure: writes gpurun_out/r1_vs_reference_cuda.json."""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from oracle import build_ref_cuda
from understanding_flow_robustness_b200 import AlternateCorrBlock, CorrBlock, alt_cuda_corr, backend, coords_grid

res = {}


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


q = (1, 1, 21, 21, 0, 0, 1, 1, 2, 2, 1, 1)
B = 8
a = torch.randn(B, 256, 48, 160, device="cuda")
b = torch.randn(B, 256, 48, 160, device="cuda")
g = torch.randn(B, 21, 21, 48, 160, device="cuda")
ours_f = timeit(lambda: backend.forward(a, b, *q))
ours_b = timeit(lambda: backend.backward(a, b, g, *q))
res["sampler_B8"] = {"ours_fwd_ms": ours_f, "ours_bwd_ms": ours_b}
try:
    ref = build_ref_cuda.load_module("ref_sampler_cuda")
    rf = timeit(lambda: ref.forward(a, b, *q), n=5, warm=1)
    rb = timeit(lambda: ref.backward(a, b, g, *q), n=3, warm=1)
    o1, o2 = backend.forward(a, b, *q), ref.forward(a, b, *q)
    res["sampler_B8"].update({"reference_cuda_fwd_ms": rf, "reference_cuda_bwd_ms": rb,
                              "speedup_fwd": rf / ours_f, "speedup_bwd": rb / ours_b,
                              "max_rel_diff_fwd": float((o1 - o2).abs().max() / o2.abs().max())})
    g1, g2 = backend.backward(a, b, g, *q)
    r1, r2 = ref.backward(a, b, g, *q)
    res["sampler_B8"]["max_rel_diff_bwd"] = max(float((g1 - r1).abs().max() / r1.abs().max()),
                                                float((g2 - r2).abs().max() / r2.abs().max()))
except Exception as e:
    res["sampler_B8"]["reference_cuda"] = f"unavailable: {type(e).__name__}: {e}"

# RAFT config 3
Bq, C, H, W = 4, 256, 48, 160
f1 = torch.randn(Bq, C, H, W, device="cuda")
f2 = torch.randn(Bq, C, H, W, device="cuda")
coords = coords_grid(Bq, H, W, "cuda") + 3.0 * torch.randn(Bq, 2, H, W, device="cuda")


def torch_build():
    corr = torch.matmul(f1.view(Bq, C, H * W).transpose(1, 2), f2.view(Bq, C, H * W))
    corr = corr.view(Bq, H, W, 1, H, W) / torch.sqrt(torch.tensor(C).float())
    corr = corr.reshape(Bq * H * W, 1, H, W)
    pyr = [corr]
    for _ in range(3):
        corr = F.avg_pool2d(corr, 2, stride=2)
        pyr.append(corr)
    return pyr


def torch_lookup(pyr):
    r = 4
    c = coords.permute(0, 2, 3, 1)
    outs = []
    for i in range(4):
        dx = torch.linspace(-r, r, 2 * r + 1)
        dy = torch.linspace(-r, r, 2 * r + 1)
        delta = torch.stack(torch.meshgrid(dy, dx, indexing="ij"), axis=-1).to(coords.device)   # corr.py:80-82
        cl = c.reshape(Bq * H * W, 1, 1, 2) / 2 ** i + delta.view(1, 2 * r + 1, 2 * r + 1, 2)
        Hl, Wl = pyr[i].shape[-2:]
        xg, yg = cl.split([1, 1], dim=-1)
        grid = torch.cat([2 * xg / (Wl - 1) - 1, 2 * yg / (Hl - 1) - 1], dim=-1)
        outs.append(F.grid_sample(pyr[i], grid, align_corners=True).view(Bq, H, W, -1))
    return torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous().float()


with torch.no_grad():
    torch.backends.cuda.matmul.allow_tf32 = False
    hold = [None]

    def tb():
        hold[0] = None
        hold[0] = torch_build()
    t_build_fp32 = timeit(tb, n=5, warm=2)
    torch.backends.cuda.matmul.allow_tf32 = True
    t_build_tf32 = timeit(tb, n=5, warm=2)
    t_look = timeit(lambda: torch_lookup(hold[0]), n=10, warm=2)
    hold[0] = None
    blk = [None]

    def ob():
        blk[0] = None
        blk[0] = CorrBlock(f1, f2, 4, 4)
    o_build = timeit(ob, n=10, warm=2)
    o_look = timeit(lambda: blk[0](coords), n=20, warm=2)
    res["raft_B4"] = {"torch_build_fp32_ms": t_build_fp32, "torch_build_tf32_ms": t_build_tf32, "torch_lookup_ms": t_look,
                      "ours_build_ms": o_build, "ours_lookup_ms": o_look,
                      "torch_ms_per_iter": (t_build_fp32 + 12 * t_look) / 12, "ours_ms_per_iter": (o_build + 12 * o_look) / 12}
    blk[0] = None
    # alt path
    alt = AlternateCorrBlock(f1, f2, 4, 4)
    o_alt = timeit(lambda: alt(coords), n=5, warm=1)
    res["alt_B4"] = {"ours_ms_per_iter": o_alt}
    try:
        ref_alt = build_ref_cuda.load_module("ref_alt_cuda_corr")
        f1n = f1.permute(0, 2, 3, 1).contiguous()
        pyr2 = [f2]
        for _ in range(3):
            pyr2.append(F.avg_pool2d(pyr2[-1], 2, stride=2))
        f2n = [p.permute(0, 2, 3, 1).contiguous() for p in pyr2]
        cn = coords.permute(0, 2, 3, 1)

        def ref_alt_iter():
            outs = []
            for i in range(4):
                ci = (cn / 2 ** i).reshape(Bq, 1, H, W, 2).contiguous()
                (c_,) = ref_alt.forward(f1n, f2n[i], ci, 4)
                outs.append(c_.squeeze(1))
            return torch.stack(outs, 1).reshape(Bq, -1, H, W) / math.sqrt(C)
        r_alt = timeit(ref_alt_iter, n=3, warm=1)
        res["alt_B4"].update({"reference_cuda_ms_per_iter": r_alt, "speedup": r_alt / o_alt})
    except Exception as e:
        res["alt_B4"]["reference_cuda"] = f"unavailable: {type(e).__name__}: {e}"

os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/r1_vs_reference_cuda.json", "w"), indent=1)
print(json.dumps(res, indent=1))
