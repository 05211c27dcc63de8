"""Breakdown of the CorrBlock backward (B=4, 256x48x160, 12 lookups): gradient-pyramid zero fill, 12 lookup
backward launches, then round 2's volume backward (csrc/raft_volume_bwd.cu: feature pooling + two tcgen05 GEMMs
over the unfolded pyramid + unpool) next to what it replaced (pyramid fold + two cuBLAS GEMMs) (GPU box)."""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from understanding_flow_robustness_b200 import CorrBlock, coords_grid, raft_corr

B, C, H, W = 4, 256, 48, 160
f1 = torch.randn(B, C, H, W, device="cuda")
f2 = torch.randn(B, C, H, W, device="cuda")
cs = [coords_grid(B, H, W, "cuda") + 3.0 * torch.randn(B, 2, H, W, device="cuda") for _ in range(12)]
g = torch.randn(B, 324, H, W, device="cuda")


def t(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


shapes = [(H >> l, W >> l) for l in range(4)]
glv = [torch.zeros(B * H * W, 1, h, w, device="cuda") for h, w in shapes]
res = {"zero_fill_ms": t(lambda: [v.zero_() for v in glv]),
       "lookup_bwd_x12_ms": t(lambda: [raft_corr.lookup_backward(glv, c, g, 4, H, W) for c in cs]),
       "pyramid_fold_ms": t(lambda: raft_corr.pyramid_backward(glv, B, H, W))}
res["volume_backward_tf32_ms"] = t(lambda: raft_corr.volume_backward(glv, f1, f2, 1.0 / 16.0, "tf32"))
gvol = glv[0].view(B, H * W, H * W)
for tf32 in (True, False):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    res["bmm_x2_%s_ms" % ("tf32" if tf32 else "fp32")] = t(lambda: (torch.bmm(f2.reshape(B, C, H * W), gvol.transpose(1, 2)),
                                                                  torch.bmm(f1.reshape(B, C, H * W), gvol)), 3)
f1g, f2g = f1.clone().requires_grad_(), f2.clone().requires_grad_()


def whole():
    blk = CorrBlock(f1g, f2g, 4, 4, precision="tf32")
    loss = sum(blk(c).sum() for c in cs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss.backward()
    e1.record()
    torch.cuda.synchronize()
    f1g.grad = f2g.grad = None
    return e0.elapsed_time(e1)


whole()
res["corrblock_backward_ms"] = min(whole() for _ in range(3))
print(json.dumps(res))
