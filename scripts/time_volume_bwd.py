"""CorrBlock volume backward (csrc/raft_volume_bwd.cu) alone at BASELINE config 3: dF1, dF2 from a gradient pyramid."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from understanding_flow_robustness_b200 import raft_corr  # noqa: E402

B, C, H, W, L = 4, 256, 48, 160, 4
torch.manual_seed(0)
f1 = torch.randn(B, C, H, W, device="cuda")
f2 = torch.randn(B, C, H, W, device="cuda")
glv = [torch.randn(B * H * W, 1, H >> l, W >> l, device="cuda") for l in range(L)]
for prec in ("tf32",):
    for _ in range(2):
        raft_corr.volume_backward(glv, f1, f2, 1 / math.sqrt(C), prec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        raft_corr.volume_backward(glv, f1, f2, 1 / math.sqrt(C), prec)
    e1.record()
    torch.cuda.synchronize()
    print(prec, "volume_backward ms:", e0.elapsed_time(e1) / 5)
