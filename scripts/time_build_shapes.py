import os, sys
sys.path.insert(0, os.getcwd())
import torch
from understanding_flow_robustness_b200 import raft_corr
for (H, W) in ((68, 120), (48, 160), (55, 128), (46, 100)):
    for B in (1, 4):
        f1 = torch.randn(B, 256, H, W, device="cuda"); f2 = torch.randn(B, 256, H, W, device="cuda")
        keep = [None]
        def build():
            keep[0] = None
            keep[0] = raft_corr.allpairs_pyramid(f1, f2, 4, "tf32")
        for _ in range(3): build()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): build()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gb = B * (H * W) ** 2 * 4 * (1 + 0.25 + 1 / 16 + 1 / 64) / 1e9
        print(f"{H}x{W} B={B}: build {ms:.4f} ms  {gb / ms * 1e3:.0f} GB/s")
for prec in ("tf32", "tf32x3", "fp32"):
    f1 = torch.randn(4, 256, 48, 160, device="cuda"); f2 = torch.randn(4, 256, 48, 160, device="cuda")
    keep = [None]
    def build():
        keep[0] = None
        keep[0] = raft_corr.allpairs_pyramid(f1, f2, 4, prec)
    for _ in range(2): build()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): build()
    e1.record(); torch.cuda.synchronize()
    print(f"48x160 B=4 precision {prec}: build {e0.elapsed_time(e1) / 5:.4f} ms")
