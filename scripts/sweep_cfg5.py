"""BASELINE config 5: correlation sweep at Sintel (436x1024 -> 55x128 features) and FlyingThings
(540x960 -> 68x120), batch 1..64: FlowNetC patch-21 sampler fwd / bwd vs RAFT all-pairs build + lookup.
Writes gpurun_out/r1_sweep_cfg5.json.  (The reference CPU path at these sizes is covered by
bench.py's cpu_baseline scaling: its cost is linear in B*C*H*W*441.)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from understanding_flow_robustness_b200 import CorrBlock, backend, coords_grid

Q = (1, 1, 21, 21, 0, 0, 1, 1, 2, 2, 1, 1)


def timeit(fn, n=10, warm=2):
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


def inb(C, H, W):
    sh = sum(max(0, H - abs(2 * d)) for d in range(-10, 11))
    sw = sum(max(0, W - abs(2 * d)) for d in range(-10, 11))
    return 2.0 * C * sh * sw


rows = []
for name, (H, W) in {"kitti_48x160": (48, 160), "sintel_55x128": (55, 128), "things_68x120": (68, 120)}.items():
    for B in (1, 2, 4, 8, 16, 32, 64):
        a = torch.randn(B, 256, H, W, device="cuda")
        b = torch.randn(B, 256, H, W, device="cuda")
        g = torch.randn(B, 21, 21, H, W, device="cuda")
        tf = timeit(lambda: backend.forward(a, b, *Q))
        tb = timeit(lambda: backend.backward(a, b, g, *Q))
        del g
        row = {"dataset": name, "B": B, "sampler_fwd_ms": tf, "sampler_bwd_ms": tb,
               "sampler_pairs_per_s": B / (tf + tb) * 1e3,
               "sampler_fwd_bwd_inbounds_tflops": 3 * inb(256, H, W) * B / ((tf + tb) * 1e-3) / 1e12}
        vol_gb = B * (H * W) ** 2 * 4 * (1 + 0.25 + 1 / 16 + 1 / 64) / 1e9
        if vol_gb < 60:
            with torch.no_grad():
                blk = [None]

                def build():
                    blk[0] = None
                    blk[0] = CorrBlock(a, b, 4, 4)
                tbuild = timeit(build, n=5, warm=2)
                c = coords_grid(B, H, W, "cuda") + 3.0 * torch.randn(B, 2, H, W, device="cuda")
                tl = timeit(lambda: blk[0](c), n=10, warm=2)
                blk[0] = None
            row.update({"raft_build_ms": tbuild, "raft_lookup_ms": tl, "raft_ms_per_iter": (tbuild + 12 * tl) / 12,
                        "raft_volume_GB": vol_gb, "raft_build_GBps": vol_gb / (tbuild * 1e-3)})
        rows.append(row)
        print(json.dumps(row))
        del a, b
        torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/r1_sweep_cfg5.json", "w"), indent=1)
