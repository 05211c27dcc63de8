"""BASELINE config 5: correlation sweep at Sintel (436x1024 -> 55x128 features) and FlyingThings
(540x960 -> 68x120), batch 1..64: FlowNetC patch-21 sampler fwd / bwd vs RAFT all-pairs build + lookup.
Writes gpurun_out/r2_sweep_cfg5.json.  The "vs reference CPU on host cores" column: the reference's own
correlation.cpp (oracle/_ref, all cores, batch = core count so that its batch-parallel backward uses them all, 32 of
256 channels timed and scaled x8 -- the loops are linear in C) and the reference's torch CorrBlock on CPU tensors
(matmul + 3 avg_pool2d + 12 x 4 grid_sample), once per dataset."""
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


def reference_cpu(H, W):
    """pairs/s of the reference CPU sampler fwd+bwd and ms/iter of the reference torch CorrBlock on the host cores."""
    import time

    import torch.nn.functional as F
    out = {"cores": os.cpu_count()}
    torch.set_num_threads(os.cpu_count())
    try:
        from oracle import build_ref
        be = build_ref.load_backend()
        Bc, Cs = os.cpu_count() or 8, 32
        a, b = torch.randn(Bc, Cs, H, W), torch.randn(Bc, Cs, H, W)
        t0 = time.perf_counter()
        o = be.forward(a, b, *Q)
        t1 = time.perf_counter()
        be.backward(a, b, torch.ones_like(o), *Q)
        t2 = time.perf_counter()
        out["sampler_cpu"] = {"batch": Bc, "channels_timed": Cs, "scaled_by": 256 // Cs, "fwd_s_scaled": (t1 - t0) * 256 / Cs,
                              "bwd_s_scaled": (t2 - t1) * 256 / Cs, "pairs_per_s": Bc / ((t2 - t0) * 256 / Cs)}
    except Exception as e:
        out["sampler_cpu"] = {"unavailable": f"{type(e).__name__}: {e}"}
    # models/raft/corr.py on CPU tensors, B = 1
    f1, f2 = torch.randn(1, 256, H, W), torch.randn(1, 256, H, W)
    t0 = time.perf_counter()
    corr = torch.matmul(f1.view(1, 256, -1).transpose(1, 2), f2.view(1, 256, -1)).view(H * W, 1, H, W) / 16.0
    pyr = [corr]
    for _ in range(3):
        pyr.append(F.avg_pool2d(pyr[-1], 2, stride=2))
    t1 = time.perf_counter()
    c = (coords_grid(1, H, W) + 3.0 * torch.randn(1, 2, H, W)).permute(0, 2, 3, 1)
    r = 4
    d = torch.linspace(-r, r, 2 * r + 1)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), axis=-1)
    for i in range(4):
        cl = c.reshape(H * W, 1, 1, 2) / 2 ** i + delta.view(1, 9, 9, 2)
        Hl, Wl = pyr[i].shape[-2:]
        xg, yg = cl.split([1, 1], dim=-1)
        F.grid_sample(pyr[i], torch.cat([2 * xg / (Wl - 1) - 1, 2 * yg / (Hl - 1) - 1], dim=-1), align_corners=True)
    t2 = time.perf_counter()
    out["raft_cpu_B1"] = {"build_ms": (t1 - t0) * 1e3, "lookup_ms": (t2 - t1) * 1e3,
                          "ms_per_iter": ((t1 - t0) + 12 * (t2 - t1)) * 1e3 / 12}
    return out


rows = []
cpu_rows = {}
for name, (H, W) in {"kitti_48x160": (48, 160), "sintel_55x128": (55, 128), "things_68x120": (68, 120)}.items():
    cpu_rows[name] = reference_cpu(H, W)
    print(json.dumps({"dataset": name, "reference_cpu": cpu_rows[name]}))
    for B in (1, 2, 4, 8, 16, 32, 64):
        a = torch.randn(B, 256, H, W, device="cuda")
        b = torch.randn(B, 256, H, W, device="cuda")
        g = torch.randn(B, 21, 21, H, W, device="cuda")
        tf = timeit(lambda: backend.forward(a, b, *Q))
        tb = timeit(lambda: backend.backward(a, b, g, *Q))
        del g
        row = {"dataset": name, "B": B, "sampler_fwd_ms": tf, "sampler_bwd_ms": tb,
               "sampler_pairs_per_s": B / (tf + tb) * 1e3,
               "sampler_vs_reference_cpu": B / (tf + tb) * 1e3 / cpu_rows[name].get("sampler_cpu", {}).get("pairs_per_s", float("nan")),
               "sampler_fwd_bwd_inbounds_tflops": 3 * inb(256, H, W) * B / ((tf + tb) * 1e-3) / 1e12}
        vol_gb = B * (H * W) ** 2 * 4 * (1 + 0.25 + 1 / 16 + 1 / 64) / 1e9
        if vol_gb < 60:
            with torch.no_grad():
                blk = [None]

                def build():
                    blk[0] = None
                    blk[0] = CorrBlock(a, b, 4, 4, precision="tf32")
                tbuild = timeit(build, n=5, warm=2)
                c = coords_grid(B, H, W, "cuda") + 3.0 * torch.randn(B, 2, H, W, device="cuda")
                tl = timeit(lambda: blk[0](c), n=10, warm=2)
                blk[0] = None
            row.update({"raft_build_ms": tbuild, "raft_lookup_ms": tl, "raft_ms_per_iter": (tbuild + 12 * tl) / 12,
                        "raft_vs_reference_cpu_per_sample": cpu_rows[name]["raft_cpu_B1"]["ms_per_iter"] / ((tbuild + 12 * tl) / 12 / B),
                        "raft_volume_GB": vol_gb, "raft_build_GBps": vol_gb / (tbuild * 1e-3)})
        rows.append(row)
        print(json.dumps(row))
        del a, b
        torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"reference_cpu": cpu_rows, "rows": rows}, open("gpurun_out/r2_sweep_cfg5.json", "w"), indent=1)
