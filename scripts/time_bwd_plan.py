"""Backward plan: makespan of the schedule (host) and kernel time (GPU) per B200CORR_BWD_UNIT_OVERHEAD value.

    python scripts/time_bwd_plan.py [--cpu]      (--cpu: only the host-side makespan table)
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from understanding_flow_robustness_b200 import _lib

HYPER = (1, 1, 21, 21, 0, 0, 1, 1, 2, 2, 1, 1)
COST = [14, 18, 22, 22, 18, 14] * 2          # source rows per unit of the 12 row groups at H = 48


def plan(B, C, H, W, overhead):
    os.environ["B200CORR_BWD_UNIT_OVERHEAD"] = str(overhead)
    L = _lib.lib()
    L.b200corr_sampler_backward_workspace_bytes.restype = ctypes.c_size_t
    n = L.b200corr_sampler_backward_workspace_bytes(B, C, H, W, *HYPER, 0)
    buf = np.zeros(n // 4, dtype=np.int32)
    _lib.check(L.b200corr_sampler_backward_plan(B, C, H, W, *HYPER, 0, buf.ctypes.data_as(ctypes.c_void_p), n), "plan")
    return buf


def makespan(buf, grid, per_sample=120, per_group=10):
    offs, ids = buf[:grid + 1], buf[grid + 1:]
    loads = [sum(COST[(u % per_sample) // per_group] for u in ids[offs[c]:offs[c + 1]]) for c in range(grid)]
    return max(loads), sum(loads) / grid, max(offs[c + 1] - offs[c] for c in range(grid))


if __name__ == "__main__":
    grid = 148
    for B in (4, 8, 16):
        for ov in (0, 2, 4, 6):
            b = plan(B, 256, 48, 160, ov)
            assert sorted(b[grid + 1:]) == list(range(B * 120)), "plan is not a permutation of the units"
            mx, mean, cnt = makespan(b, grid)
            print(f"B {B} overhead {ov}: max {mx} rows, mean {mean:.1f}, ratio {mx / mean:.4f}, most units on a CTA {cnt}")
    if "--cpu" in sys.argv:
        sys.exit(0)
    from understanding_flow_robustness_b200 import backend
    for B in (4, 8):
        a = torch.randn(B, 256, 48, 160, device="cuda")
        c = torch.randn(B, 256, 48, 160, device="cuda")
        g = torch.randn(B, 21, 21, 48, 160, device="cuda")
        for ov in (-1, 0, 1, 2, 4, -1, 0):
            os.environ["B200CORR_BWD_PLAN_LS"] = "0" if ov < 0 else "1"     # -1: plain LPT (round 1's schedule)
            os.environ["B200CORR_BWD_UNIT_OVERHEAD"] = str(max(ov, 0))
            backend._PLANS.clear()
            for _ in range(5):
                backend.backward(a, c, g, *HYPER)
            torch.cuda.synchronize()
            ts = []
            for _ in range(30):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                backend.backward(a, c, g, *HYPER)
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            print(f"B {B} overhead {ov}: backward {ts[len(ts) // 2]:.4f} ms (min {ts[0]:.4f})")
