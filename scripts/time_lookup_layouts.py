"""RAFT pyramid build + lookup (4 levels, radius 4, C=256; CUDA-graph replay, 12 lookups with fresh coordinates)
with the row-major and the blocked volume layout at the BASELINE config-5 feature sizes (GPU box).
    python scripts/time_lookup_layouts.py [B ...]      -> one JSON line per (shape, B)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from understanding_flow_robustness_b200 import coords_grid, raft_corr

BS = [int(x) for x in sys.argv[1:]] or [4]


def graph_time(fn, reps):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if __name__ == "__main__":
  for name, (H, W) in {"kitti_48x160": (48, 160), "sintel_55x128": (55, 128), "things_68x120": (68, 120)}.items():
      for B in BS:
          f1 = torch.randn(B, 256, H, W, device="cuda")
          f2 = torch.randn(B, 256, H, W, device="cuda")
          cs = [coords_grid(B, H, W, "cuda") + 3.0 * torch.randn(B, 2, H, W, device="cuda") for _ in range(12)]
          res = {"shape": name, "B": B}
          vol_gb = B * (H * W) ** 2 * 4 * (1 + 0.25 + 1 / 16 + 1 / 64) / 1e9
          for lay in ("rowmajor", "blocked"):
              hold = [None]

              def build():
                  hold[0] = None
                  r = raft_corr.allpairs_pyramid(f1, f2, 4, "tf32", blocked=(lay == "blocked"))
                  hold[0] = r if lay == "blocked" else (r, 0)
              with torch.no_grad():
                  tb = graph_time(build, 5)
                  build()
                  pyr, mask = hold[0]
                  tl = graph_time(lambda: [raft_corr.lookup_forward(pyr, c, 4, H, W, blocked_levels=mask) for c in cs], 5) / 12
              res[lay] = {"mask": mask, "build_ms": round(tb, 4), "build_GBps": round(vol_gb / tb * 1e3), "lookup_us": round(tl * 1e3, 2),
                          "ms_per_iter": round((tb + 12 * tl) / 12, 4)}
              hold[0] = None
              del pyr
          print(json.dumps(res), flush=True)
          del f1, f2, cs
          torch.cuda.empty_cache()
