"""RAFT build + lookup with fp32 and fp16 storage of the blocked levels (B=4, 48x160 and the config-5 shapes;
CUDA-graph replay, 12 lookups with fresh coordinates).  python scripts/time_fp16_storage.py [B ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from understanding_flow_robustness_b200 import coords_grid, raft_corr
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from time_lookup_layouts import graph_time  # noqa: E402

BS = [int(x) for x in sys.argv[1:]] or [4]
for name, (H, W) in {"kitti_48x160": (48, 160), "sintel_55x128": (55, 128), "things_68x120": (68, 120)}.items():
    for B in BS:
        f1 = torch.randn(B, 256, H, W, device="cuda")
        f2 = torch.randn(B, 256, H, W, device="cuda")
        cs = [coords_grid(B, H, W, "cuda") + 3.0 * torch.randn(B, 2, H, W, device="cuda") for _ in range(12)]
        res = {"shape": name, "B": B}
        for storage in ("fp32", "fp16"):
            hold = [None]

            def build():
                hold[0] = None
                hold[0] = raft_corr.allpairs_pyramid(f1, f2, 4, "tf32", blocked=True, storage=storage)
            with torch.no_grad():
                tb = graph_time(build, 5)
                build()
                pyr, mask = hold[0]
                tl = graph_time(lambda: [raft_corr.lookup_forward(pyr, c, 4, H, W, blocked_levels=mask) for c in cs], 5) / 12
            res[storage] = {"build_ms": round(tb, 4), "lookup_us": round(tl * 1e3, 2), "ms_per_iter": round((tb + 12 * tl) / 12, 4),
                            "volume_MB": round(sum(v.numel() * v.element_size() for v in pyr) / 1e6)}
            hold[0] = None
            del pyr
        print(json.dumps(res), flush=True)
        del f1, f2, cs
        torch.cuda.empty_cache()
