"""Where a patch-attack iteration spends its time at N ranks (the evidence behind DESIGN.md section 4).

    torchrun --nproc-per-node N scripts/attack_timeline.py        (or plain python for N = 1)

Rank 0 runs 3 iterations of the bench's attack loop (global batch 64, 384x1280, FlowNetC with the fused merge block,
eager AND graph-replayed) under torch.profiler (CUPTI kernel records) and reports, per iteration: wall time,
GPU-busy time (union of kernel intervals), idle gaps, and kernel time by class -- NCCL all-reduce, cuDNN/cuBLAS
library kernels, this library's kernels, ATen elementwise.  Writes gpurun_out/r2_attack_timeline_g{N}.json."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from understanding_flow_robustness_b200 import attack  # noqa: E402
from understanding_flow_robustness_b200.harness import FlowNetCHarness  # noqa: E402

OURS = ("sampler_", "merge_grad", "patch_compose", "patch_reduce")


def classify(name):
    n = name.lower()
    if "nccl" in n:
        return "nccl"
    if any(k in name for k in OURS):
        return "this_library"
    if any(k in n for k in ("cudnn", "cutlass", "xmma", "gemm", "conv", "implicit", "wgrad", "dgrad", "sm90", "sm100", "nchw", "nhwc")):
        return "library_conv"
    return "aten_elementwise"


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    fmt = torch.channels_last
    net = FlowNetCHarness(fused_merge=True).to(dev).eval().to(memory_format=fmt)
    for q in net.parameters():
        q.requires_grad_(False)
    G, H, W, p = 64, 384, 1280, 100
    idx = attack.shard_slice(G, rank, world)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    i1 = torch.rand(len(idx), 3, H, W, device=dev, generator=g).contiguous(memory_format=fmt)
    i2 = torch.rand(len(idx), 3, H, W, device=dev, generator=g).contiguous(memory_format=fmt)
    patch0 = torch.rand(1, 3, p, p, device=dev)
    mask = attack.circle_mask(p, dev)
    cfg = attack.PatchAttackConfig()
    target = torch.empty(len(idx), 2, H, W, device=dev)
    s_patch = patch0.clone()
    s_pl = attack.sample_placements(len(idx), H, W, p, cfg, g, dev)

    def clean_fn(a, b):
        with torch.no_grad():
            return -net(a, b)

    def grad_fn(pt, pl):
        return attack.patch_gradient(net, i1, i2, pt, mask, patch0, pl, target, G, cfg.alpha)

    g_clean = attack.GraphedGradient(clean_fn, [i1, i2])
    g_grad = attack.GraphedGradient(grad_fn, [s_patch, s_pl])
    state = {"patch": patch0.clone()}

    def iteration(graphed):
        target.copy_(g_clean() if graphed else clean_fn(i1, i2))
        pl = attack.sample_placements(len(idx), H, W, p, cfg, g, dev)
        pt = state["patch"]
        for _ in range(cfg.max_count):
            packed = g_grad(pt, pl) if graphed else grad_fn(pt, pl)
            if world > 1:
                dist.all_reduce(packed)
            pt, _ = attack.apply_patch_step(pt, packed, cfg)
        state["patch"] = pt

    out = {"n_gpus": world, "pairs_per_rank": len(idx)}
    for graphed in (False, True):
        for _ in range(2):
            iteration(graphed)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        iters = 3
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as prof:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                iteration(graphed)
            e1.record()
            torch.cuda.synchronize()
        wall_ms = e0.elapsed_time(e1) / iters
        if rank == 0:
            evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time_total > 0]
            spans = sorted((e.time_range.start, e.time_range.end) for e in evs)
            busy, cur_s, cur_e = 0.0, None, None
            for s, e in spans:
                if cur_e is None or s > cur_e:
                    if cur_e is not None:
                        busy += cur_e - cur_s
                    cur_s, cur_e = s, e
                else:
                    cur_e = max(cur_e, e)
            if cur_e is not None:
                busy += cur_e - cur_s
            by = {}
            for e in evs:
                c = classify(e.name)
                by[c] = by.get(c, 0.0) + e.device_time_total
            nccl = [e.device_time_total for e in evs if classify(e.name) == "nccl"]
            out["graphed" if graphed else "eager"] = {
                "ms_per_iter": wall_ms, "gpu_busy_ms_per_iter": busy / 1e3 / iters,
                "idle_ms_per_iter": wall_ms - busy / 1e3 / iters, "kernels_per_iter": len(evs) // iters,
                "kernel_ms_per_iter_by_class": {k: v / 1e3 / iters for k, v in sorted(by.items())},
                "nccl_allreduce_us_each": sorted(x for x in nccl)[len(nccl) // 2] if nccl else None,
                "nccl_calls_per_iter": len(nccl) // iters}
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"r2_attack_timeline_g{world}.json"), "w"), indent=1)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
