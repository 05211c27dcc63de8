#!/usr/bin/env python
"""bench.py -- correlation hot path on B200: `python bench.py --gpus N --steps K --warmup W`.

Headline metric (BASELINE.json): FlowNetC corr fwd+bwd pairs/s -- `spatial_correlation_sample`
(kernel 1, patch 21, dilation_patch 2) forward + backward w.r.t. both feature maps on
synthetic KITTI-shaped features (B, 256, 48, 160), B = 8 pairs per GPU (BASELINE config 2's
correlation layer).  One "step" = one forward + one backward over one batch.  The same JSON line
carries RAFT corr+lookup ms/iter (config 3) under "raft".

  value     : pairs/s with inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       : same metric through the public operator with HOST (pinned) buffers: H2D of
              in1/in2/grad_out and D2H of out/grad_in1/grad_in2 inside the timed region
  roofline  : dominant kernel (sampler_bwd_kernel, two launches per step) against the FP32-FMA peak
              measured in this run (the path is FP32-pipe bound); in-bounds FLOPs
  cpu_baseline / --impl reference : the reference's own CPU extension (oracle/_ref, compiled from
              correlation.cpp) on the host cores, bounded sample, scaled to pairs/s
N > 1 (torchrun): the headline stays the sampler line (weak scaling, every rank its own batch of 8 pairs, no
data-path collective -- SURVEY.md 8e: the correlation is independent per image pair; barrier +
max-over-ranks timing).  The metric of the path that HAS a collective is in the same line and repeated in
its last key `summary` (so a truncated tail still carries it):
  attack                 : BASELINE config 4, universal 100x100 patch, global batch 64 sharded r::N (strong
                           scaling), 120 KB gradient all-reduce (NCCL) per inner step, >= 20 timed iterations
  universal_perturbation : global_attacks/universal_perturbation.py loop, batch 64 at 256x640, 3.9 MB all-reduce
  nccl_value_check       : N-rank all-reduced gradients == rank 0's single-process gradients of the same batch
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = dict(B=8, C=256, H=48, W=160, patch=21, dilation_patch=2)
RAFT_CFG = dict(B=4, C=256, H=48, W=160, levels=4, radius=4, iters=12)
Q = (1, 1, 21, 21, 0, 0, 1, 1, 2, 2, 1, 1)


def inbounds_macs_per_pair(C, H, W, P, dp):
    r = (P - 1) // 2
    sh = sum(max(0, H - abs(d * dp)) for d in range(-r, r + 1))
    sw = sum(max(0, W - abs(d * dp)) for d in range(-r, r + 1))
    return C * sh * sw


# ---------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def wait_first_sample(self, timeout=8.0):
        """NVML start-up holds the driver lock for a while: never let it fall into a timed region."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.samples and time.perf_counter() - t0 < timeout:
            time.sleep(0.05)

    def close(self):
        if self.proc is not None:
            self.proc.terminate()

    def stop(self, t0, t1):
        r = self.window(t0, t1)
        self.close()
        return r

    def window(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        mhz, mx, reasons = [], None, set()
        for (t, line) in self.samples:
            if t < t0 or t > t1 + 0.1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                mhz.append(float(f[0]))
                mx = float(f[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mhz.sort()
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(mhz), "window_s": round(t1 - t0, 3)}


# ---------------------------------------------------------------------------- reference CPU arm
def reference_cpu_pairs_per_s(steps, warmup, budget_s=2.5, exact_config1=False):
    """Times the reference's own CPU path (oracle/_ref, else the C oracle port) on all host cores.

    A step is the bench workload (batch 8, 48x160, patch 21, dilation 2) at a reduced channel count
    C_s chosen so one step costs ~budget_s; the work is exactly linear in C (correlation.cpp:20-35
    loops over c), so pairs/s is scaled by C_s / 256."""
    import numpy as np
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    kind, fwd, bwd = None, None, None
    try:
        from oracle import build_ref
        be = build_ref.load_backend()
        kind = "reference"

        def fwd(a, b):
            return be.forward(a, b, *Q)

        def bwd(a, b, g):
            return be.backward(a, b, g, *Q)
    except Exception:
        from oracle import sampler_oracle
        kind = "port"

        def fwd(a, b):
            return torch.from_numpy(sampler_oracle.forward(a.numpy(), b.numpy(), 1, 21, 1, 0, 1, 2))

        def bwd(a, b, g):
            return sampler_oracle.backward(a.numpy(), b.numpy(), g.numpy(), 1, 21, 1, 0, 1, 2)

    B, H, W = CFG["B"], CFG["H"], CFG["W"]

    def run(C):
        g = torch.Generator().manual_seed(0)
        a = torch.randn(B, C, H, W, generator=g)
        b = torch.randn(B, C, H, W, generator=g)
        t0 = time.perf_counter()
        out = fwd(a, b)
        go = torch.ones_like(out) if isinstance(out, torch.Tensor) else torch.ones(out.shape)
        bwd(a, b, go)
        return time.perf_counter() - t0

    t4 = run(4)                                    # calibration
    Cs = 4
    while Cs < CFG["C"] and t4 * (2 * Cs / 4) <= budget_s:
        Cs *= 2
    for _ in range(warmup):
        run(Cs)
    ts = [run(Cs) for _ in range(steps)]
    t = sum(ts) / len(ts)
    pairs_per_s = B / (t * CFG["C"] / Cs)
    sample = (f"batch {B} x {Cs} of {CFG['C']} channels x {H}x{W} fwd+bwd per step "
              f"({Cs}/{CFG['C']} of the MACs, time EXTRAPOLATED x{CFG['C'] // Cs}: the CPU loops are linear in C, "
              f"correlation.cpp:20-35); {steps} steps, mean")
    cb = {"value": pairs_per_s, "unit": "pairs/s", "cores": cores, "kind": kind, "sample": sample,
          "channel_subsample": {"channels_timed": Cs, "channels_workload": CFG["C"], "time_scaled_by": CFG["C"] // Cs}}
    if exact_config1:
        # BASELINE config 1 exactly (B=1, C=256): no extrapolation.  The backward parallelises over the batch
        # only (correlation.cpp:148-149), so at B=1 it runs on ONE thread whatever the core count.
        g = torch.Generator().manual_seed(1)
        a = torch.randn(1, CFG["C"], H, W, generator=g)
        b = torch.randn(1, CFG["C"], H, W, generator=g)
        t0 = time.perf_counter()
        out = fwd(a, b)
        t1 = time.perf_counter()
        bwd(a, b, torch.ones(tuple(out.shape)))
        t2 = time.perf_counter()
        cb["config1_exact"] = {"what": f"(1,{CFG['C']},{H},{W}) patch 21 dilation_patch 2, one forward + one backward, not extrapolated",
                               "fwd_s": t1 - t0, "bwd_s": t2 - t1, "pairs_per_s": 1.0 / (t2 - t0),
                               "threads": {"fwd": cores, "bwd": 1}}
    return pairs_per_s, t * 1e3, cb


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, ms, cb = reference_cpu_pairs_per_s(max(1, args.steps), max(0, args.warmup), exact_config1=True)
    line = {"impl": "reference", "metric": "FlowNetC corr fwd+bwd pairs/s", "value": v, "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.gpus), "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "channel_subsample": cb["channel_subsample"], "config1_exact": cb.get("config1_exact")}
    _emit(line)


def workload_config(n):
    c = CFG
    return {"workload": f"spatial_correlation_sample fwd+bwd, features ({c['B']},{c['C']},{c['H']},{c['W']}) per GPU "
                        f"(FlowNetC corr layer of 384x1280 KITTI pairs, batch 8), kernel 1, patch {c['patch']}, "
                        f"dilation_patch {c['dilation_patch']}",
            "pairs_per_gpu": c["B"], "global_pairs": c["B"] * n, "parallelism": f"pair-sharded x{n}, no collective",
            "l2": "working set per step 470 MB > 126 MB L2 (inputs 126 MB, grad_out 108 MB, outputs 234 MB); no explicit flush"}


def gpu_cpu_affinity(index):
    """CPU ids local to GPU `index` (NUMA node of its PCIe root), from `nvidia-smi topo -m`; None if unknown."""
    try:
        txt = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        lines = [ln for ln in txt.splitlines() if ln.strip()]
        hdr = next(ln for ln in lines if "CPU Affinity" in ln)
        cols = [c.strip() for c in hdr.split("\t")]
        ci = cols.index("CPU Affinity")
        row = next(ln for ln in lines if ln.split("\t")[0].strip() == f"GPU{index}")
        spec = [c.strip() for c in row.split("\t")][ci]
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return cpus & os.sched_getaffinity(0) or None
    except Exception:
        return None


# ---------------------------------------------------------------------------- GPU arm
_REAL_STDOUT = None


def _claim_stdout():
    """Libraries print to fd 1 (NCCL's version banner, torchrun children): keep the real stdout for the
    one JSON line and send everything else to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-raft", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-attack", action="store_true")
    ap.add_argument("--attack-batch", type=int, default=64)
    ap.add_argument("--attack-iters", type=int, default=20)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    from understanding_flow_robustness_b200 import _lib, backend

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    B, C, H, W, P = CFG["B"], CFG["C"], CFG["H"], CFG["W"], CFG["patch"]
    torch.manual_seed(rank)
    in1 = torch.randn(B, C, H, W, device=dev)
    in2 = torch.randn(B, C, H, W, device=dev)
    gout = torch.randn(B, P, P, H, W, device=dev)

    def step():
        out = backend.forward(in1, in2, *Q)
        g1, g2 = backend.backward(in1, in2, gout, *Q)
        return out, g1, g2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        clocks.wait_first_sample()
    for _ in range(args.warmup):
        step()
    # the GPU comes out of idle here (the wait for the first clock sample above): W steps are ~10 ms,
    # far less than the boost ramp.  Keep stepping, untimed, until the clocks have settled.
    t_ramp = time.perf_counter()
    while time.perf_counter() - t_ramp < 0.5:
        for _ in range(20):
            step()
        torch.cuda.synchronize()
    barrier()
    # ---- timed region: K steps, per-kernel-group events on the launching (current) stream.  The C-ABI
    # launches of one step are captured once into two CUDA graphs (forward; backward) and replayed, so a
    # slow or noisy host cannot starve the GPU inside the timed region (one B200 step is 0.8 ms; the
    # eager path costs the host ~0.2 ms per step, 1 ms+ on a loaded box).
    timed_loop = "cuda-graph replay of the step's C-ABI launches (forward graph + backward graph)"
    per_step_launches = None
    try:
        n_c = L.b200corr_launch_count()
        g_fwd, g_bwd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_fwd):
            out = backend.forward(in1, in2, *Q)
        with torch.cuda.graph(g_bwd):
            g1, g2 = backend.backward(in1, in2, gout, *Q)
        per_step_launches = int(L.b200corr_launch_count() - n_c)
        for _ in range(3):
            g_fwd.replay()
            g_bwd.replay()
        torch.cuda.synchronize()

        def run_fwd():
            g_fwd.replay()

        def run_bwd():
            g_bwd.replay()
    except Exception as e:   # capture unavailable: time the eager calls
        timed_loop = f"eager C-ABI calls (graph capture failed: {type(e).__name__})"
        torch.cuda.synchronize()

        def run_fwd():
            backend.forward(in1, in2, *Q)

        def run_bwd():
            backend.backward(in1, in2, gout, *Q)
    n0 = L.b200corr_launch_count()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    e_start = torch.cuda.Event(enable_timing=True)
    e_end = torch.cuda.Event(enable_timing=True)
    e_start.record()
    for k in range(args.steps):
        ev[k][0].record()
        run_fwd()
        ev[k][1].record()
        run_bwd()
        ev[k][2].record()
    e_end.record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = L.b200corr_launch_count() - n0
    if per_step_launches is not None:
        launches = per_step_launches * args.steps   # replayed kernels are not seen by the library's counter
    # the clock record needs a few 200 ms samples under this load: keep the same step loop running
    # (untimed) until the sampled window is >= 0.7 s
    while time.perf_counter() - t_wall0 < 0.7:
        for _ in range(20):
            step()
        torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    ms_total = e_start.elapsed_time(e_end)
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    bwd_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    clk = clocks.window(t_wall0, t_wall1) if rank == 0 else None
    # FP32 FFMA peak (the roofline denominator), measured right behind the timed region: same clocks,
    # same power / thermal state as the kernels it is compared with (measured after the attack bench it
    # read 65.8 instead of 72.6 TFLOP/s on one box: the power cap of a GPU that had just run the conv stack)
    import ctypes
    pk = ctypes.c_float()
    t_p0 = time.perf_counter()
    peaks_seen = []
    while time.perf_counter() - t_p0 < 0.45 or not peaks_seen:      # long enough for two 200 ms clock samples
        _lib.check(L.b200corr_measure_fp32_peak(4000, ctypes.byref(pk), _lib.current_stream(dev)), "fp32 peak")
        peaks_seen.append(float(pk.value))
    t_p1 = time.perf_counter()
    peaks_seen.sort()
    peak = peaks_seen[len(peaks_seen) // 2]
    peak_clk = None
    if rank == 0:
        pc = clocks.stop(t_p0, t_p1)
        peak_clk = {"sm_mhz": pc.get("sm_mhz"), "reasons": pc.get("reasons"), "probe_runs": len(peaks_seen),
                    "tflops_min_median_max": [peaks_seen[0], peak, peaks_seen[-1]]}
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region.  The public
    # host-buffer front end (SamplerHostPipeline) overlaps H2D / kernels / D2H of consecutive batches.
    from understanding_flow_robustness_b200.host_pipeline import SamplerHostPipeline
    # pinned buffers on the NUMA node of this GPU's PCIe root (first touch happens under this affinity)
    all_cpus = os.sched_getaffinity(0)
    near = gpu_cpu_affinity(local)
    if near:
        os.sched_setaffinity(0, near)
    nset = 2
    h_sets = [[torch.randn(B, C, H, W).pin_memory(), torch.randn(B, C, H, W).pin_memory(),
               torch.randn(B, P, P, H, W).pin_memory(), torch.empty(B, P, P, H, W).pin_memory(),
               torch.empty(B, C, H, W).pin_memory(), torch.empty(B, C, H, W).pin_memory()] for _ in range(nset)]
    pipe = SamplerHostPipeline((B, C, H, W), Q, dev)
    e2e_steps = max(4, min(args.steps, 20))
    t_w = time.perf_counter()
    k = 0
    while k < 4 or time.perf_counter() - t_w < 0.3:   # PCIe link and copy engines up to speed
        pipe.submit(*h_sets[k % nset])
        k += 1
        if k % 4 == 0:
            pipe.synchronize()
    pipe.synchronize()
    barrier()
    t_e0 = time.perf_counter()
    for k in range(e2e_steps):
        pipe.submit(*h_sets[k % nset])
    pipe.synchronize()
    t_e1 = time.perf_counter()
    barrier()
    t = torch.tensor([(t_e1 - t_e0) * 1e3], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / (float(t.item()) * 1e-3)
    h2d = sum(x.numel() for x in h_sets[0][:3]) * 4
    d2h = sum(x.numel() for x in h_sets[0][3:]) * 4
    # the roof of this number: the same copies with no kernels in between (both directions at once, as the
    # pipeline runs them), same pinned buffers, same slots
    s_a, s_b = pipe.s_in, pipe.s_out
    pe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    reps = 6
    pe[0].record()
    s_a.wait_event(pe[0])
    s_b.wait_event(pe[0])
    for k in range(reps):
        hs, di, do = h_sets[k % nset], pipe.dev_in[k % pipe.depth], pipe.dev_out[k % pipe.depth]
        with torch.cuda.stream(s_a):
            for d, h_ in zip(di, hs[:3]):
                d.copy_(h_, non_blocking=True)
        with torch.cuda.stream(s_b):
            for h_, d in zip(hs[3:], do):
                h_.copy_(d, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s_a)
    torch.cuda.current_stream().wait_stream(s_b)
    pe[1].record()
    torch.cuda.synchronize()
    copy_ms = pe[0].elapsed_time(pe[1]) / reps
    e2e_roof = {"copies_only_ms_per_step": copy_ms, "pairs_per_s": B / (copy_ms * 1e-3),
                "gb_per_s_each_way": h2d / (copy_ms * 1e-3) / 1e9,
                "what": "H2D and D2H of one step's buffers, concurrently, no kernels (this rank alone)"}
    del pipe
    os.sched_setaffinity(0, all_cpus)   # the CPU baseline below uses every host core

    attack_res = pert_res = nccl_res = None
    if not args.no_attack:
        del h_sets
        torch.cuda.empty_cache()
        attack_res = attack_bench(dev, rank, world, args.attack_batch, args.attack_iters)
        pert_res = perturbation_bench(dev, rank, world, args.attack_batch)
        if world > 1:
            nccl_res = nccl_check_bench(dev, rank, world, args.attack_batch)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (sampler_bwd_kernel: two launches per step)
    macs = inbounds_macs_per_pair(C, H, W, P, CFG["dilation_patch"])
    flop_launch = 2.0 * macs * B                      # one gradient (or the forward): in-bounds FLOPs
    bwd_launch_ms = bwd_ms / 2.0
    ach = flop_launch / (bwd_launch_ms * 1e-3) / 1e12
    traffic = None
    prof = os.path.join(ROOT, "profiles", "sampler_traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("sampler_bwd_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "fp32_fma", "kernel": "sampler_bwd_kernel (2 launches/step)", "achieved": ach,
                "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "peak_source": "FP32 FFMA micro-benchmark measured in this run right behind the timed region "
                               "(MEASURED_PEAKS.json has no FP32 figure); median of the probe runs",
                "peak_probe_clocks": peak_clk,
                "flops_counted": "in-bounds MACs x2 (dense count is 1.369x larger)",
                "step": {"fwd_ms": fwd_ms, "bwd_ms": bwd_ms,
                         "fwd_frac": flop_launch / (fwd_ms * 1e-3) / 1e12 / peak,
                         "fwd_bwd_frac": 3 * flop_launch / ((fwd_ms + bwd_ms) * 1e-3) / 1e12 / peak}}

    line = {"metric": "FlowNetC corr fwd+bwd pairs/s", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world), "timed_loop": timed_loop, "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "pcie_roof": e2e_roof, "frac_of_pcie_roof": e2e_value / world / e2e_roof["pairs_per_s"],
                    "how": "pinned host in1/in2/grad_out -> device -> fwd+bwd -> pinned host out/grad_in1/grad_in2, every step; "
                           "3-stream double-buffered pipeline, wall-clock over the steps incl. final drain" + ("; host buffers on the GPU's NUMA node" if near else "")},
            "gpu_launches": int(launches), "roofline": roofline}

    if attack_res is not None:
        line["attack"] = attack_res
        line["universal_perturbation"] = pert_res
        if nccl_res is not None:
            line["nccl_value_check"] = nccl_res
    if not args.no_raft and world == 1:
        try:
            line["reference_cuda"] = reference_cuda_bench(dev)
        except Exception as e:
            line["reference_cuda"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    if not args.no_raft:
        line["raft"] = raft_bench(dev)
        try:
            line["merge_block"] = merge_bench(dev)
        except Exception as e:
            line["merge_block"] = {"error": f"{type(e).__name__}: {e}"}
    if not args.no_cpu_baseline:
        try:
            _, _, cb = reference_cpu_pairs_per_s(steps=3, warmup=1, budget_s=3.0)
            line["cpu_baseline"] = cb
        except Exception as e:  # never lose the GPU numbers to a host-side problem
            line["cpu_baseline"] = {"value": None, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "unavailable",
                                    "sample": f"{type(e).__name__}: {e}"}
    # last key: the numbers a truncated tail must still carry (<= 300 bytes)
    summ = {"n_gpus": world, "corr_pairs_s": round(value, 1), "roofline_step": round(roofline["step"]["fwd_bwd_frac"], 3)}
    if attack_res is not None:
        summ.update(attack_iters_s=round(attack_res["value"], 3), attack_allreduce_B=attack_res["allreduce_bytes"],
                    pert_iters_s=round(pert_res["value"], 3), pert_allreduce_B=pert_res["allreduce_bytes"])
    if nccl_res is not None:
        summ["nccl_check_ok"] = nccl_res.get("ok")
    if "raft" in line:
        summ.update(raft_ms_iter=round(line["raft"]["ms_per_iter"], 4), raft_roofline=round(line["raft"]["roofline"]["frac"], 3))
    line["summary"] = summ
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def _timed_iters(fn, iters, dev, world, sync_each=False):
    """CUDA-event time of `iters` calls of fn, barrier + synchronize on both sides, MAX over ranks -> ms per call."""
    import torch
    import torch.distributed as dist

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / iters


def attack_bench(dev, rank, world, global_batch, iters=20):
    """BASELINE config 4: universal 100x100 patch attack on FlowNetC (random init), global batch of 384x1280
    pairs sharded r::G over the ranks, patch gradient all-reduced (NCCL) every inner step.  Strong scaling: the
    global batch is fixed.  iteration = clean forward + max_count x (compose + forward + backward + all-reduce
    + update).

    Network: the reference's own models/FlowNetC.py through the shims where its file is present (row
    `reference_body`, eager), and the same network (same weights: state_dict copied) with the merge block on the
    fused kernel -- the headline row, its per-rank work (clean forward; gradient step) replayed from CUDA graphs
    so the ~600 launches per step cost no host time; the all-reduce and the update stay eager between replays."""
    import torch
    import torch.distributed as dist

    from understanding_flow_robustness_b200 import attack
    from understanding_flow_robustness_b200.harness import FlowNetCHarness, reference_models
    torch.manual_seed(0)                                   # same weights / patch on every rank
    fmt = torch.channels_last                              # cuDNN's preferred conv layout
    net = FlowNetCHarness(fused_merge=True).to(dev).eval().to(memory_format=fmt)
    for q in net.parameters():
        q.requires_grad_(False)
    H, W, p = 384, 1280, 100
    idx = attack.shard_slice(global_batch, rank, world)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    i1 = torch.rand(len(idx), 3, H, W, device=dev, generator=g).contiguous(memory_format=fmt)
    i2 = torch.rand(len(idx), 3, H, W, device=dev, generator=g).contiguous(memory_format=fmt)
    patch0 = torch.rand(1, 3, p, p, device=dev)
    mask = attack.circle_mask(p, dev)
    cfg = attack.PatchAttackConfig()

    # BASELINE config 2 on the way: whole-network forward+backward pairs/s on 8 pairs of this rank's shard
    nb = min(8, len(idx))

    def net_step():
        a = i1[:nb].clone().requires_grad_(True)
        net(a, i2[:nb]).mean().backward()
    net_step()
    net_ms = _timed_iters(net_step, 5, dev, 1)
    net_pairs_per_s = nb / (net_ms * 1e-3)

    # ---- headline: graph-replayed per-rank work
    target = torch.empty(len(idx), 2, H, W, device=dev)
    s_patch = patch0.clone()
    s_pl = attack.sample_placements(len(idx), H, W, p, cfg, g, dev)

    def clean_fn(a, b):
        with torch.no_grad():
            return -net(a, b)

    def grad_fn(pt, pl):
        return attack.patch_gradient(net, i1, i2, pt, mask, patch0, pl, target, global_batch, cfg.alpha)

    how = "cuda-graph replay of the per-rank work (clean forward; gradient step), all-reduce + update eager"
    try:
        g_clean = attack.GraphedGradient(clean_fn, [i1, i2])
        g_grad = attack.GraphedGradient(grad_fn, [s_patch, s_pl])
    except Exception as e:                               # capture unavailable: time the eager loop
        torch.cuda.synchronize()
        how = f"eager (graph capture failed: {type(e).__name__}: {e})"[:200]
        g_clean = lambda: clean_fn(i1, i2)                # noqa: E731
        g_grad = grad_fn
    state = {"patch": patch0.clone(), "loss": None}

    def iteration():
        target.copy_(g_clean())
        pl = attack.sample_placements(len(idx), H, W, p, cfg, g, dev)
        pt = state["patch"]
        for _ in range(cfg.max_count):
            packed = g_grad(pt, pl)
            if world > 1:
                dist.all_reduce(packed)                   # the one collective of the path (NCCL over NVLink)
            pt, state["loss"] = attack.apply_patch_step(pt, packed, cfg)
        state["patch"] = pt

    for _ in range(2):
        iteration()
    ms = _timed_iters(iteration, iters, dev, world)
    res = {"metric": "patch-attack iters/s", "value": 1e3 / ms, "unit": "iters/s", "ms_per_iter": ms, "iters_timed": iters,
           "pairs_per_s": global_batch * 1e3 / ms, "scaling": "strong", "global_batch": global_batch,
           "pairs_per_rank": len(idx), "n_gpus": world, "timed_loop": how,
           "config": "FlowNetC (reference parameters, random init, merge block on the fused kernel) 384x1280, 100x100 "
                     "circular patch placed by the compose kernel, max_count 2, cosine loss, patch-gradient all-reduce "
                     "(NCCL) per inner step",
           "allreduce_bytes": int(patch0.numel() * 4 + 4), "allreduces_per_iter": cfg.max_count,
           "final_loss": float(state["loss"]),
           "flownetc_fwd_bwd_pairs_per_s_per_gpu": net_pairs_per_s,
           "flownetc_config": f"BASELINE config 2: FlowNetC random init, forward+backward, 384x1280, batch {nb}, 1 GPU, eager"}
    del g_clean, g_grad
    # ---- same loop, eager, fused network (what the graph replaces) and on the reference's unmodified FlowNetC.py
    def eager_iter(model):
        def it():
            state["patch"] = attack.patch_attack_iteration(model, i1, i2, state["patch"], mask, patch0, cfg,
                                                           global_batch, g)[0]
        return it
    it = eager_iter(net)
    it()
    res["eager_ms_per_iter"] = _timed_iters(it, 5, dev, world)
    if reference_models.available():
        try:
            # the reference's memory format (contiguous NCHW): its correlate() hands conv3's output to the
            # operator as is, and the operator -- like the reference's, CHECK_CONTIGUOUS -- refuses channels_last
            ref = reference_models.reference_flownetc().to(dev).eval()
            ref.load_state_dict(net.state_dict())
            for q in ref.parameters():
                q.requires_grad_(False)
            j1, j2 = i1.contiguous(), i2.contiguous()

            def it():
                state["patch"] = attack.patch_attack_iteration(ref, j1, j2, state["patch"], mask, patch0, cfg,
                                                               global_batch, g)[0]
            it()
            rms = _timed_iters(it, 5, dev, world)
            res["reference_body"] = {"what": "the reference's UNMODIFIED models/FlowNetC.py (correlate() -> this package's "
                                             "sampler through the shims), same weights, eager, NCHW", "ms_per_iter": rms,
                                     "iters_per_s": 1e3 / rms, "tree": reference_models.reference_root()}
            del ref, j1, j2
        except Exception as e:
            res["reference_body"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    else:
        res["reference_body"] = {"error": "reference model files not staged (baseline/stage_reference.py)"}
    del net
    torch.cuda.empty_cache()
    return res


def perturbation_bench(dev, rank, world, global_batch, iters=5):
    """global_attacks/universal_perturbation.py:452-530, batched and pair-sharded: batch 64 at the reference's
    training resolution 256x640, n_step 10 (its default), learning rate 2e-3, sign steps; the (2,3,256,640)
    gradient (3.9 MB) is all-reduced every step.  iteration = clean forward + n_step x (forward + backward +
    all-reduce + update)."""
    import torch
    import torch.distributed as dist

    from understanding_flow_robustness_b200 import attack
    from understanding_flow_robustness_b200.harness import FlowNetCHarness
    torch.manual_seed(0)
    fmt = torch.channels_last
    net = FlowNetCHarness(fused_merge=True).to(dev).eval().to(memory_format=fmt)
    for q in net.parameters():
        q.requires_grad_(False)
    H, W, n_step, lr, eps = 256, 640, 10, 2e-3, 0.02
    idx = attack.shard_slice(global_batch, rank, world)
    g = torch.Generator(device=dev).manual_seed(200 + rank)
    i1 = torch.rand(len(idx), 3, H, W, device=dev, generator=g).contiguous(memory_format=fmt)
    i2 = torch.rand(len(idx), 3, H, W, device=dev, generator=g).contiguous(memory_format=fmt)
    target = torch.empty(len(idx), 2, H, W, device=dev)
    s_delta = torch.zeros(1, 2, 3, H, W, device=dev)

    def clean_fn(a, b):
        with torch.no_grad():
            return -net(a, b)

    def grad_fn(d):
        return attack.perturbation_gradient(net, i1, i2, d, target, global_batch)

    how = "cuda-graph replay of the per-rank work, all-reduce + update eager"
    try:
        g_clean = attack.GraphedGradient(clean_fn, [i1, i2])
        g_grad = attack.GraphedGradient(grad_fn, [s_delta])
    except Exception as e:
        torch.cuda.synchronize()
        how = f"eager (graph capture failed: {type(e).__name__}: {e})"[:200]
        g_clean = lambda: clean_fn(i1, i2)                # noqa: E731
        g_grad = grad_fn
    state = {"delta": torch.zeros_like(s_delta), "loss": None}

    def iteration():
        target.copy_(g_clean())
        d = state["delta"]
        for _ in range(n_step):
            packed = g_grad(d)
            if world > 1:
                dist.all_reduce(packed)
            d, state["loss"] = attack.apply_perturbation_step(d, packed, eps, lr)
        state["delta"] = d

    iteration()
    ms = _timed_iters(iteration, iters, dev, world)
    out = {"metric": "universal-perturbation iters/s", "value": 1e3 / ms, "unit": "iters/s", "ms_per_iter": ms,
           "iters_timed": iters, "pairs_per_s": global_batch * 1e3 / ms, "scaling": "strong",
           "global_batch": global_batch, "pairs_per_rank": len(idx), "n_gpus": world, "timed_loop": how,
           "allreduce_bytes": int(s_delta.numel() * 4 + 4), "allreduces_per_iter": n_step,
           "final_loss": float(state["loss"]), "delta_linf": float(state["delta"].abs().max()),
           "config": f"FlowNetC (fused merge block) {H}x{W}, delta (2,3,{H},{W}), n_step {n_step}, lr {lr}, eps {eps}, "
                     "I-FGSM sign steps, cosine loss"}
    del net, g_clean, g_grad
    torch.cuda.empty_cache()
    return out


def nccl_check_bench(dev, rank, world, global_batch):
    """All-reduced N-rank gradients of the bench's global batch == rank 0's single-process gradients."""
    import torch

    from understanding_flow_robustness_b200 import attack
    from understanding_flow_robustness_b200.harness import FlowNetCHarness
    torch.manual_seed(0)
    net = FlowNetCHarness(fused_merge=True).to(dev).eval()
    for q in net.parameters():
        q.requires_grad_(False)
    res = attack.nccl_value_check(net, dev, rank, world, global_pairs=global_batch, H=384, W=1280, p=100)
    # the same check in fp64 (sampler's fp64 kernels, torch-op placement) on a small batch: pins the plumbing to 1e-9
    net64 = FlowNetCHarness(fused_merge=False).to(dev).eval().double()
    r64 = attack.nccl_value_check(net64, dev, rank, world, global_pairs=2 * world, H=64, W=128, p=16, dtype=torch.float64)
    res["fp64"] = r64
    if rank == 0:
        res["ok"] = bool(res["ok"] and r64["ok"])
    del net, net64
    torch.cuda.empty_cache()
    return res


def reference_cuda_bench(dev):
    """The reference's own CUDA kernels, compiled unmodified for sm_100a (oracle/_ref, built by
    oracle/build_ref_cuda.py), timed on this GPU next to ours: the kernels to beat.  Never on the product path."""
    import math

    import torch
    import torch.nn.functional as F

    from oracle import build_ref_cuda
    from understanding_flow_robustness_b200 import backend
    out = {}

    def timeit(fn, n, warm=1):
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
    B, C, H, W, P = CFG["B"], CFG["C"], CFG["H"], CFG["W"], CFG["patch"]
    torch.manual_seed(0)
    a = torch.randn(B, C, H, W, device=dev)
    b = torch.randn(B, C, H, W, device=dev)
    g = torch.randn(B, P, P, H, W, device=dev)
    try:
        ref = build_ref_cuda.load_module("ref_sampler_cuda")
        rf = timeit(lambda: ref.forward(a, b, *Q), 3)
        rb = timeit(lambda: ref.backward(a, b, g, *Q), 2)
        of = timeit(lambda: backend.forward(a, b, *Q), 20, 3)
        ob = timeit(lambda: backend.backward(a, b, g, *Q), 20, 3)
        o, r = backend.forward(a, b, *Q), ref.forward(a, b, *Q)
        out["sampler"] = {"config": f"({B},{C},{H},{W}) patch 21 dilation_patch 2", "reference_fwd_ms": rf,
                          "reference_bwd_ms": rb, "ours_fwd_ms": of, "ours_bwd_ms": ob, "speedup_fwd": rf / of,
                          "speedup_bwd": rb / ob, "speedup_fwd_bwd": (rf + rb) / (of + ob),
                          "max_rel_diff_fwd": float((o - r).abs().max() / r.abs().max())}
        del o, r
    except Exception as e:
        out["sampler"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    del a, b, g
    try:
        ref_alt = build_ref_cuda.load_module("ref_alt_cuda_corr")
        from understanding_flow_robustness_b200 import AlternateCorrBlock, coords_grid
        c = RAFT_CFG
        Bq, Cq, Hq, Wq = c["B"], c["C"], c["H"], c["W"]
        f1 = torch.randn(Bq, Cq, Hq, Wq, device=dev)
        f2 = torch.randn(Bq, Cq, Hq, Wq, device=dev)
        coords = coords_grid(Bq, Hq, Wq, dev) + 3.0 * torch.randn(Bq, 2, Hq, Wq, device=dev)
        with torch.no_grad():
            f1n = f1.permute(0, 2, 3, 1).contiguous()
            pyr2 = [f2]
            for _ in range(3):
                pyr2.append(F.avg_pool2d(pyr2[-1], 2, stride=2))
            f2n = [q.permute(0, 2, 3, 1).contiguous() for q in pyr2]
            cn = coords.permute(0, 2, 3, 1)

            def ref_alt_iter():          # models/raft/corr.py:120-137
                outs = []
                for i in range(4):
                    ci = (cn / 2 ** i).reshape(Bq, 1, Hq, Wq, 2).contiguous()
                    (c_,) = ref_alt.forward(f1n, f2n[i], ci, 4)
                    outs.append(c_.squeeze(1))
                return torch.stack(outs, 1).reshape(Bq, -1, Hq, Wq) / math.sqrt(Cq)
            r_alt = timeit(ref_alt_iter, 2)
            alt = AlternateCorrBlock(f1, f2, 4, 4)
            o_alt = timeit(lambda: alt(coords), 5)
            pure = AlternateCorrBlock(f1, f2, 4, 4, dense_max_keys=0)
            o_pure = timeit(lambda: pure(coords), 5)
        out["alt_cuda_corr"] = {"config": f"B={Bq}, {Cq}x{Hq}x{Wq}, 4 levels, radius 4, one lookup", "reference_ms_per_iter": r_alt,
                                "ours_ms_per_iter": o_alt, "ours_pure_alt_kernel_ms_per_iter": o_pure,
                                "speedup": r_alt / o_alt, "speedup_pure_alt_kernel": r_alt / o_pure}
    except Exception as e:
        out["alt_cuda_corr"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    return out


def merge_bench(dev):
    """the FlowNetC merge block (correlate -> /C -> LeakyReLU -> cat, forward + backward
    w.r.t. both feature maps and the redir map) as one fused operator vs the same chain of torch ops on this
    package's unfused sampler.  Feature shape of BASELINE config 2."""
    import torch
    import torch.nn.functional as F

    from understanding_flow_robustness_b200 import correlate_merge, spatial_correlation_sample
    B, C, H, W = CFG["B"], CFG["C"], CFG["H"], CFG["W"]
    torch.manual_seed(0)
    a = torch.randn(B, C, H, W, device=dev, requires_grad=True)
    b = torch.randn(B, C, H, W, device=dev, requires_grad=True)
    r = torch.randn(B, 32, H, W, device=dev, requires_grad=True)
    g = torch.randn(B, 32 + 441, H, W, device=dev)

    def fused():
        correlate_merge(a, b, r, 21, 2, 0.1).backward(g)

    def unfused():
        out = spatial_correlation_sample(a, b, kernel_size=1, patch_size=21, stride=1, padding=0, dilation_patch=2)
        out = out.view(B, 441, H, W) / a.size(1)
        torch.cat((r, F.leaky_relu(out, 0.1)), 1).backward(g)

    def timeit(fn, n=20):
        for _ in range(5):
            fn()
            a.grad = b.grad = r.grad = None
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
            a.grad = b.grad = r.grad = None
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    tf, tu = timeit(fused), timeit(unfused)
    cv = B * 441 * H * W * 4
    return {"metric": "FlowNetC merge block fwd+bwd pairs/s", "fused_ms": tf, "unfused_ms": tu,
            "fused_pairs_per_s": B / (tf * 1e-3), "unfused_pairs_per_s": B / (tu * 1e-3),
            "config": f"features ({B},{C},{H},{W}), redir 32 channels, LeakyReLU 0.1; eager autograd calls, CUDA events",
            "cost_volume_passes": {"fused": "1 write (fwd) + 2 reads + 1 write + 2 reads (bwd)",
                                   "unfused": "7 (fwd) + 7 + 2 reads (bwd)", "bytes_per_pass": cv}}


def raft_bench(dev):
    """RAFT CorrBlock, BASELINE config 3: ms/iter := (pyramid build + 12 lookups) / 12, HBM roofline."""
    import math

    import torch

    from understanding_flow_robustness_b200 import AlternateCorrBlock, CorrBlock, coords_grid
    c = RAFT_CFG
    B, C, H, W = c["B"], c["C"], c["H"], c["W"]
    torch.manual_seed(0)
    f1 = torch.randn(B, C, H, W, device=dev)
    f2 = torch.randn(B, C, H, W, device=dev)
    coords = [coords_grid(B, H, W, dev) + 3.0 * torch.randn(B, 2, H, W, device=dev) for _ in range(c["iters"])]

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    blk = [None]

    def build():
        blk[0] = None   # RAFT holds one block per forward: the previous 1.25 GB pyramid is released first
        blk[0] = CorrBlock(f1, f2, c["levels"], c["radius"], precision="tf32")

    def lookups():
        for cc in coords:
            blk[0](cc)

    def graphed(fn):
        """fn's launches captured once and replayed (a 36 us lookup is launch-bound on a slow host)."""
        try:
            fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            return g.replay, "cuda-graph replay"
        except Exception as e:
            torch.cuda.synchronize()
            return fn, f"eager ({type(e).__name__})"

    with torch.no_grad():
        build_fn, how_build = graphed(build)
        build_ms = timed(build_fn, 5)
        build()   # an eager pyramid for the lookups below (the captured one lives in the graph's pool)
        look_fn, how_look = graphed(lookups)
        look_ms = timed(look_fn, 5) / c["iters"]
        # the same two numbers with the reference's row-major volume layout (layout="rowmajor")
        rm = [None]

        def build_rm():
            rm[0] = None
            rm[0] = CorrBlock(f1, f2, c["levels"], c["radius"], precision="tf32", layout="rowmajor")

        def lookups_rm():
            for cc in coords:
                rm[0](cc)
        blk[0] = None
        brm_fn, _ = graphed(build_rm)
        build_rm_ms = timed(brm_fn, 5)
        build_rm()
        lrm_fn, _ = graphed(lookups_rm)
        look_rm_ms = timed(lrm_fn, 5) / c["iters"]
        layout_mask = None
        rm[0] = None
        # opt-in fp16 storage of the two blocked levels (include/b200corr.h): same metric, half the volume bytes
        hf = [None]

        def build_hf():
            hf[0] = None
            hf[0] = CorrBlock(f1, f2, c["levels"], c["radius"], precision="tf32", storage="fp16")

        def lookups_hf():
            for cc in coords:
                hf[0](cc)
        bhf_fn, _ = graphed(build_hf)
        build_hf_ms = timed(bhf_fn, 5)
        build_hf()
        lhf_fn, _ = graphed(lookups_hf)
        look_hf_ms = timed(lhf_fn, 5) / c["iters"]
        hf_bytes = int(sum(v.numel() * v.element_size() for v in hf[0]._levels))
        hf[0] = None
        build()
        layout_mask = blk[0]._blocked
        fp32_bytes = int(sum(v.numel() * v.element_size() for v in blk[0]._levels))
        from understanding_flow_robustness_b200 import raft_corr
        alt = AlternateCorrBlock(f1, f2, c["levels"], c["radius"])
        alt_ms = timed(lambda: alt(coords[0]), 3)
        # the same block with every level on the alt_cuda_corr kernel (the reference's structure), with level 1
        # dense as well, and what building the block costs once per RAFT forward
        alt_rows = {}
        for name, dmk in (("pure_alt_cuda_corr", 0), ("dense_max_keys_2048", 2048)):
            a2 = AlternateCorrBlock(f1, f2, c["levels"], c["radius"], dense_max_keys=dmk)
            alt_rows[name + "_ms_per_iter"] = timed(lambda: a2(coords[0]), 3)
            del a2
        # alt_cuda_corr.backward at level 0 (the path's backward scatter): f1 gradient in registers, f2 gradient by atomics
        f1n = f1.permute(0, 2, 3, 1).contiguous()
        f2n = f2.permute(0, 2, 3, 1).contiguous()
        cn = coords[0].permute(0, 2, 3, 1).reshape(B, 1, H, W, 2).contiguous()
        cg = torch.randn(B, 1, 81, H, W, device=dev)
        ab_ms = timed(lambda: raft_corr.alt_cuda_corr.backward(f1n, f2n, cn, cg, c["radius"]), 5)
        af_ms = timed(lambda: raft_corr.alt_cuda_corr.forward(f1n, f2n, cn, c["radius"]), 5)
        alt_bytes = B * H * W * (81 * 4 + 2 * C * 4 + 100 * C * 4 * 2)     # grad read + f1 r/w + window read + RMW
        alt_rows["alt_cuda_corr_level0"] = {
            "forward_ms": af_ms, "backward_ms": ab_ms,
            "forward_tflops": 2.0 * B * H * W * 100 * C / (af_ms * 1e-3) / 1e12,
            "backward_tflops": 4.0 * B * H * W * 100 * C / (ab_ms * 1e-3) / 1e12,
            "backward_scatter_gbs": alt_bytes / (ab_ms * 1e-3) / 1e9,
            "what": "one level-0 call (B,1,H,W,2) coords; scatter bytes = every window pixel's C floats read once and "
                    "read-modify-written once (L2 atomics), + f1, f1_grad, corr_grad"}
        del f1n, f2n, cn, cg
        alt_rows["block_init_ms"] = timed(lambda: AlternateCorrBlock(f1, f2, c["levels"], c["radius"]), 3)
        alt_rows["dense_levels_from"] = alt._dense_from
        alt_rows["dense_bytes"] = int(sum(v.numel() * 4 for v in alt._dense))
        # extra rows of the same path (not part of ms/iter): lookup backward into a resident gradient pyramid,
        # and the volume in split-TF32 (fp32-level accuracy on the tensor cores)
        from understanding_flow_robustness_b200 import raft_corr
        glv = [v.new_zeros((v.shape[0], 1, H >> i, W >> i)) for i, v in enumerate(blk[0]._levels)]   # row-major gradient pyramid
        gout = torch.randn(B, c["levels"] * (2 * c["radius"] + 1) ** 2, H, W, device=dev)
        lbwd_fn, _ = graphed(lambda: raft_corr.lookup_backward(glv, coords[0], gout, c["radius"], H, W))
        lookup_bwd_ms = timed(lbwd_fn, 10)
        vbwd_fn, _ = graphed(lambda: raft_corr.volume_backward(glv, f1, f2, 1.0 / math.sqrt(C), "tf32"))
        vol_bwd_ms = timed(vbwd_fn, 5)
        del glv, gout
        # the lookup fused with the motion encoder's convc1 (1x1, 324 -> 256) + bias + ReLU
        # (update.py:104,111) against lookup -> cuDNN 1x1 convolution -> ReLU (torch's default allow_tf32)
        conv = torch.nn.Conv2d(c["levels"] * (2 * c["radius"] + 1) ** 2, 256, 1).to(dev)
        import torch.nn.functional as F
        fused_fn, how_fused = graphed(lambda: blk[0].lookup_convc1(coords[0], conv.weight, conv.bias, impl="fused"))
        fused_ms = timed(fused_fn, 10)
        pipe_fn, _ = graphed(lambda: blk[0].lookup_convc1(coords[0], conv.weight, conv.bias, impl="pipelined"))
        pipe_ms = timed(pipe_fn, 10)
        unf_fn, _ = graphed(lambda: F.relu(conv(blk[0](coords[0]))))
        unf_ms = timed(unf_fn, 10)
        look1_fn, _ = graphed(lambda: blk[0](coords[0]))
        look1_ms = timed(look1_fn, 10)
        fuse_row = {"pipelined_ms": pipe_ms, "fused_ms": fused_ms, "unfused_ms": unf_ms, "lookup_alone_ms": look1_ms,
                    "speedup_pipelined": unf_ms / pipe_ms, "speedup_fused": unf_ms / fused_ms,
                    "timed_loop": how_fused,
                    "bytes_not_moved_per_iter": 2 * B * 324 * H * W * 4,
                    "what": "CorrBlock.lookup_convc1 vs lookup kernel + cuDNN 1x1 conv (TF32 allowed) + ReLU.  fused: ONE "
                            "tcgen05 kernel (gather -> TF32 operand tile -> MMA over the 4 levels -> bias + ReLU), the "
                            "(B,324,H,W) lookup result is neither written nor read back; pipelined: lookup kernel + a "
                            "tcgen05 1x1-conv kernel (MN-major operand straight from the lookup result, which is still in "
                            "the L2), bias + ReLU in its epilogue"}
        del conv
        blk[0] = None
        x3_fn, _ = graphed(lambda: raft_corr.allpairs_pyramid(f1, f2, c["levels"], "tf32x3"))
        build_x3_ms = timed(x3_fn, 3)
    HW = H * W
    vol_bytes = B * (4 * HW * HW * (1 + 0.25 + 1 / 16 + 1 / 64) + 2 * C * HW * 4)
    look_bytes = B * (324 * HW * 4 + 4 * HW * 100 * 4)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    tr = {}
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "raft_traffic.json")))
    except Exception:
        pass

    def traffic(kernel):
        r_, w_ = tr.get(kernel + "_dram_read_bytes"), tr.get(kernel + "_dram_write_bytes")
        return None if r_ is None else {"dram_read": r_, "dram_write": w_, "source": tr.get("source"),
                                        "note": "writes still dirty in the 126 MB L2 when the launch ends are not counted by ncu"}
    return {"metric": "RAFT corr+lookup ms/iter", "ms_per_iter": (build_ms + c["iters"] * look_ms) / c["iters"],
            "build_ms": build_ms, "lookup_ms": look_ms,
            "volume_layout": {"blocked_levels_mask": layout_mask,
                              "what": "levels in the mask are stored as 8x8 tiles of 64 floats (include/b200corr.h); "
                                      "corr_pyramid / get_corr_pyramid() convert to the reference's row-major view on demand",
                              "rowmajor_build_ms": build_rm_ms, "rowmajor_lookup_ms": look_rm_ms,
                              "rowmajor_ms_per_iter": (build_rm_ms + c["iters"] * look_rm_ms) / c["iters"]},
            "fp16_storage": {"build_ms": build_hf_ms, "lookup_ms": look_hf_ms,
                             "ms_per_iter": (build_hf_ms + c["iters"] * look_hf_ms) / c["iters"],
                             "volume_bytes": hf_bytes, "volume_bytes_fp32": fp32_bytes,
                             "what": "CorrBlock(storage='fp16'), opt-in: levels 0-1 (94 % of the volume) stored as fp16 tiles, "
                                     "each value rounded once from the fp32 accumulator (relative 2^-11, saturating); NOT the "
                                     "configuration of ms_per_iter / roofline above, which keep the reference's fp32 volume"},
            "alt_corr_ms_per_iter": alt_ms,
            "alt_corr": alt_rows, "lookup_backward_ms": lookup_bwd_ms, "build_tf32x3_ms": build_x3_ms,
            "lookup_convc1": fuse_row,
            "volume_backward_ms": vol_bwd_ms,
            "corrblock_backward_12_lookups_ms": 12 * lookup_bwd_ms + vol_bwd_ms,
            "config": f"B={B}, {C}x{H}x{W}, {c['levels']} levels, radius {c['radius']}, {c['iters']} lookups, TF32 volume",
            "timed_loop": {"build": how_build, "lookups": how_look},
            "roofline_build": {"bound": "hbm", "achieved": vol_bytes / (build_ms * 1e-3) / 1e9, "peak": hbm,
                               "unit": "GB/s", "frac": vol_bytes / (build_ms * 1e-3) / 1e9 / hbm, "peak_source": src,
                               "bytes": "volume + 3 pooled levels written once + features read once",
                               "traffic": traffic("allpairs_tc_kernel")},
            "roofline_lookup": {"bound": "hbm", "achieved": look_bytes / (look_ms * 1e-3) / 1e9, "peak": hbm,
                                "unit": "GB/s", "frac": look_bytes / (look_ms * 1e-3) / 1e9 / hbm, "peak_source": src,
                                "bytes": "324-channel output written + 4 levels x 10x10 window read per query",
                                "traffic": traffic("lookup_fwd_kernel"),
                                "what_bounds_it": "not the gather: with the window loads switched off the kernel takes the same time, "
                                                  "and the pieces (CTA prologue 7.4, tap tables 5.2, parking 2.1, loads 6, sampling "
                                                  "6.6, stores 3.4 us) add up -- a per-CTA latency chain at 4 CTAs per SM "
                                                  "(DESIGN.md 2.4, 2.6)"},
            "roofline": {"bound": "hbm", "what": "(build + 12 lookups) / 12 against the algorithmic bytes of both at the HBM copy peak",
                         "bound_ms_per_iter": (vol_bytes + c["iters"] * look_bytes) / (hbm * 1e9) * 1e3 / c["iters"],
                         "frac": (vol_bytes + c["iters"] * look_bytes) / (hbm * 1e9) * 1e3 / (build_ms + c["iters"] * look_ms)}}


if __name__ == "__main__":
    main()
