"""SM probes (run on the GPU box): LDS.128 multicast cost, FP32 FFMA peak, first sampler timings."""
import ctypes
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import subprocess

from understanding_flow_robustness_b200 import _lib, spatial_correlation_sample

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_probes():
    """libb200probes.so: the SM probes + the product library's error / launch-count plumbing (common.cu)."""
    out = os.path.join(ROOT, "scripts", "probes", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libb200probes.so")
    csrc = os.path.join(ROOT, "understanding_flow_robustness_b200", "csrc")
    if not os.path.exists(so):
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                               "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-I", csrc,
                               os.path.join(ROOT, "scripts", "probes", "sm_probes.cu"), os.path.join(csrc, "common.cu"),
                               "-o", so, "-lcudart"])
    P = ctypes.CDLL(so)
    for name, args in (("b200corr_probe_lds", [ctypes.c_int] * 3 + [ctypes.POINTER(ctypes.c_float), ctypes.c_void_p]),
                       ("b200corr_probe_ffma2_peak", [ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_void_p]),
                       ("b200corr_probe_ffma_toeplitz", [ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_void_p])):
        getattr(P, name).argtypes = args
        getattr(P, name).restype = ctypes.c_int
    return P


L = _lib.lib()
PR = build_probes()
res = {}
st = _lib.current_stream(torch.device("cuda:0"))
v = ctypes.c_float()
for warps in (16,):
    for pat in range(11):
        _lib.check(PR.b200corr_probe_lds(pat, warps, 2000, ctypes.byref(v), st), "probe_lds")
        res[f"lds_p{pat}_w{warps}"] = round(v.value, 3)
_lib.check(L.b200corr_measure_fp32_peak(20000, ctypes.byref(v), st), "fp32 peak")
res["fp32_peak_tflops"] = round(v.value, 2)
_lib.check(PR.b200corr_probe_ffma2_peak(4000, ctypes.byref(v), st), "ffma2 peak")
res["ffma2_peak_tflops"] = round(v.value, 2)
_lib.check(PR.b200corr_probe_ffma_toeplitz(2000, ctypes.byref(v), st), "toeplitz")
res["ffma_toeplitz_tflops"] = round(v.value, 2)
print(json.dumps(res, indent=1))


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


if "--sampler" in sys.argv:
    for B in (1, 8):
        a = torch.randn(B, 256, 48, 160, device="cuda", requires_grad=True)
        b = torch.randn(B, 256, 48, 160, device="cuda", requires_grad=True)
        g = torch.randn(B, 21, 21, 48, 160, device="cuda")
        from understanding_flow_robustness_b200 import backend
        q = (1, 1, 21, 21, 0, 0, 1, 1, 2, 2, 1, 1)
        tf = timeit(lambda: backend.forward(a.detach(), b.detach(), *q))
        tb = timeit(lambda: backend.backward(a.detach(), b.detach(), g, *q))
        inb = 256 * 788 * 3140 * 2 * B
        print(json.dumps({"B": B, "fwd_ms": round(tf, 4), "bwd_ms": round(tb, 4),
                          "fwd_inbounds_tflops": round(inb / tf / 1e9, 2),
                          "bwd_inbounds_tflops": round(2 * inb / tb / 1e9, 2),
                          "pairs_per_s": round(B / (tf + tb) * 1e3, 1)}))
