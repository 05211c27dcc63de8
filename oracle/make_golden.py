"""Generate tests/golden/*.npz from the REFERENCE itself (run in the build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference cannot travel to the GPU box, so the vectors it
produces are committed as small fixtures together with this script:

  sampler_*.npz : inputs + outputs of the reference CPU extension (oracle/_ref, compiled from
                  correlation.cpp / correlation_sampler.cpp in place) driven through the reference's
                  own Python wrapper spatial_correlation_sampler.py (autograd Function, forward and
                  backward of a seeded upstream gradient).  Parameter grid = the reference's
                  check.py:76-89 and grad_check.py:9-24 defaults, the FlowNetC (21/2) and PWC-Net
                  (9/1) call-site configurations, an even patch size and an anisotropic case.
  raft_*.npz    : CorrBlock pyramid + lookup from the reference's models/raft/corr.py imported with
                  models/__init__.py bypassed, CPU tensors.

Usage: python oracle/make_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
REF = "/root/reference"

SAMPLER_CASES = {
    # name: (B, C, H, W, dtype, kernel, patch, stride, pad, dil, dil_patch)
    "check_default": (1, 10, 10, 10, "float64", 3, 3, 2, 5, 2, 2),   # check.py:76-89
    "gradcheck":     (2, 2, 10, 10, "float64", 3, 3, 2, 1, 2, 2),    # grad_check.py:9-24
    "p5_dp2":        (2, 3, 9, 10, "float32", 1, 5, 1, 0, 1, 2),
    "p4_dp2_even":   (2, 3, 9, 10, "float32", 1, 4, 1, 0, 1, 2),
    "k2_p3_pad1":    (2, 3, 9, 10, "float32", 2, 3, 1, 1, 1, 1),
    "pwc_p9":        (2, 5, 9, 10, "float32", 1, 9, 1, 0, 1, 1),     # PWCNet.py:43-45
    "flownetc_p21":  (1, 8, 12, 16, "float32", 1, 21, 1, 0, 1, 2),   # submodules.py:124-138
    "flownetc_odd":  (2, 4, 11, 13, "float32", 1, 21, 1, 0, 1, 2),   # odd H/W, sub-lattices ragged
    "aniso":         (1, 3, 8, 11, "float32", (1, 3), (3, 5), (1, 2), (0, 1), (1, 2), (2, 1)),
}

RAFT_CASES = {
    # name: (B, C, H, W, levels, radius, coord noise sigma)
    "r4_16x16": (1, 8, 16, 16, 4, 4, 3.0),
    "r2_10x14": (2, 4, 10, 14, 3, 2, 2.0),
    "r4_16x16_far": (1, 8, 16, 16, 4, 4, 40.0),
}


def import_reference():
    from oracle import build_ref

    backend = build_ref.load_backend()
    sys.modules["spatial_correlation_sampler_backend"] = backend
    sys.path.insert(0, os.path.join(
        REF, "models/Pytorch-Correlation-extension/Correlation_Module"))
    import spatial_correlation_sampler as scs  # the reference's own wrapper package

    models = types.ModuleType("models")
    models.__path__ = [os.path.join(REF, "models")]
    sys.modules["models"] = models
    import importlib

    corr = importlib.import_module("models.raft.corr")
    utils = importlib.import_module("models.raft.utils.utils")
    return scs, corr, utils


def main():
    scs, corr, utils = import_reference()
    os.makedirs(GOLD, exist_ok=True)
    for i, (name, (B, C, H, W, dt, k, p, s, pad, dil, dp)) in enumerate(SAMPLER_CASES.items()):
        g = torch.Generator().manual_seed(1000 + i)
        dtype = getattr(torch, dt)
        in1 = torch.randn(B, C, H, W, generator=g, dtype=dtype).requires_grad_()
        in2 = torch.randn(B, C, H, W, generator=g, dtype=dtype).requires_grad_()
        out = scs.spatial_correlation_sample(in1, in2, k, p, s, pad, dil, dp)
        gout = torch.randn(out.shape, generator=g, dtype=dtype)
        out.backward(gout)
        np.savez(os.path.join(GOLD, f"sampler_{name}.npz"),
                 in1=in1.detach().numpy(), in2=in2.detach().numpy(), out=out.detach().numpy(),
                 gout=gout.numpy(), gin1=in1.grad.numpy(), gin2=in2.grad.numpy(),
                 params=np.array([np.broadcast_to(np.array(v), (2,)) for v in (k, p, s, pad, dil, dp)]))
        print("sampler", name, tuple(out.shape))
    for i, (name, (B, C, H, W, L, r, sig)) in enumerate(RAFT_CASES.items()):
        g = torch.Generator().manual_seed(2000 + i)
        f1 = torch.randn(B, C, H, W, generator=g)
        f2 = torch.randn(B, C, H, W, generator=g)
        coords = utils.coords_grid(B, H, W) + sig * torch.randn(B, 2, H, W, generator=g)
        blk = corr.CorrBlock(f1, f2, num_levels=L, radius=r)
        out = blk(coords)
        pyr = {f"pyr{l}": v.numpy() for l, v in enumerate(blk.get_corr_pyramid())}
        np.savez(os.path.join(GOLD, f"raft_{name}.npz"), f1=f1.numpy(), f2=f2.numpy(),
                 coords=coords.numpy(), out=out.numpy(), levels=np.array(L), radius=np.array(r),
                 **pyr)
        print("raft", name, tuple(out.shape))


if __name__ == "__main__":
    main()
