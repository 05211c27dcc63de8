"""Parity of the CUDA spatial correlation sampler (through the C-ABI) against the CPU oracle.

Mirrors the reference's own checks: check.py:8-59 (forward and gradients of out.sum(), CPU vs CUDA)
and grad_check.py:9-55 (fp64 gradcheck), on the reference-generated golden vectors and on fresh
seeded inputs.  Tolerance: fp32 max|diff| <= 1e-5 * max|ref| (BASELINE.json north_star); fp64 1e-12.
"""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

SAMPLER = sorted(glob.glob(os.path.join(GOLDEN, "sampler_*.npz")))


def _params(z):
    p = z["params"]
    return dict(kernel_size=tuple(int(v) for v in p[0]), patch_size=tuple(int(v) for v in p[1]),
                stride=tuple(int(v) for v in p[2]), padding=tuple(int(v) for v in p[3]),
                dilation=tuple(int(v) for v in p[4]), dilation_patch=tuple(int(v) for v in p[5]))


def _rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _run(in1, in2, gout, **kw):
    from understanding_flow_robustness_b200 import spatial_correlation_sample
    a = torch.from_numpy(in1).cuda().requires_grad_()
    b = torch.from_numpy(in2).cuda().requires_grad_()
    out = spatial_correlation_sample(a, b, **kw)
    out.backward(torch.from_numpy(gout).cuda())
    torch.cuda.synchronize()
    return out.detach().cpu().numpy(), a.grad.cpu().numpy(), b.grad.cpu().numpy()


@pytest.mark.parametrize("path", SAMPLER, ids=[os.path.basename(p)[:-4] for p in SAMPLER])
def test_golden_vectors(path):
    z = np.load(path)
    out, g1, g2 = _run(z["in1"], z["in2"], z["gout"], **_params(z))
    tol = 1e-12 if z["in1"].dtype == np.float64 else 1e-5
    assert out.shape == z["out"].shape
    assert _rel(out, z["out"]) <= tol
    assert _rel(g1, z["gin1"]) <= tol
    assert _rel(g2, z["gin2"]) <= tol


def test_generic_kernels_are_bit_exact_vs_reference_golden():
    """sampler_generic.cu accumulates in the reference's order without FMA contraction."""
    for path in SAMPLER:
        z = np.load(path)
        kw = _params(z)
        out, g1, g2 = _run(z["in1"], z["in2"], z["gout"], **kw)
        from understanding_flow_robustness_b200 import _lib
        B, C, H, W = z["in1"].shape
        q = (*kw["kernel_size"], *kw["patch_size"], *kw["padding"], *kw["dilation"],
             *kw["dilation_patch"], *kw["stride"])
        dt = 0 if z["in1"].dtype == np.float32 else 1
        if _lib.lib().b200corr_sampler_uses_fast_path(B, C, H, W, *q, dt, 0):
            continue
        np.testing.assert_array_equal(out, z["out"], err_msg=path)


FAST_CASES = [
    # B, C, H, W, patch, dilation_patch        (C % 128 == 0 -> fast backward too)
    (1, 8, 12, 16, 21, 2),
    (2, 16, 11, 20, 21, 2),      # odd H: ragged row parity classes
    (1, 128, 9, 36, 21, 2),      # fast backward, W not a multiple of the 32-pixel unit
    (2, 128, 48, 160, 21, 2),    # FlowNetC shape, half the channels
    (1, 256, 23, 44, 21, 2),
    (1, 128, 10, 24, 9, 1),      # PWC-Net structure
    (2, 8, 7, 12, 9, 1),
    (1, 128, 3, 8, 21, 2),       # image smaller than the patch radius
    (2, 96, 12, 40, 9, 1),       # PWC-Net level: C % 32 == 0 backward (2 channels per thread)
    (1, 32, 24, 80, 9, 1),
    (1, 64, 13, 24, 21, 2),      # patch 21 with the 32-channel unit
    (2, 196, 6, 20, 9, 1),       # PWC-Net level 6: C % 8 != 0 -> 4-D tensor maps, zero-filled channel tail
    (3, 5, 9, 12, 9, 1),         # fewer channels than one chunk, several samples (a 3-D map would read the next sample)
    (2, 37, 10, 16, 21, 2),      # odd channel count on the patch-21 kernels
    (2, 130, 7, 12, 21, 2),      # C % 128 != 0 and C % 32 != 0: 32-channel backward units with a tail
    (2, 256, 55, 128, 21, 2),    # BASELINE config 5: Sintel 436x1024 feature shape (odd H), full channel count
    (2, 256, 68, 120, 21, 2),    # BASELINE config 5: FlyingThings 540x960 feature shape (W % 32 != 0)
]


@pytest.mark.parametrize("case", FAST_CASES, ids=[str(c) for c in FAST_CASES])
def test_fast_path_vs_oracle(case):
    from oracle import sampler_oracle
    from understanding_flow_robustness_b200 import _lib
    B, C, H, W, P, dp = case
    rng = np.random.default_rng(hash(case) & 0xFFFF)
    in1 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    in2 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    gout = rng.standard_normal((B, P, P, H, W)).astype(np.float32)
    assert _lib.lib().b200corr_sampler_uses_fast_path(B, C, H, W, 1, 1, P, P, 0, 0, 1, 1, dp, dp, 1, 1, 0, 0) == 1
    out, g1, g2 = _run(in1, in2, gout, kernel_size=1, patch_size=P, dilation_patch=dp)
    ref = sampler_oracle.forward(in1, in2, 1, P, 1, 0, 1, dp)
    r1, r2 = sampler_oracle.backward(in1, in2, gout, 1, P, 1, 0, 1, dp)
    assert _rel(out, ref) <= 1e-5
    assert _rel(g1, r1) <= 1e-5
    assert _rel(g2, r2) <= 1e-5


def test_fast_path_row_dilation_differs_from_column_dilation():
    """dilation_patch=(3, 2): three row parity classes, column structure of the 21x2 kernels."""
    from oracle import sampler_oracle
    rng = np.random.default_rng(5)
    in1 = rng.standard_normal((1, 8, 14, 16)).astype(np.float32)
    in2 = rng.standard_normal((1, 8, 14, 16)).astype(np.float32)
    gout = rng.standard_normal((1, 21, 21, 14, 16)).astype(np.float32)
    out, g1, g2 = _run(in1, in2, gout, kernel_size=1, patch_size=21, dilation_patch=(3, 2))
    ref = sampler_oracle.forward(in1, in2, 1, 21, 1, 0, 1, (3, 2))
    r1, r2 = sampler_oracle.backward(in1, in2, gout, 1, 21, 1, 0, 1, (3, 2))
    assert _rel(out, ref) <= 1e-5 and _rel(g1, r1) <= 1e-5 and _rel(g2, r2) <= 1e-5


def test_full_size_flownetc_properties():
    """BASELINE config 2 shape (8,256,48,160): size-independent properties + spot checks vs oracle."""
    from oracle import sampler_oracle
    from understanding_flow_robustness_b200 import spatial_correlation_sample
    torch.manual_seed(0)
    a = torch.randn(8, 256, 48, 160, device="cuda")
    b = torch.randn(8, 256, 48, 160, device="cuda")
    out = spatial_correlation_sample(a, b, 1, 21, 1, 0, 1, 2)
    assert out.shape == (8, 21, 21, 48, 160)
    # zero displacement plane == per-pixel dot product
    dot = (a * b).sum(1)
    assert _rel(out[:, 10, 10].cpu().numpy(), dot.cpu().numpy()) <= 1e-5
    # bilinearity in input1
    out2 = spatial_correlation_sample(2.0 * a, b, 1, 21, 1, 0, 1, 2)
    assert _rel(out2.cpu().numpy(), (2.0 * out).cpu().numpy()) <= 1e-6
    # swap symmetry: corr(a,b)[dy,dx][h,w] == corr(b,a)[-dy,-dx][h+dy,w+dx]
    outs = spatial_correlation_sample(b, a, 1, 21, 1, 0, 1, 2)
    assert _rel(out[:, 12, 7, 0:40, 6:160].cpu().numpy(), outs[:, 8, 13, 4:44, 0:154].cpu().numpy()) <= 1e-5
    # out-of-image displacements are exactly zero
    assert float(out[:, 0, :, 0:20].abs().max()) == 0.0
    assert float(out[:, :, 20, :, 140:].abs().max()) == 0.0
    # one sample against the oracle (forward), and the adjoint identity for the backward
    ref = sampler_oracle.forward(a[3:4].cpu().numpy(), b[3:4].cpu().numpy(), 1, 21, 1, 0, 1, 2)
    assert _rel(out[3:4].cpu().numpy(), ref) <= 1e-5
    a.requires_grad_()
    b.requires_grad_()
    g = torch.randn_like(out)
    o = spatial_correlation_sample(a, b, 1, 21, 1, 0, 1, 2)
    o.backward(g)
    # <g, J_a da> == <J_a^T g, da> with da := a (forward is linear in a): <g, out> == <grad_a, a>
    lhs = float((g.double() * o.detach().double()).sum())
    assert abs(float((a.grad.double() * a.detach().double()).sum()) - lhs) <= 1e-5 * abs(lhs) + 1e-2
    assert abs(float((b.grad.double() * b.detach().double()).sum()) - lhs) <= 1e-5 * abs(lhs) + 1e-2
    r1, r2 = sampler_oracle.backward(a[5:6].detach().cpu().numpy(), b[5:6].detach().cpu().numpy(),
                                     g[5:6].cpu().numpy(), 1, 21, 1, 0, 1, 2)
    assert _rel(a.grad[5:6].cpu().numpy(), r1) <= 1e-5
    assert _rel(b.grad[5:6].cpu().numpy(), r2) <= 1e-5


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
def test_half_precision_storage(dtype):
    """The reference's CUDA dispatch accepts at::Half (correlation_cuda_kernel.cu:262,297).  Here fp16 /
    bf16 tensors keep their type end to end, the arithmetic is fp32: the result must equal the fp32
    operator applied to the same (already rounded) inputs up to ONE rounding to the storage type."""
    from understanding_flow_robustness_b200 import spatial_correlation_sample
    torch.manual_seed(5)
    kw = dict(kernel_size=1, patch_size=21, stride=1, padding=0, dilation=1, dilation_patch=2)
    for shape, kw_ in [((2, 32, 12, 20), kw), ((1, 6, 9, 10), dict(kernel_size=3, patch_size=3, stride=2, padding=1, dilation=1, dilation_patch=2))]:
        a = torch.randn(*shape, device="cuda").to(dtype).requires_grad_()
        b = torch.randn(*shape, device="cuda").to(dtype).requires_grad_()
        out = spatial_correlation_sample(a, b, **kw_)
        assert out.dtype == dtype
        g = torch.randn(out.shape, device="cuda").to(dtype)
        out.backward(g)
        a32 = a.detach().float().requires_grad_()
        b32 = b.detach().float().requires_grad_()
        ref = spatial_correlation_sample(a32, b32, **kw_)
        ref.backward(g.float())
        # one rounding to the storage type (fp16: 2^-11, bf16: 2^-8 relative) on top of fp32 arithmetic
        eps = 2.0 ** -11 if dtype == torch.float16 else 2.0 ** -8
        for got, want in ((out.detach(), ref.detach()), (a.grad, a32.grad), (b.grad, b32.grad)):
            assert got.dtype == dtype
            err = (got.float() - want).abs()
            assert bool((err <= 1.01 * eps * want.abs() + 1e-5 * float(want.abs().max())).all()), float(err.max())


def test_gradcheck_fp64():
    """grad_check.py:9-55 equivalent."""
    from understanding_flow_robustness_b200 import SpatialCorrelationSampler
    torch.manual_seed(0)
    a = torch.randn(2, 2, 10, 10, dtype=torch.float64, device="cuda", requires_grad=True)
    b = torch.randn(2, 2, 10, 10, dtype=torch.float64, device="cuda", requires_grad=True)
    m = SpatialCorrelationSampler(3, 3, 2, 1, 2, 2)
    assert torch.autograd.gradcheck(m, [a, b])


def test_error_behaviour():
    from understanding_flow_robustness_b200 import spatial_correlation_sample
    a = torch.randn(1, 4, 8, 8)
    with pytest.raises(RuntimeError):
        spatial_correlation_sample(a, a, 1, 3)            # CPU tensor: no CPU path in this build
    c = torch.randn(1, 4, 8, 8, device="cuda")
    with pytest.raises(RuntimeError):
        spatial_correlation_sample(c.transpose(2, 3), c, 1, 3)   # non-contiguous (CHECK_CONTIGUOUS)
    with pytest.raises(RuntimeError):
        spatial_correlation_sample(c, c[:, :2].contiguous(), 1, 3)   # shape mismatch
    out = spatial_correlation_sample(c[:0], c[:0], 1, 3)  # empty batch
    assert out.shape == (0, 3, 3, 8, 8)


def test_second_device_in_the_same_process():
    """check.py:62-73 runs the op on every visible GPU from one process."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from oracle import sampler_oracle
    from understanding_flow_robustness_b200 import CorrBlock, coords_grid, spatial_correlation_sample
    rng = np.random.default_rng(3)
    in1 = rng.standard_normal((1, 128, 12, 32)).astype(np.float32)
    in2 = rng.standard_normal((1, 128, 12, 32)).astype(np.float32)
    ref = sampler_oracle.forward(in1, in2, 1, 21, 1, 0, 1, 2)
    for dev in ("cuda:0", "cuda:1"):
        a = torch.from_numpy(in1).to(dev).requires_grad_()
        b = torch.from_numpy(in2).to(dev)
        out = spatial_correlation_sample(a, b, 1, 21, 1, 0, 1, 2)
        out.sum().backward()
        assert out.device == a.device
        assert _rel(out.detach().cpu().numpy(), ref) <= 1e-5
        f = torch.randn(1, 32, 8, 32, device=dev)
        blk = CorrBlock(f, f, 2, 2)
        assert blk(coords_grid(1, 8, 32, dev)).device == f.device
