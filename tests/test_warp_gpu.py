"""PWC-Net warp() kernel against the reference function (models/PWCNet.py:164-204).

Checked against golden vectors the reference method itself produced (tests/golden/warp_*.npz,
oracle/make_golden_warp.py) and, on fresh inputs, against oracle/warp_oracle.py -- that method restated with its
`.cuda()` calls made device-agnostic (pinned to the same golden vectors in tests/test_oracle_cpu.py), run on the
same GPU in fp32.
Tolerance: forward 1e-6 of max|ref| (same arithmetic; the two may differ in FMA contraction), gradients 1e-5
(atomics: summation order).
"""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _reference_warp(x, flo):
    from oracle import warp_oracle
    return warp_oracle.warp(x, flo)


WARP_GOLDEN = sorted(glob.glob(os.path.join(GOLDEN, "warp_*.npz")))


@pytest.mark.parametrize("path", WARP_GOLDEN, ids=[os.path.basename(p)[:-4] for p in WARP_GOLDEN])
def test_warp_golden_vectors(path):
    """Vectors produced by the reference method itself on CPU (oracle/make_golden_warp.py).  torch's CPU `x / 19`
    is a true division, its CUDA kernel (and this one) multiplies by fl(1/19): coordinates differ by an ulp, hence
    1e-5 / 1e-4 here; the same-device comparison below is exact."""
    from understanding_flow_robustness_b200 import warp
    z = np.load(path)
    x = torch.from_numpy(z["x"]).cuda().requires_grad_()
    flo = torch.from_numpy(z["flo"]).cuda().requires_grad_()
    out = warp(x, flo)
    gx, gf = torch.autograd.grad(out, (x, flo), torch.from_numpy(z["gout"]).cuda())
    assert np.abs(out.detach().cpu().numpy() - z["out"]).max() <= 1e-5 * np.abs(z["out"]).max()
    assert np.abs(gx.cpu().numpy() - z["gx"]).max() <= 1e-4 * np.abs(z["gx"]).max()
    assert np.abs(gf.cpu().numpy() - z["gflo"]).max() <= 1e-4 * np.abs(z["gflo"]).max()


CASES = [(2, 32, 24, 80, 3.0), (1, 196, 6, 20, 1.5), (2, 5, 9, 13, 8.0), (1, 64, 48, 160, 30.0), (1, 3, 1, 1, 0.5),
         (2, 16, 12, 40, 0.0)]


@pytest.mark.parametrize("case", CASES, ids=str)
def test_warp_forward_backward_vs_reference_function(case):
    from understanding_flow_robustness_b200 import warp
    B, C, H, W, sigma = case
    torch.manual_seed(B * 100 + C + H + W)
    x = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
    flo = (sigma * torch.randn(B, 2, H, W, device="cuda")).requires_grad_()
    g = torch.randn(B, C, H, W, device="cuda")
    out = warp(x, flo)
    gx, gf = torch.autograd.grad(out, (x, flo), g)
    ref = _reference_warp(x, flo)
    rx, rf = torch.autograd.grad(ref, (x, flo), g)
    s = float(ref.abs().max()) + 1e-30
    assert out.shape == ref.shape
    assert torch.equal(out, ref)                      # same op sequence as torch's CUDA kernels: bit-identical
    assert float((gx - rx).abs().max()) <= 1e-5 * (float(rx.abs().max()) + 1e-30)
    assert float((gf - rf).abs().max()) <= 1e-5 * (float(rf.abs().max()) + 1e-30)


def test_warp_zero_flow_and_out_of_image():
    from understanding_flow_robustness_b200 import warp
    x = torch.randn(1, 4, 10, 12, device="cuda")
    # zero flow is NOT the identity: the reference normalises for align_corners=True and samples with False
    z = warp(x, torch.zeros(1, 2, 10, 12, device="cuda"))
    assert torch.allclose(z, _reference_warp(x, torch.zeros(1, 2, 10, 12, device="cuda")), rtol=0, atol=1e-6)
    far = torch.full((1, 2, 10, 12), 1000.0, device="cuda")
    assert float(warp(x, far).abs().max()) == 0.0                        # everything masked
    nan = torch.full((1, 2, 10, 12), float("nan"), device="cuda")
    assert warp(x, nan).shape == x.shape                                  # no fault on non-finite flow


def test_warp_feeds_the_correlation_like_pwcnet():
    """PWCNet.py:293-296: corr5 = leakyRELU(corr(c15, warp(c25, up_flow6 * 0.625))), fused tail included."""
    from understanding_flow_robustness_b200 import correlate_merge, spatial_correlation_sample, warp
    torch.manual_seed(7)
    c1 = torch.randn(2, 128, 12, 40, device="cuda", requires_grad=True)
    c2 = torch.randn(2, 128, 12, 40, device="cuda", requires_grad=True)
    up_flow = torch.randn(2, 2, 12, 40, device="cuda", requires_grad=True)
    up_feat = torch.randn(2, 2, 12, 40, device="cuda")
    x = correlate_merge(c1, warp(c2, up_flow * 0.625), None, 9, 1, 0.1, after=(c1, up_flow, up_feat))
    corr = spatial_correlation_sample(c1, _reference_warp(c2, up_flow * 0.625), kernel_size=1, patch_size=9, stride=1)
    ref = torch.cat((torch.nn.functional.leaky_relu(corr.view(2, 81, 12, 40) / 128, 0.1), c1, up_flow, up_feat), 1)
    g = torch.randn_like(ref)
    ga = torch.autograd.grad(x, (c1, c2, up_flow), g)
    gb = torch.autograd.grad(ref, (c1, c2, up_flow), g)
    assert float((x - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    for a, b in zip(ga, gb):
        assert float((a - b).abs().max()) <= 1e-4 * float(b.abs().max())
