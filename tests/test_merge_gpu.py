"""Fused FlowNetC merge block against the chain it replaces.

Reference chain: models/submodules.py:124-138 (`correlate`: sampler, view, `/ C`) ->
models/FlowNetC.py:138 (`corr_activation` = LeakyReLU(0.1)) -> :147 (`torch.cat((out_conv_redir, out_corr), 1)`).
Checked (a) against the CPU oracle of the sampler with the chain applied in numpy (tolerance 1e-5 relative,
the sampler's own bound) and (b) against the same chain run with torch ops on this package's unfused operator,
where forward and backward must agree bit for bit for power-of-two channel counts (same kernel arithmetic,
same elementwise arithmetic) and to 1 ulp otherwise.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _chain_torch(a, b, redir, patch, dp, slope):
    from understanding_flow_robustness_b200 import spatial_correlation_sample
    out = spatial_correlation_sample(a, b, kernel_size=1, patch_size=patch, stride=1, padding=0, dilation_patch=dp)
    B, ph, pw, h, w = out.shape
    out = out.view(B, ph * pw, h, w) / a.size(1)
    return torch.cat((redir, F.leaky_relu(out, slope)), 1)


def _rel(x, y):
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    return np.abs(x - y).max() / max(np.abs(y).max(), 1e-30)


CASES = [
    # B, C, H, W, patch, dp, c_redir
    (2, 64, 12, 32, 21, 2, 32),      # FlowNetC structure, wide (256-bit) stores
    (1, 128, 9, 20, 21, 2, 32),      # W % 8 != 0: 128-bit store path, odd H (ragged parity classes)
    (2, 32, 10, 24, 9, 1, 8),        # PWC-Net structure
    (1, 24, 7, 16, 21, 2, 3),        # C not a power of two, odd slice offset
    (3, 256, 6, 8, 21, 2, 0),        # no redir channels at all
]


@pytest.mark.parametrize("case", CASES, ids=str)
def test_merge_forward_backward_vs_unfused_chain_and_oracle(case):
    from oracle import sampler_oracle
    from understanding_flow_robustness_b200 import correlate_merge
    B, C, H, W, patch, dp, cr = case
    slope = 0.1
    rng = np.random.default_rng(B * 1000 + C + H + W)
    in1 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    in2 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    red = rng.standard_normal((B, cr, H, W)).astype(np.float32)
    gm = rng.standard_normal((B, cr + patch * patch, H, W)).astype(np.float32)

    def run(fn):
        a = torch.from_numpy(in1).cuda().requires_grad_()
        b = torch.from_numpy(in2).cuda().requires_grad_()
        r = torch.from_numpy(red).cuda().requires_grad_()
        out = fn(a, b, r, patch, dp, slope)
        out.backward(torch.from_numpy(gm).cuda())
        torch.cuda.synchronize()
        return out.detach(), a.grad, b.grad, r.grad

    fo, fa, fb, fr = run(correlate_merge)
    uo, ua, ub, ur = run(_chain_torch)
    assert fo.shape == uo.shape == (B, cr + patch * patch, H, W)
    pow2 = C & (C - 1) == 0
    if pow2:
        assert torch.equal(fo, uo)
        assert torch.equal(fa, ua) and torch.equal(fb, ub)
    else:
        assert _rel(fo.cpu(), uo.cpu()) <= 2e-7
        assert _rel(fa.cpu(), ua.cpu()) <= 1e-6 and _rel(fb.cpu(), ub.cpu()) <= 1e-6
    assert torch.equal(fr, ur)                       # the redir gradient is the slice of grad_merged
    assert torch.equal(fo[:, :cr].cpu(), torch.from_numpy(red))

    # against the CPU oracle (correlation.cpp restatement) with the chain in numpy
    ref = sampler_oracle.forward(in1, in2, 1, patch, 1, 0, 1, dp).reshape(B, patch * patch, H, W) / np.float32(C)
    act = np.where(ref > 0, ref, ref * np.float32(slope))
    assert _rel(fo[:, cr:].cpu().numpy(), act) <= 1e-5
    # mask from the oracle's own sign (elements within rounding of zero may flip: exclude |ref| < 1e-6)
    safe = np.abs(ref) > 1e-6
    gcorr = np.where(ref > 0, gm[:, cr:], gm[:, cr:] * np.float32(slope)) / np.float32(C)
    gcorr = np.where(safe, gcorr, 0).astype(np.float32)
    gm_safe = gm.copy()
    gm_safe[:, cr:] = np.where(safe, gm[:, cr:], 0)
    a = torch.from_numpy(in1).cuda().requires_grad_()
    b = torch.from_numpy(in2).cuda().requires_grad_()
    out = correlate_merge(a, b, torch.from_numpy(red).cuda(), patch, dp, slope)
    out.backward(torch.from_numpy(gm_safe).cuda())
    r1, r2 = sampler_oracle.backward(in1, in2, gcorr.reshape(B, patch, patch, H, W), 1, patch, 1, 0, 1, dp)
    assert _rel(a.grad.cpu().numpy(), r1) <= 1e-5
    assert _rel(b.grad.cpu().numpy(), r2) <= 1e-5


def test_merge_full_size_properties():
    """BASELINE config 2 feature shape: bit-identical to the unfused chain; untouched redir slice."""
    from understanding_flow_robustness_b200 import correlate_merge
    torch.manual_seed(0)
    a = torch.randn(4, 256, 48, 160, device="cuda", requires_grad=True)
    b = torch.randn(4, 256, 48, 160, device="cuda", requires_grad=True)
    r = torch.randn(4, 32, 48, 160, device="cuda")
    g = torch.randn(4, 473, 48, 160, device="cuda")
    fo = correlate_merge(a, b, r)
    fo.backward(g)
    fa, fb = a.grad.clone(), b.grad.clone()
    a.grad = b.grad = None
    uo = _chain_torch(a, b, r, 21, 2, 0.1)
    uo.backward(g)
    assert torch.equal(fo, uo) and torch.equal(fa, a.grad) and torch.equal(fb, b.grad)


def test_merge_pwc_call_site_layout():
    """PWCNet.py:286-292: x = cat((leakyRELU(corr(c15, warp5)), c15, up_flow6, up_feat6), 1), C = 196 (channel tail)."""
    from understanding_flow_robustness_b200 import correlate_merge, spatial_correlation_sample
    torch.manual_seed(3)
    c1 = torch.randn(2, 196, 6, 20, device="cuda", requires_grad=True)
    w2 = torch.randn(2, 196, 6, 20, device="cuda", requires_grad=True)
    flow = torch.randn(2, 2, 6, 20, device="cuda", requires_grad=True)
    feat = torch.randn(2, 2, 6, 20, device="cuda", requires_grad=True)
    g = torch.randn(2, 81 + 196 + 4, 6, 20, device="cuda")
    x = correlate_merge(c1, w2, None, 9, 1, 0.1, after=(c1, flow, feat))
    grads = torch.autograd.grad(x, (c1, w2, flow, feat), g)
    corr = spatial_correlation_sample(c1, w2, kernel_size=1, patch_size=9, stride=1)
    ref = torch.cat((F.leaky_relu(corr.view(2, 81, 6, 20) / 196, 0.1), c1, flow, feat), 1)
    rgrads = torch.autograd.grad(ref, (c1, w2, flow, feat), g)
    assert x.shape == ref.shape and _rel(x.detach().cpu(), ref.detach().cpu()) <= 2e-7
    for a, b in zip(grads, rgrads):
        assert _rel(a.cpu(), b.cpu()) <= 1e-6


def test_merge_rejects_what_it_does_not_cover():
    from understanding_flow_robustness_b200 import correlate_merge
    a = torch.randn(1, 16, 8, 16, device="cuda")
    r = torch.randn(1, 4, 8, 16, device="cuda")
    with pytest.raises(RuntimeError):
        correlate_merge(a, a, r, patch_size=5, dilation_patch=1)       # no register-blocked instantiation
    with pytest.raises(RuntimeError):
        correlate_merge(a.cpu(), a.cpu(), r.cpu())                      # no CPU path
    with pytest.raises(RuntimeError):
        correlate_merge(a, a, r, negative_slope=-0.5)                   # mask is recovered from the output's sign
    with pytest.raises(RuntimeError):
        correlate_merge(a, a, r[:, :, :4])                              # redir of another spatial size


def test_flownetc_harness_fused_equals_unfused():
    from understanding_flow_robustness_b200.harness.flownetc import FlowNetCHarness
    torch.manual_seed(1)
    net = FlowNetCHarness().cuda().eval()
    x1 = torch.rand(1, 3, 128, 256, device="cuda", requires_grad=True)
    x2 = torch.rand(1, 3, 128, 256, device="cuda")
    net.fused_merge = False
    y0 = net(x1, x2)
    (g0,) = torch.autograd.grad(y0.square().mean(), x1)
    net.fused_merge = True
    y1 = net(x1, x2)
    (g1,) = torch.autograd.grad(y1.square().mean(), x1)
    # the merge block itself is bit-identical (tests above); around it cuDNN's conv kernels are not run-to-run
    # exact (algorithm choice, split-K atomics), so the network-level check carries a tolerance
    assert float((y0 - y1).abs().max()) <= 1e-5 * float(y0.abs().max())
    assert float((g0 - g1).abs().max()) <= 1e-3 * float(g0.abs().max())
