"""FlowNet2's native ops against the numpy restatement of the reference kernels
(oracle/flownet2_oracle.py) and, for resample2d, against the reference's own extension compiled unmodified for sm_100a
(oracle/_ref/ref_resample2d_cuda; channelnorm does not compile against torch 2.11).  Tolerances: forward 1e-6 of
max|ref| (resample2d forward vs the compiled reference: bit-exact), gradients 1e-5 (atomics / summation order)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("shape", [(2, 3, 16, 24), (1, 2, 7, 9), (3, 8, 5, 4)], ids=str)
def test_channelnorm(shape):
    from oracle import flownet2_oracle as O
    from understanding_flow_robustness_b200 import ChannelNorm
    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal(shape).astype(np.float32)
    g = rng.standard_normal((shape[0], 1, *shape[2:])).astype(np.float32)
    t = torch.from_numpy(x).cuda().requires_grad_()
    out = ChannelNorm()(t)
    out.backward(torch.from_numpy(g).cuda())
    ref = O.channelnorm_forward(x)
    assert out.shape == (shape[0], 1, *shape[2:])
    assert _rel(out.detach().cpu().numpy(), ref) <= 1e-6
    assert _rel(t.grad.cpu().numpy(), O.channelnorm_backward(x, ref, g)) <= 1e-5
    # the same formula with torch ops
    assert _rel(out.detach().cpu().numpy(), torch.from_numpy(x).pow(2).sum(1, keepdim=True).sqrt().numpy()) <= 1e-6


@pytest.mark.parametrize("case", [(2, 3, 16, 24, 2.0, True), (1, 2, 9, 7, 8.0, True), (2, 4, 8, 8, 1.0, False),
                                  (1, 3, 32, 48, 0.3, True)], ids=str)
def test_resample2d_vs_oracle(case):
    from oracle import flownet2_oracle as O
    from understanding_flow_robustness_b200 import Resample2d
    B, C, H, W, sigma, bilinear = case
    rng = np.random.default_rng(B + C + H + W)
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    f = (sigma * rng.standard_normal((B, 2, H, W))).astype(np.float32)
    g = rng.standard_normal((B, C, H, W)).astype(np.float32)
    tx = torch.from_numpy(x).cuda().requires_grad_()
    tf = torch.from_numpy(f).cuda().requires_grad_()
    out = Resample2d(bilinear=bilinear)(tx, tf)
    out.backward(torch.from_numpy(g).cuda())
    assert _rel(out.detach().cpu().numpy(), O.resample2d_forward(x, f, bilinear)) <= 1e-6
    g1, g2 = O.resample2d_backward(x, f, g)
    assert _rel(tx.grad.cpu().numpy(), g1) <= 1e-5
    assert _rel(tf.grad.cpu().numpy(), g2) <= 1e-5


def test_resample2d_vs_compiled_reference_extension():
    from oracle import build_ref_cuda
    from understanding_flow_robustness_b200.flownet2_natives import resample2d_cuda
    if not os.path.exists(build_ref_cuda.so_path("ref_resample2d_cuda")):
        pytest.skip("oracle/_ref/ref_resample2d_cuda not built")
    ref = build_ref_cuda.load_module("ref_resample2d_cuda")
    torch.manual_seed(5)
    for (B, C, H, W, sigma) in [(2, 3, 24, 40, 3.0), (1, 2, 17, 9, 10.0)]:
        x = torch.randn(B, C, H, W, device="cuda")
        f = sigma * torch.randn(B, 2, H, W, device="cuda")
        g = torch.randn(B, C, H, W, device="cuda")
        o1, o2 = torch.zeros_like(x), torch.zeros_like(x)
        resample2d_cuda.forward(x, f, o1, 1, True)
        ref.forward(x, f, o2, 1, True)
        assert torch.equal(o1, o2)
        a1, a2 = torch.zeros_like(x), torch.zeros_like(f)
        b1, b2 = torch.zeros_like(x), torch.zeros_like(f)
        resample2d_cuda.backward(x, f, g, a1, a2, 1, True)
        ref.backward(x, f, g, b1, b2, 1, True)
        assert float((a1 - b1).abs().max()) <= 1e-5 * float(b1.abs().max())
        assert float((a2 - b2).abs().max()) <= 1e-5 * float(b2.abs().max())


def test_reference_wrapper_modules_resolve_to_this_implementation():
    import sys

    from understanding_flow_robustness_b200 import flownet2_natives, install_reference_shims
    install_reference_shims(raft_package=None)
    assert sys.modules["channelnorm_cuda"] is flownet2_natives.channelnorm_cuda
    assert sys.modules["resample2d_cuda"] is flownet2_natives.resample2d_cuda
    x = torch.randn(1, 3, 8, 8, device="cuda")
    with pytest.raises(RuntimeError):
        flownet2_natives.resample2d_cuda.forward(x, torch.zeros(1, 2, 8, 8, device="cuda"), torch.empty_like(x), 3, True)
    with pytest.raises(RuntimeError):
        flownet2_natives.channelnorm_cuda.forward(x.cpu(), torch.empty(1, 1, 8, 8), 2)
