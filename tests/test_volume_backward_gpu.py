"""Backward of the all-pairs volume + pyramid (csrc/raft_volume_bwd.cu) against what autograd derives for the
reference's matmul + 3 x avg_pool2d (models/raft/corr.py:55-64,98-106).

  precision "fp32": exact CUDA-core kernels            <= 2e-5 relative (fp32 summation order)
  precision "tf32": tcgen05, operands truncated to TF32 <= 2^-9 * sum |g f| per entry (two truncations of
                    relative 2^-10 each), K-major (dF1) and MN-major (dF2) operand paths"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_library_math():
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32 = prev


def _reference(glv, f1, f2, scale):
    """fold the per-level gradients to level 0 (avg_pool2d backward), then the two products in fp64."""
    B, C, H, W = f1.shape
    g = glv[-1].double()
    for l in range(len(glv) - 2, -1, -1):
        h, w = glv[l].shape[-2:]
        up = F.interpolate(g, scale_factor=2, mode="nearest") / 4
        full = torch.zeros_like(glv[l], dtype=torch.float64)
        full[:, :, :up.shape[-2], :up.shape[-1]] = up
        g = glv[l].double() + full
    G = g.view(B, H * W, H * W)
    d1 = scale * torch.bmm(f2.double().view(B, C, -1), G.transpose(1, 2)).view_as(f1)
    d2 = scale * torch.bmm(f1.double().view(B, C, -1), G).view_as(f2)
    a1 = scale * torch.bmm(f2.abs().double().view(B, C, -1), G.abs().transpose(1, 2)).view_as(f1)
    a2 = scale * torch.bmm(f1.abs().double().view(B, C, -1), G.abs()).view_as(f2)
    return d1, d2, a1, a2


@pytest.mark.parametrize("B,C,H,W,L", [
    (1, 16, 8, 16, 3), (2, 32, 12, 16, 3), (1, 64, 16, 32, 4), (2, 256, 16, 24, 4),
    (1, 40, 13, 20, 3),       # odd height: floor-mode pooling drops the last row; C not a multiple of 16
    (1, 24, 17, 30, 4),       # level 1 = 8x15 = 120 keys fine, level 2 = 4x7 = 28, level 3 = 2x3 = 6: not 16-byte rows -> exact kernels
    (1, 320, 8, 16, 2),       # C > 256 -> exact kernels
])
@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_volume_backward_equals_fold_plus_matmul(B, C, H, W, L, precision):
    from understanding_flow_robustness_b200 import _lib, raft_corr
    g = torch.Generator(device="cuda").manual_seed(B * 100 + C + H)
    f1 = torch.randn(B, C, H, W, device="cuda", generator=g)
    f2 = torch.randn(B, C, H, W, device="cuda", generator=g)
    glv = [torch.randn(B * H * W, 1, H >> l, W >> l, device="cuda", generator=g) for l in range(L)]
    # make the gradient pyramid sparse like the lookups leave it (most entries zero)
    glv = [v * (torch.rand(v.shape, device="cuda", generator=g) < 0.3) for v in glv]
    scale = 1.0 / math.sqrt(C)
    n0 = _lib.lib().b200corr_launch_count()
    g1, g2 = raft_corr.volume_backward(glv, f1, f2, scale, precision)
    assert _lib.lib().b200corr_launch_count() > n0
    d1, d2, a1, a2 = _reference(glv, f1, f2, scale)
    for got, want, mag in ((g1, d1, a1), (g2, d2, a2)):
        err = (got.double() - want).abs()
        if precision == "fp32":
            assert float(err.max()) <= 2e-5 * float(want.abs().max())
        else:
            assert bool((err <= 2.0 ** -9 * mag + 1e-5 * float(want.abs().max())).all()), float((err / (mag + 1e-9)).max())


def test_volume_backward_full_size_tensor_core_paths():
    """BASELINE config 3 (B=4, 256x48x160, 4 levels): K-split items, atomics, MN-major operand, all levels."""
    from understanding_flow_robustness_b200 import raft_corr
    B, C, H, W, L = 4, 256, 48, 160, 4
    g = torch.Generator(device="cuda").manual_seed(7)
    f1 = torch.randn(B, C, H, W, device="cuda", generator=g)
    f2 = torch.randn(B, C, H, W, device="cuda", generator=g)
    glv = [torch.zeros(B * H * W, 1, H >> l, W >> l, device="cuda") for l in range(L)]
    # sparse structured gradient: a few windows per query
    for l, v in enumerate(glv):
        idx = torch.randint(0, v[0].numel(), (v.shape[0], 8), device="cuda", generator=g)
        v.view(v.shape[0], -1).scatter_(1, idx, torch.randn(v.shape[0], 8, device="cuda", generator=g))
    scale = 1.0 / 16.0
    g1, g2 = raft_corr.volume_backward(glv, f1, f2, scale, "tf32")
    # reference by linearity in fp32 with torch ops: dF1 = sum_l G_l pool_l(F2)^T
    f2p = [f2]
    for _ in range(L - 1):
        f2p.append(F.avg_pool2d(f2p[-1], 2, stride=2))
    want1 = sum(torch.bmm(f2p[l].view(B, C, -1), glv[l].view(B, H * W, -1).transpose(1, 2)) for l in range(L)).view_as(f1) * scale
    assert float((g1 - want1).abs().max()) <= 4e-3 * float(want1.abs().max())
    d2l = [torch.bmm(f1.view(B, C, -1), glv[l].view(B, H * W, -1)).view(B, C, H >> l, W >> l) * scale for l in range(L)]
    want2 = d2l[0].clone()
    for l in range(1, L):
        up = F.interpolate(d2l[l], scale_factor=2 ** l, mode="nearest") / 4 ** l
        want2[:, :, :up.shape[-2], :up.shape[-1]] += up
    assert float((g2 - want2).abs().max()) <= 4e-3 * float(want2.abs().max())
