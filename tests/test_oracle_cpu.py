"""Pin the CPU oracle (oracle/) against vectors produced by the reference itself.

tests/golden/sampler_*.npz come from the reference's compiled CPU extension driven through its own
Python wrapper; tests/golden/raft_*.npz from the reference's models/raft/corr.py (see
oracle/make_golden.py).  When the compiled reference (oracle/_ref) is present it is additionally
exercised live on fresh seeds.
"""
import glob
import os

import numpy as np
import pytest

from oracle import raft_oracle, sampler_oracle

from conftest import GOLDEN

SAMPLER = sorted(glob.glob(os.path.join(GOLDEN, "sampler_*.npz")))
RAFT = sorted(glob.glob(os.path.join(GOLDEN, "raft_*.npz")))


def _params(z):
    p = z["params"]
    return dict(kernel_size=tuple(p[0]), patch_size=tuple(p[1]), stride=tuple(p[2]),
                padding=tuple(p[3]), dilation=tuple(p[4]), dilation_patch=tuple(p[5]))


def test_golden_present():
    assert len(SAMPLER) >= 9 and len(RAFT) >= 3


@pytest.mark.parametrize("path", SAMPLER, ids=[os.path.basename(p)[:-4] for p in SAMPLER])
def test_sampler_oracle_matches_reference_golden(path):
    z = np.load(path)
    kw = _params(z)
    out = sampler_oracle.forward(z["in1"], z["in2"], **kw)
    g1, g2 = sampler_oracle.backward(z["in1"], z["in2"], z["gout"], **kw)
    # same loops, same accumulation order as correlation.cpp -> bit-exact
    np.testing.assert_array_equal(out, z["out"])
    np.testing.assert_array_equal(g1, z["gin1"])
    np.testing.assert_array_equal(g2, z["gin2"])


@pytest.mark.parametrize("path", RAFT, ids=[os.path.basename(p)[:-4] for p in RAFT])
def test_raft_oracle_matches_reference_golden(path):
    z = np.load(path)
    L, r = int(z["levels"]), int(z["radius"])
    pyr = raft_oracle.build_pyramid(z["f1"], z["f2"], L)
    for l in range(L):
        ref = z[f"pyr{l}"]
        scale = np.abs(ref).max()
        # fp32 SGEMM accumulation order is a BLAS detail: <= 1e-5 of the volume's range
        assert np.abs(pyr[l] - ref).max() <= 1e-5 * scale, l
    # lookup from the REFERENCE pyramid isolates the sampler arithmetic; golden was made on CPU
    ref_pyr = [z[f"pyr{l}"] for l in range(L)]
    out = raft_oracle.lookup(ref_pyr, z["coords"], r, unnorm="cpu")
    scale = np.abs(z["out"]).max()
    assert np.abs(out - z["out"]).max() <= 2e-6 * scale
    # the CUDA un-normalisation differs from the CPU one by an ulp of the coordinate
    out_cuda = raft_oracle.lookup(ref_pyr, z["coords"], r, unnorm="cuda")
    assert np.abs(out_cuda - z["out"]).max() <= 1e-4 * scale
    # and the round-trip-free formula (used by alt_cuda_corr) stays within the survey's bound
    out_exact = raft_oracle.lookup(ref_pyr, z["coords"], r, roundtrip=False)
    assert np.abs(out_exact - z["out"]).max() <= 1e-4 * scale


@pytest.mark.parametrize("path", RAFT[:2], ids=[os.path.basename(p)[:-4] for p in RAFT[:2]])
def test_alt_oracle_equals_corrblock_by_linearity(path):
    """pooling the feature map == pooling the volume (avg-pool is linear)."""
    z = np.load(path)
    L, r = int(z["levels"]), int(z["radius"])
    alt = raft_oracle.alternate_corr_block(z["f1"], z["f2"], z["coords"], L, r)
    scale = np.abs(z["out"]).max()
    assert alt.shape == z["out"].shape
    assert np.abs(alt - z["out"]).max() <= 1e-4 * scale


def test_alt_oracle_backward_is_adjoint():
    """<J^T g, (df1, df2)> == <g, J (df1, df2)> for the bilinear-in-features forward."""
    rng = np.random.default_rng(0)
    B, H, W, C, r = 1, 5, 6, 3, 2
    f1 = rng.standard_normal((B, H, W, C)).astype(np.float32)
    f2 = rng.standard_normal((B, H, W, C)).astype(np.float32)
    coords = (np.stack(np.meshgrid(np.arange(W), np.arange(H)), -1)[None, None]
              + 1.5 * rng.standard_normal((B, 1, H, W, 2))).astype(np.float32)
    g = rng.standard_normal((B, 1, (2 * r + 1) ** 2, H, W)).astype(np.float32)
    g1, g2, gc = raft_oracle.alt_corr_backward(f1, f2, coords, g, r)
    assert not gc.any()
    d1 = rng.standard_normal(f1.shape).astype(np.float32)
    d2 = rng.standard_normal(f2.shape).astype(np.float32)
    # forward is bilinear: directional derivative = F(d1, f2) + F(f1, d2)
    jvp = (raft_oracle.alt_corr_forward(d1, f2, coords, r).astype(np.float64)
           + raft_oracle.alt_corr_forward(f1, d2, coords, r))
    lhs = float((g * jvp).sum())
    rhs = float((g1.astype(np.float64) * d1).sum() + (g2.astype(np.float64) * d2).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))


def test_sampler_oracle_live_against_compiled_reference():
    from oracle import build_ref
    if not os.path.exists(build_ref.so_path()):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    import torch
    backend = build_ref.load_backend()
    rng = np.random.default_rng(7)
    for (B, C, H, W, k, p, s, pad, dil, dp) in [(2, 6, 13, 17, 1, 21, 1, 0, 1, 2),
                                                (1, 4, 9, 12, 3, 5, 2, 2, 1, 3),
                                                (2, 3, 8, 8, 1, 9, 1, 0, 1, 1)]:
        a = rng.standard_normal((B, C, H, W)).astype(np.float32)
        b = rng.standard_normal((B, C, H, W)).astype(np.float32)
        ref = backend.forward(torch.from_numpy(a), torch.from_numpy(b), k, k, p, p, pad, pad,
                              dil, dil, dp, dp, s, s).numpy()
        out = sampler_oracle.forward(a, b, k, p, s, pad, dil, dp)
        np.testing.assert_array_equal(out, ref)
        g = rng.standard_normal(ref.shape).astype(np.float32)
        r1, r2 = backend.backward(torch.from_numpy(a), torch.from_numpy(b), torch.from_numpy(g),
                                  k, k, p, p, pad, pad, dil, dil, dp, dp, s, s)
        o1, o2 = sampler_oracle.backward(a, b, g, k, p, s, pad, dil, dp)
        np.testing.assert_array_equal(o1, r1.numpy())
        np.testing.assert_array_equal(o2, r2.numpy())


WARP = sorted(glob.glob(os.path.join(GOLDEN, "warp_*.npz")))


@pytest.mark.parametrize("path", WARP, ids=[os.path.basename(p)[:-4] for p in WARP])
def test_warp_oracle_matches_reference_golden(path):
    """oracle/warp_oracle.py vs vectors from the reference's own PWCDCNet.warp (oracle/make_golden_warp.py)."""
    import torch

    from oracle import warp_oracle
    z = np.load(path)
    x = torch.from_numpy(z["x"]).requires_grad_()
    flo = torch.from_numpy(z["flo"]).requires_grad_()
    out = warp_oracle.warp(x, flo)
    gx, gf = torch.autograd.grad(out, (x, flo), torch.from_numpy(z["gout"]))
    np.testing.assert_array_equal(out.detach().numpy(), z["out"])
    np.testing.assert_allclose(gx.numpy(), z["gx"], rtol=0, atol=1e-6 * np.abs(z["gx"]).max())
    np.testing.assert_allclose(gf.numpy(), z["gflo"], rtol=0, atol=1e-6 * np.abs(z["gflo"]).max())


def test_flownet2_oracle_against_torch_formulas():
    """oracle/flownet2_oracle.py (restated from channelnorm_kernel.cu / resample2d_kernel.cu) against independent torch
    formulas: the channel 2-norm and its gradient; resample2d == grid_sample(bilinear, padding_mode="border",
    align_corners=True) at pixel + flow (clamping the four corners is border replication), forward and image gradient
    (away from integer positions, where the reference's `xf - int(xf)` and floor agree for xf >= 0)."""
    import torch

    from oracle import flownet2_oracle as O
    rng = np.random.default_rng(7)
    x = rng.standard_normal((2, 3, 9, 11)).astype(np.float32)
    g1 = rng.standard_normal((2, 1, 9, 11)).astype(np.float32)
    t = torch.from_numpy(x).requires_grad_()
    n = t.pow(2).sum(1, keepdim=True).sqrt()
    n.backward(torch.from_numpy(g1))
    out = O.channelnorm_forward(x)
    np.testing.assert_allclose(out, n.detach().numpy(), rtol=1e-6)
    np.testing.assert_allclose(O.channelnorm_backward(x, out, g1), t.grad.numpy(), rtol=1e-5, atol=1e-7)

    B, C, H, W = 2, 3, 9, 11
    flow = (2.5 * rng.standard_normal((B, 2, H, W))).astype(np.float32)
    flow[:, 0] = np.abs(flow[:, 0])           # xf, yf >= 0: truncation == floor in the backward weights
    flow[:, 1] = np.abs(flow[:, 1])
    g = rng.standard_normal((B, C, H, W)).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    gx = 2 * (xs[None] + flow[:, 0]) / (W - 1) - 1
    gy = 2 * (ys[None] + flow[:, 1]) / (H - 1) - 1
    grid = torch.from_numpy(np.stack([gx, gy], -1).astype(np.float32))
    t = torch.from_numpy(x).requires_grad_()
    ref = torch.nn.functional.grid_sample(t, grid, mode="bilinear", padding_mode="border", align_corners=True)
    ref.backward(torch.from_numpy(g))
    np.testing.assert_allclose(O.resample2d_forward(x, flow), ref.detach().numpy(), rtol=0, atol=2e-5)
    np.testing.assert_allclose(O.resample2d_backward(x, flow, g)[0], t.grad.numpy(), rtol=0, atol=2e-5)
