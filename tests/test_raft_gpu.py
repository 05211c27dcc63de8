"""Parity of the RAFT correlation kernels (all-pairs volume + pyramid, lookup, alt_cuda_corr) through
the C-ABI against the numpy oracle, the reference-generated golden vectors and -- for the pieces
whose reference arithmetic lives in torch (matmul / avg_pool2d / grid_sample) -- torch itself on the
same device.  Tolerances:
  fp32 volume / pyramid  : <= 1e-5 * max|vol|
  tf32 volume            : |d| <= 2^-10 * scale * sum_c|f1 f2|  (+ fp32 accumulation slack)
  lookup vs grid_sample  : <= 1e-4 * max|vol| (coordinate round trip); vs the oracle's restatement
                           of the CUDA arithmetic: <= 2e-6 * max|vol|
  alt vs CorrBlock       : <= 1e-4 * max|out|
"""
import glob
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

RAFT = sorted(glob.glob(os.path.join(GOLDEN, "raft_*.npz")))


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _maxabs(x):
    return float(np.abs(x).max())


def _torch_corrblock(f1, f2, coords, L, r):
    """models/raft/corr.py restated with torch ops (the reference's third-party arithmetic)."""
    B, C, H, W = f1.shape
    corr = torch.matmul(f1.view(B, C, H * W).transpose(1, 2), f2.view(B, C, H * W))
    corr = corr.view(B, H, W, 1, H, W) / torch.sqrt(torch.tensor(C).float())
    corr = corr.reshape(B * H * W, 1, H, W)
    pyr = [corr]
    for _ in range(L - 1):
        corr = F.avg_pool2d(corr, 2, stride=2)
        pyr.append(corr)
    c = coords.permute(0, 2, 3, 1)
    outs = []
    for i in range(L):
        dx = torch.linspace(-r, r, 2 * r + 1, device=f1.device)
        dy = torch.linspace(-r, r, 2 * r + 1, device=f1.device)
        delta = torch.stack(torch.meshgrid(dy, dx, indexing="ij"), axis=-1)
        cl = c.reshape(B * H * W, 1, 1, 2) / 2 ** i + delta.view(1, 2 * r + 1, 2 * r + 1, 2)
        Hl, Wl = pyr[i].shape[-2:]
        xg, yg = cl.split([1, 1], dim=-1)
        grid = torch.cat([2 * xg / (Wl - 1) - 1, 2 * yg / (Hl - 1) - 1], dim=-1)
        outs.append(F.grid_sample(pyr[i], grid, align_corners=True).view(B, H, W, -1))
    return torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous().float(), pyr


@pytest.mark.parametrize("path", RAFT, ids=[os.path.basename(p)[:-4] for p in RAFT])
@pytest.mark.parametrize("precision", ["fp32", "tf32", "tf32x3"])
def test_pyramid_vs_reference_golden(path, precision):
    from understanding_flow_robustness_b200 import raft_corr
    z = np.load(path)
    L = int(z["levels"])
    f1, f2 = _cuda(z["f1"]), _cuda(z["f2"])
    pyr = raft_corr.allpairs_pyramid(f1, f2, L, precision)
    C = z["f1"].shape[1]
    # sum_c |f1 f2| / sqrt(C): the scale of the TF32 rounding bound
    absdot = np.einsum("bcm,bcn->bmn", np.abs(z["f1"]).reshape(*z["f1"].shape[:2], -1),
                       np.abs(z["f2"]).reshape(*z["f2"].shape[:2], -1)).max() / math.sqrt(C)
    for l in range(L):
        ref = z[f"pyr{l}"]
        got = pyr[l].cpu().numpy()
        assert got.shape == ref.shape
        tol = 1e-5 * _maxabs(z["pyr0"]) if precision != "tf32" else 2.0 ** -10 * absdot + 1e-5 * _maxabs(z["pyr0"])
        assert _maxabs(got - ref) <= tol, (l, _maxabs(got - ref), tol)


@pytest.mark.parametrize("path", RAFT, ids=[os.path.basename(p)[:-4] for p in RAFT])
def test_lookup_vs_reference_golden(path):
    """Lookup from the REFERENCE pyramid isolates the sampling arithmetic."""
    from oracle import raft_oracle
    from understanding_flow_robustness_b200 import raft_corr
    z = np.load(path)
    L, r = int(z["levels"]), int(z["radius"])
    B, _, H, W = z["coords"].shape
    pyr = [_cuda(z[f"pyr{l}"]) for l in range(L)]
    coords = _cuda(z["coords"])
    scale = _maxabs(z["out"])
    out = raft_corr.lookup_forward(pyr, coords, r, H, W, "grid_sample").cpu().numpy()
    assert out.shape == z["out"].shape
    assert _maxabs(out - z["out"]) <= 1e-4 * scale                      # golden made by torch on CPU
    ora = raft_oracle.lookup([z[f"pyr{l}"] for l in range(L)], z["coords"], r, unnorm="cuda")
    assert _maxabs(out - ora) <= 2e-6 * scale                           # same arithmetic restated
    # grid_sample on this device: the reference's actual code path on a GPU
    tref, _ = None, None
    pt = [p for p in pyr]
    c = coords.permute(0, 2, 3, 1)
    outs = []
    for i in range(L):
        d = torch.linspace(-r, r, 2 * r + 1, device="cuda")
        delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), axis=-1)
        cl = c.reshape(B * H * W, 1, 1, 2) / 2 ** i + delta.view(1, 2 * r + 1, 2 * r + 1, 2)
        Hl, Wl = pt[i].shape[-2:]
        xg, yg = cl.split([1, 1], dim=-1)
        grid = torch.cat([2 * xg / (Wl - 1) - 1, 2 * yg / (Hl - 1) - 1], dim=-1)
        outs.append(F.grid_sample(pt[i], grid, align_corners=True).view(B, H, W, -1))
    tref = torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous().cpu().numpy()
    assert _maxabs(out - tref) <= 5e-6 * scale   # ATen may contract the un-normalise into an FMA
    # direct mode: within the documented round-trip bound of the reference
    outd = raft_corr.lookup_forward(pyr, coords, r, H, W, "direct").cpu().numpy()
    assert _maxabs(outd - z["out"]) <= 1e-4 * scale


LOOKUP_CASES = [  # (B, H, W, levels, radius): level widths hit the 256-bit, 128-bit and scalar staging
    (2, 16, 64, 4, 4), (1, 12, 40, 3, 3), (2, 9, 13, 2, 2), (1, 8, 24, 3, 1), (1, 20, 36, 2, 4), (1, 33, 32, 4, 4)]


@pytest.mark.parametrize("case", LOOKUP_CASES, ids=[str(c) for c in LOOKUP_CASES])
def test_lookup_shapes_vs_grid_sample(case):
    """Every staging flavour of the lookup (row tails, misaligned level bases, integer / far /
    non-finite coordinates) against F.grid_sample on the same device and the numpy oracle."""
    from oracle import raft_oracle
    from understanding_flow_robustness_b200 import coords_grid, raft_corr
    B, H, W, L, r = case
    torch.manual_seed(H * 100 + W)
    pyr = [torch.randn(B * H * W, 1, H // 2 ** l, W // 2 ** l, device="cuda") for l in range(L)]
    # one level lives at a 4-byte offset inside a bigger buffer: exercises the alignment dispatch
    buf = torch.randn(pyr[-1].numel() + 1, device="cuda")
    pyr[-1] = buf[1:].view_as(pyr[-1])
    grid = coords_grid(B, H, W, "cuda")
    for name, coords in [
        ("integer", grid.clone()),                                   # round-trip noise around integers
        ("near", grid + 2.5 * torch.randn(B, 2, H, W, device="cuda")),
        ("far", grid + 60.0 * torch.randn(B, 2, H, W, device="cuda")),
    ]:
        out = raft_corr.lookup_forward(pyr, coords, r, H, W, "grid_sample")
        c = coords.permute(0, 2, 3, 1)
        outs = []
        for i in range(L):
            d = torch.linspace(-r, r, 2 * r + 1, device="cuda")
            delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), axis=-1)
            cl = c.reshape(B * H * W, 1, 1, 2) / 2 ** i + delta.view(1, 2 * r + 1, 2 * r + 1, 2)
            Hl, Wl = pyr[i].shape[-2:]
            xg, yg = cl.split([1, 1], dim=-1)
            g = torch.cat([2 * xg / (Wl - 1) - 1, 2 * yg / (Hl - 1) - 1], dim=-1)
            outs.append(F.grid_sample(pyr[i], g, align_corners=True).view(B, H, W, -1))
        tref = torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous()
        scale = float(tref.abs().max()) + 1e-6
        # a level of extent 1 divides by zero in the reference's normalisation (NaN/inf there too)
        ok = torch.isfinite(tref)
        assert float((out - tref)[ok].abs().max()) <= 1e-5 * scale, name
        assert bool((torch.isfinite(out) == ok).all()), name
        ora = raft_oracle.lookup([p_.cpu().numpy() for p_ in pyr], coords.cpu().numpy(), r, unnorm="cuda")
        okn = np.isfinite(ora)
        assert _maxabs((out.cpu().numpy() - ora)[okn]) <= 2e-6 * scale, name
        # backward: gradient w.r.t. every pyramid level vs autograd through grid_sample; two calls
        # accumulate into the same gradient pyramid (what 12 RAFT iterations do)
        if bool(ok.all()):
            gout = torch.randn_like(out)
            leaves = [p_.detach().clone().requires_grad_() for p_ in pyr]
            outs = []
            for i in range(L):
                d = torch.linspace(-r, r, 2 * r + 1, device="cuda")
                delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), axis=-1)
                cl = c.reshape(B * H * W, 1, 1, 2) / 2 ** i + delta.view(1, 2 * r + 1, 2 * r + 1, 2)
                Hl, Wl = leaves[i].shape[-2:]
                xg, yg = cl.split([1, 1], dim=-1)
                g = torch.cat([2 * xg / (Wl - 1) - 1, 2 * yg / (Hl - 1) - 1], dim=-1)
                outs.append(F.grid_sample(leaves[i], g, align_corners=True).view(B, H, W, -1))
            (torch.cat(outs, dim=-1).permute(0, 3, 1, 2) * gout).sum().backward()
            glv = [torch.zeros_like(p_) for p_ in pyr]
            glv[-1] = torch.zeros(pyr[-1].numel() + 1, device="cuda")[1:].view_as(pyr[-1])   # misaligned level
            raft_corr.lookup_backward(glv, coords, gout, r, H, W, "grid_sample")
            raft_corr.lookup_backward(glv, coords, gout, r, H, W, "grid_sample")
            for i in range(L):
                want = 2 * leaves[i].grad
                assert float((glv[i] - want).abs().max()) <= 1e-5 * float(want.abs().max()) + 1e-6, (name, i)
    # non-finite coordinates must not fault; finite queries are unaffected
    bad = grid + 2.5 * torch.randn(B, 2, H, W, device="cuda")
    good = raft_corr.lookup_forward(pyr, bad, r, H, W, "direct")
    bad2 = bad.clone()
    bad2[0, 0, 0, 0] = float("nan")
    bad2[0, 1, H - 1, W - 1] = float("inf")
    bad2[0, 0, H // 2, W // 2] = 3e30
    got = raft_corr.lookup_forward(pyr, bad2, r, H, W, "direct")
    mask = torch.ones(B, 1, H, W, dtype=torch.bool, device="cuda")
    mask[0, 0, 0, 0] = mask[0, 0, H - 1, W - 1] = mask[0, 0, H // 2, W // 2] = False
    assert bool((got == good)[mask.expand_as(got)].all())
    assert float(got[0, :, H // 2, W // 2].abs().max()) == 0.0


TC_SHAPES = [(1, 32, 8, 32), (2, 64, 11, 20), (1, 40, 16, 36), (1, 256, 24, 64), (3, 8, 5, 8)]


@pytest.mark.parametrize("shape", TC_SHAPES, ids=[str(s) for s in TC_SHAPES])
def test_tcgen05_volume_ragged_shapes(shape):
    """tcgen05 path (W % 4 == 0) with M / patch tails, channel padding, and 1..5 levels."""
    from understanding_flow_robustness_b200 import raft_corr
    B, C, H, W = shape
    torch.manual_seed(B * 1000 + C + H + W)
    f1 = torch.randn(B, C, H, W, device="cuda")
    f2 = torch.randn(B, C, H, W, device="cuda")
    L = 1
    while min(H, W) // 2 ** L >= 1 and L < 5:
        L += 1
    pyr = raft_corr.allpairs_pyramid(f1, f2, L, "tf32")
    a = f1.double().view(B, C, -1)
    b = f2.double().view(B, C, -1)
    exact = torch.einsum("bcm,bcn->bmn", a, b) / math.sqrt(C)
    bound = 2.0 ** -10 * torch.einsum("bcm,bcn->bmn", a.abs(), b.abs()) / math.sqrt(C) + 1e-5
    v0 = pyr[0].view(B, H * W, H * W).double()
    assert bool(((v0 - exact).abs() <= bound).all()), float(((v0 - exact).abs() - bound).max())
    ref = v0.float().view(B * H * W, 1, H, W)
    for l in range(1, L):
        ref = F.avg_pool2d(ref, 2, stride=2)
        assert pyr[l].shape == ref.shape
        assert float((pyr[l] - ref).abs().max()) <= 2e-6 * float(v0.abs().max()) + 1e-7, l
    # exact kernel agrees with the fp64 contraction to fp32 accuracy, and so does split TF32 (lo*hi + hi*lo +
    # hi*hi on the tensor cores: the dropped lo*lo term is 2^-22 relative per product)
    e32 = raft_corr.allpairs_pyramid(f1, f2, 1, "fp32")[0].view(B, H * W, H * W).double()
    assert float((e32 - exact).abs().max()) <= 1e-5 * float(exact.abs().max())
    x3 = raft_corr.allpairs_pyramid(f1, f2, L, "tf32x3")
    v3 = x3[0].view(B, H * W, H * W).double()
    bound3 = 2.0 ** -20 * torch.einsum("bcm,bcn->bmn", a.abs(), b.abs()) / math.sqrt(C) + 2e-6 * float(exact.abs().max())
    assert bool(((v3 - exact).abs() <= bound3).all()), float(((v3 - exact).abs() - bound3).max())
    assert [tuple(t.shape) for t in x3] == [tuple(t.shape) for t in pyr]


def test_corrblock_end_to_end_and_autograd():
    """CorrBlock (fp32 volume) forward + gradients w.r.t. both feature maps vs torch autograd through
    matmul / avg_pool2d / grid_sample; two lookups accumulate into one gradient pyramid."""
    from understanding_flow_robustness_b200 import CorrBlock, coords_grid
    torch.manual_seed(3)
    B, C, H, W, L, r = 2, 16, 12, 16, 3, 3
    f1 = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
    f2 = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
    ca = coords_grid(B, H, W, "cuda") + 2.5 * torch.randn(B, 2, H, W, device="cuda")
    cb = coords_grid(B, H, W, "cuda") + 9.0 * torch.randn(B, 2, H, W, device="cuda")
    blk = CorrBlock(f1, f2, num_levels=L, radius=r, precision="fp32")
    oa, ob = blk(ca), blk(cb)
    ga, gb = torch.randn_like(oa), torch.randn_like(ob)
    (oa * ga).sum().add((ob * gb).sum()).backward()
    g1, g2 = f1.grad.clone(), f2.grad.clone()
    f1.grad = f2.grad = None
    ra, _ = _torch_corrblock(f1, f2, ca, L, r)
    rb, _ = _torch_corrblock(f1, f2, cb, L, r)
    (ra * ga).sum().add((rb * gb).sum()).backward()
    s = float(ra.abs().max())
    assert float((oa - ra).abs().max()) <= 1e-5 * s and float((ob - rb).abs().max()) <= 1e-5 * s
    assert float((g1 - f1.grad).abs().max()) <= 2e-5 * float(f1.grad.abs().max())
    assert float((g2 - f2.grad).abs().max()) <= 2e-5 * float(f2.grad.abs().max())
    # no-grad construction gives the same values and the public pyramid layout (corr.py:66-67)
    with torch.no_grad():
        blk2 = CorrBlock(f1, f2, num_levels=L, radius=r, precision="fp32")
        assert [tuple(v.shape) for v in blk2.get_corr_pyramid()] == [(B * H * W, 1, H // 2 ** l, W // 2 ** l) for l in range(L)]
        assert float((blk2(ca) - oa).abs().max()) == 0.0


def test_corrblock_tf32_gradients():
    """precision="tf32": volume and both gradient GEMMs on the tensor cores; against the exact path
    the gradients stay within the TF32 input-rounding bound (relative 2^-10 per product)."""
    from understanding_flow_robustness_b200 import CorrBlock, coords_grid
    torch.manual_seed(11)
    B, C, H, W, L, r = 1, 64, 16, 32, 3, 3
    f1 = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
    f2 = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
    c = coords_grid(B, H, W, "cuda") + 2.0 * torch.randn(B, 2, H, W, device="cuda")
    grads = {}
    for prec in ("fp32", "tf32"):
        f1.grad = f2.grad = None
        blk = CorrBlock(f1, f2, num_levels=L, radius=r, precision=prec)
        out = blk(c)
        torch.manual_seed(12)
        (out * torch.randn_like(out)).sum().backward()
        grads[prec] = (f1.grad.clone(), f2.grad.clone())
    for a, b in zip(grads["tf32"], grads["fp32"]):
        assert float((a - b).abs().max()) <= 4e-3 * float(b.abs().max())
        assert float((a - b).abs().max()) > 0.0   # the TF32 path really ran


@pytest.mark.parametrize("path", RAFT[:2], ids=[os.path.basename(p)[:-4] for p in RAFT[:2]])
def test_alt_corr_vs_oracle_and_corrblock(path):
    from oracle import raft_oracle
    from understanding_flow_robustness_b200 import AlternateCorrBlock, alt_cuda_corr
    z = np.load(path)
    L, r = int(z["levels"]), int(z["radius"])
    f1n = np.ascontiguousarray(z["f1"].transpose(0, 2, 3, 1))
    f2n = np.ascontiguousarray(z["f2"].transpose(0, 2, 3, 1))
    B, H, W, C = f1n.shape
    coords = np.ascontiguousarray(z["coords"].transpose(0, 2, 3, 1).reshape(B, 1, H, W, 2))
    (corr,) = alt_cuda_corr.forward(_cuda(f1n), _cuda(f2n), _cuda(coords), r)
    ref = raft_oracle.alt_corr_forward(f1n, f2n, coords, r)
    assert corr.shape == ref.shape
    assert _maxabs(corr.cpu().numpy() - ref) <= 1e-5 * _maxabs(ref)
    g = np.random.default_rng(1).standard_normal(ref.shape).astype(np.float32)
    g1, g2, gc = alt_cuda_corr.backward(_cuda(f1n), _cuda(f2n), _cuda(coords), _cuda(g), r)
    r1, r2, rc = raft_oracle.alt_corr_backward(f1n, f2n, coords, g, r)
    assert _maxabs(g1.cpu().numpy() - r1) <= 1e-5 * _maxabs(r1)
    assert _maxabs(g2.cpu().numpy() - r2) <= 1e-5 * _maxabs(r2)
    assert not gc.any()
    # AlternateCorrBlock == CorrBlock output (golden) by linearity of the average pool
    alt = AlternateCorrBlock(_cuda(z["f1"]), _cuda(z["f2"]), num_levels=L, radius=r)(_cuda(z["coords"]))
    assert alt.shape == z["out"].shape
    assert _maxabs(alt.cpu().numpy() - z["out"]) <= 1e-4 * _maxabs(z["out"])


@pytest.mark.parametrize("shape,dense_from", [((2, 64, 24, 48), 1), ((1, 32, 16, 32), 1), ((1, 16, 8, 16), 0),
                                              ((1, 32, 16, 20), 3)], ids=str)
def test_alt_corr_hybrid_equals_alt_kernels(shape, dense_from):
    """AlternateCorrBlock serves the coarse levels (small key maps) from dense split-TF32 volumes + the
    lookup kernel; the result must equal the pure alt_cuda_corr structure (dense_max_keys=0) and CorrBlock."""
    from understanding_flow_robustness_b200 import AlternateCorrBlock, CorrBlock, coords_grid
    B, C, H, W = shape
    torch.manual_seed(H + W)
    f1 = torch.randn(B, C, H, W, device="cuda")
    f2 = torch.randn(B, C, H, W, device="cuda")
    for sigma in (2.0, 30.0):
        c = coords_grid(B, H, W, "cuda") + sigma * torch.randn(B, 2, H, W, device="cuda")
        with torch.no_grad():
            pure = AlternateCorrBlock(f1, f2, 3, 3, dense_max_keys=0)
            hyb = AlternateCorrBlock(f1, f2, 3, 3, dense_max_keys=400)
            assert pure._dense_from == 3 and hyb._dense_from == dense_from   # (16x20: level widths 10, 5 -> not applicable)
            a, b = pure(c), hyb(c)
            ref = CorrBlock(f1, f2, 3, 3, precision="fp32", lookup_mode="direct")(c)
        s = float(a.abs().max())
        assert a.shape == b.shape == ref.shape
        assert float((a - b).abs().max()) <= 1e-5 * s
        assert float((b - ref).abs().max()) <= 1e-4 * s


def test_alt_corr_multi_n_and_autograd():
    from oracle import raft_oracle
    from understanding_flow_robustness_b200 import AlternateCorrBlock, alt_cuda_corr, coords_grid
    rng = np.random.default_rng(11)
    B, N, H, W, C, r = 2, 3, 6, 9, 36, 2
    f1 = rng.standard_normal((B, H, W, C)).astype(np.float32)
    f2 = rng.standard_normal((B, 4, 5, C)).astype(np.float32)       # fmap2 at another resolution
    coords = (rng.uniform(-3, 8, (B, N, H, W, 2))).astype(np.float32)
    (corr,) = alt_cuda_corr.forward(_cuda(f1), _cuda(f2), _cuda(coords), r)
    ref = raft_oracle.alt_corr_forward(f1, f2, coords, r)
    assert _maxabs(corr.cpu().numpy() - ref) <= 1e-5 * _maxabs(ref)
    g = rng.standard_normal(ref.shape).astype(np.float32)
    g1, g2, _ = alt_cuda_corr.backward(_cuda(f1), _cuda(f2), _cuda(coords), _cuda(g), r)
    r1, r2, _ = raft_oracle.alt_corr_backward(f1, f2, coords, g, r)
    assert _maxabs(g1.cpu().numpy() - r1) <= 1e-5 * _maxabs(r1)
    assert _maxabs(g2.cpu().numpy() - r2) <= 1e-5 * _maxabs(r2)
    # the block is differentiable (the reference's is not): gradient == CorrBlock's (fp32, direct)
    from understanding_flow_robustness_b200 import CorrBlock
    torch.manual_seed(0)
    a = torch.randn(1, 8, 8, 12, device="cuda", requires_grad=True)
    b = torch.randn(1, 8, 8, 12, device="cuda", requires_grad=True)
    c = coords_grid(1, 8, 12, "cuda") + 1.7 * torch.randn(1, 2, 8, 12, device="cuda")
    oa = AlternateCorrBlock(a, b, num_levels=2, radius=2)(c)
    w = torch.randn_like(oa)
    (oa * w).sum().backward()
    ga, gb = a.grad.clone(), b.grad.clone()
    a.grad = b.grad = None
    oc = CorrBlock(a, b, num_levels=2, radius=2, precision="fp32", lookup_mode="direct")(c)
    (oc * w).sum().backward()
    assert float((oa - oc).abs().max()) <= 1e-5 * float(oc.abs().max())
    assert float((ga - a.grad).abs().max()) <= 1e-4 * float(a.grad.abs().max())
    assert float((gb - b.grad).abs().max()) <= 1e-4 * float(b.grad.abs().max())


def test_alt_corr_vs_compiled_reference_extension():
    """Live pin against the reference's own alt_cuda_corr compiled for sm_100a (oracle/_ref)."""
    from oracle import build_ref_cuda
    if not os.path.exists(build_ref_cuda.so_path("ref_alt_cuda_corr")):
        pytest.skip("oracle/_ref/ref_alt_cuda_corr not built")
    try:
        ref_mod = build_ref_cuda.load_module("ref_alt_cuda_corr")
    except Exception as e:  # ABI mismatch on the box: the numpy oracle still pins the kernel
        pytest.skip(f"reference extension not loadable: {e}")
    from understanding_flow_robustness_b200 import alt_cuda_corr
    torch.manual_seed(5)
    B, H, W, C, r = 2, 16, 24, 64, 4
    f1 = torch.randn(B, H, W, C, device="cuda")
    f2 = torch.randn(B, H, W, C, device="cuda")
    coords = (torch.rand(B, 1, H, W, 2, device="cuda") * torch.tensor([W + 6.0, H + 6.0], device="cuda") - 3.0)
    (ours,) = alt_cuda_corr.forward(f1, f2, coords, r)
    (theirs,) = ref_mod.forward(f1, f2, coords, r)
    torch.cuda.synchronize()
    assert float((ours - theirs).abs().max()) <= 1e-5 * float(theirs.abs().max())
    g = torch.randn_like(ours)
    o1, o2, _ = alt_cuda_corr.backward(f1, f2, coords, g, r)
    t1, t2, _ = ref_mod.backward(f1, f2, coords, g, r)
    torch.cuda.synchronize()
    assert float((o1 - t1).abs().max()) <= 2e-5 * float(t1.abs().max())
    assert float((o2 - t2).abs().max()) <= 2e-5 * float(t2.abs().max())


def test_full_size_raft_properties():
    """BASELINE config 3: B=4, 256x48x160, 4 levels, radius 4 -- size-independent properties."""
    from understanding_flow_robustness_b200 import AlternateCorrBlock, CorrBlock, coords_grid
    torch.manual_seed(0)
    B, C, H, W = 4, 256, 48, 160
    f1 = torch.randn(B, C, H, W, device="cuda")
    f2 = torch.randn(B, C, H, W, device="cuda")
    blk = CorrBlock(f1, f2, num_levels=4, radius=4, precision="tf32")
    pyr = blk.get_corr_pyramid()
    assert [tuple(v.shape) for v in pyr] == [(B * H * W, 1, 48, 160), (B * H * W, 1, 24, 80),
                                             (B * H * W, 1, 12, 40), (B * H * W, 1, 6, 20)]
    # random entries of level 0 against fp64 dot products, within the TF32 bound
    g = torch.Generator(device="cuda").manual_seed(1)
    bi = torch.randint(0, B, (4096,), device="cuda", generator=g)
    m = torch.randint(0, H * W, (4096,), device="cuda", generator=g)
    n = torch.randint(0, H * W, (4096,), device="cuda", generator=g)
    a = f1.view(B, C, -1)[bi, :, m].double()
    b = f2.view(B, C, -1)[bi, :, n].double()
    exact = (a * b).sum(1) / 16.0
    bound = 2.0 ** -10 * (a.abs() * b.abs()).sum(1) / 16.0 + 1e-5
    got = pyr[0].view(B, H * W, H * W)[bi, m, n].double()
    assert bool(((got - exact).abs() <= bound).all())
    # every pooled level is the 2x2 mean of the level below
    for l in range(1, 4):
        ref = F.avg_pool2d(pyr[l - 1][: 2 * H * W], 2, stride=2)
        assert float((pyr[l][: 2 * H * W] - ref).abs().max()) <= 1e-5
    # lookup at integer coordinates (direct mode) returns volume entries exactly
    c = coords_grid(B, H, W, "cuda")
    out = CorrBlock(f1, f2, 4, 4, precision="tf32", lookup_mode="direct")(c)
    assert out.shape == (B, 324, H, W)
    centre = out[:, 4 * 9 + 4]                      # level 0, zero offset
    diag = pyr[0].view(B, H * W, H * W).diagonal(dim1=1, dim2=2).reshape(B, H, W)
    assert float((centre - diag).abs().max()) == 0.0
    # alternate path == volume path (same TF32 question aside: compare against the fp32 volume)
    c2 = c + 3.0 * torch.randn(B, 2, H, W, device="cuda", generator=g)
    ref = CorrBlock(f1[:1], f2[:1], 4, 4, precision="fp32", lookup_mode="direct")(c2[:1])
    alt = AlternateCorrBlock(f1[:1], f2[:1], 4, 4)(c2[:1])
    assert float((alt - ref).abs().max()) <= 1e-4 * float(ref.abs().max())


@pytest.mark.parametrize("shape,levels,precision,expect_mask", [
    ((1, 32, 16, 32), 4, "tf32", 3), ((2, 64, 32, 48), 3, "tf32", 3), ((1, 16, 48, 160), 4, "tf32x3", 3),
    ((1, 24, 16, 16), 1, "tf32", 1), ((2, 8, 16, 80), 2, "tf32", 3), ((1, 40, 64, 96), 4, "tf32", 3),
    ((1, 32, 24, 32), 4, "tf32", 3), ((1, 32, 16, 40), 3, "tf32", 3), ((1, 32, 16, 32), 4, "fp32", 0),
    # padded tiles: H not a multiple of 8, W / 2 not a multiple of 8 (FlyingThings 68x120, Sintel 55x128 shapes)
    ((1, 16, 68, 120), 4, "tf32", 3), ((1, 16, 55, 128), 4, "tf32", 3), ((2, 8, 13, 24), 3, "tf32", 3),
    ((1, 8, 9, 8), 2, "tf32", 3), ((1, 8, 20, 36), 4, "tf32", 0),
    # more levels than the fused epilogue produces (levels >= 4 come from the pooling kernel), many queries per tile tail
    ((1, 16, 32, 32), 5, "tf32", 3), ((3, 8, 24, 40), 4, "tf32x3", 3)], ids=str)
def test_blocked_volume_layout_equals_rowmajor(shape, levels, precision, expect_mask):
    """The blocked layout (8x8 tiles of 64 floats, include/b200corr.h) is another element order of the same
    values: de-blocked levels and every lookup are bit-identical to the row-major kernels."""
    from understanding_flow_robustness_b200 import CorrBlock, coords_grid
    B, C, H, W = shape
    torch.manual_seed(H * W + C)
    f1 = torch.randn(B, C, H, W, device="cuda")
    f2 = torch.randn(B, C, H, W, device="cuda")
    with torch.no_grad():
        a = CorrBlock(f1, f2, levels, 4, precision=precision, layout="auto")
        b = CorrBlock(f1, f2, levels, 4, precision=precision, layout="rowmajor")
        assert a._blocked == expect_mask and b._blocked == 0
        for la, lb in zip(a.get_corr_pyramid(), b.get_corr_pyramid()):
            assert la.shape == lb.shape and torch.equal(la, lb)
        for sigma in (0.0, 3.0, 50.0):
            c = coords_grid(B, H, W, "cuda") + sigma * torch.randn(B, 2, H, W, device="cuda")
            assert torch.equal(a(c), b(c))
        for r in (1, 2, 3):
            a.radius = b.radius = r
            c = coords_grid(B, H, W, "cuda") + 2.0 * torch.randn(B, 2, H, W, device="cuda")
            assert torch.equal(a(c), b(c))


@pytest.mark.parametrize("shape,levels,precision", [
    ((1, 32, 16, 32), 4, "tf32"), ((2, 64, 32, 48), 3, "tf32"), ((1, 16, 48, 160), 4, "tf32x3"), ((1, 24, 16, 16), 1, "tf32"),
    ((1, 16, 68, 120), 4, "tf32"), ((1, 16, 55, 128), 4, "tf32"), ((2, 8, 13, 24), 3, "tf32"), ((3, 8, 24, 40), 4, "tf32x3"),
    ((1, 16, 32, 32), 5, "tf32")], ids=str)
def test_fp16_volume_storage(shape, levels, precision):
    """storage="fp16" (include/b200corr.h, b200corr_allpairs_pyramid_storage): the two blocked levels hold the fp32
    values rounded ONCE to fp16 (round to nearest: bit-identical to `.half()` of the fp32 kernel's levels), the
    coarse levels stay fp32, and every lookup is bit-identical to the fp32 lookup of those rounded values -- hence
    within 2^-11 * max|level| of the fp32-stored block."""
    from understanding_flow_robustness_b200 import CorrBlock, coords_grid
    B, C, H, W = shape
    torch.manual_seed(H * W + C + 1)
    f1 = torch.randn(B, C, H, W, device="cuda")
    f2 = torch.randn(B, C, H, W, device="cuda")
    with torch.no_grad():
        a = CorrBlock(f1, f2, levels, 4, precision=precision, layout="auto", storage="fp16")
        ref = CorrBlock(f1, f2, levels, 4, precision=precision, layout="rowmajor")
        mask = a._blocked
        assert mask == (1 if levels == 1 else 3)
        assert [v.dtype for v in a._levels] == [torch.float16 if (mask >> l) & 1 else torch.float32 for l in range(levels)]
        rounded = CorrBlock(f1, f2, levels, 4, precision=precision, layout="rowmajor")
        rounded._levels = [v.half().float() if (mask >> l) & 1 else v for l, v in enumerate(ref.get_corr_pyramid())]
        for l, (la, lr) in enumerate(zip(a.get_corr_pyramid(), rounded._levels)):
            assert la.dtype == torch.float32 and la.shape == lr.shape and torch.equal(la, lr), l
        vmax = max(float(v.abs().max()) for v in ref.get_corr_pyramid())
        for sigma in (0.0, 3.0, 50.0):
            c = coords_grid(B, H, W, "cuda") + sigma * torch.randn(B, 2, H, W, device="cuda")
            out = a(c)
            assert torch.equal(out, rounded(c))
            assert float((out - ref(c)).abs().max()) <= 2.0 ** -11 * vmax * 1.01   # + fp32 rounding of the interpolation itself
        for r in (1, 2, 3):
            a.radius = rounded.radius = r
            c = coords_grid(B, H, W, "cuda") + 2.0 * torch.randn(B, 2, H, W, device="cuda")
            assert torch.equal(a(c), rounded(c))


def test_fp16_volume_storage_saturates_and_keeps_gradients():
    """Finite values beyond the fp16 range saturate at +-65504 (cvt.rn.satfinite), never inf; the backward does not
    read the forward volume, so gradients are those of the fp32-stored block; lookup_convc1 still works."""
    from understanding_flow_robustness_b200 import CorrBlock, coords_grid
    torch.manual_seed(3)
    B, C, H, W = 1, 32, 16, 32
    f1 = torch.randn(B, C, H, W, device="cuda")
    f2 = torch.randn(B, C, H, W, device="cuda")
    with torch.no_grad():
        big = CorrBlock(3000.0 * f1, 3000.0 * f2, 4, 4, precision="tf32", storage="fp16")
        l0 = big.get_corr_pyramid()[0]
        assert torch.isfinite(l0).all() and float(l0.abs().max()) == 65504.0
    c = coords_grid(B, H, W, "cuda") + 2.0 * torch.randn(B, 2, H, W, device="cuda")
    g = torch.randn(B, 4 * 81, H, W, device="cuda")
    grads = []
    for storage in ("fp16", "fp32"):
        a1, a2 = f1.clone().requires_grad_(), f2.clone().requires_grad_()
        blk = CorrBlock(a1, a2, 4, 4, precision="tf32", backward_precision="fp32", storage=storage)
        out = blk(c)
        grads.append(torch.autograd.grad(out, (a1, a2), g))
    for x, y in zip(*grads):
        assert torch.equal(x, y)
    with torch.no_grad():
        w = torch.randn(64, 4 * 81, 1, 1, device="cuda")
        bias = torch.randn(64, device="cuda")
        blk = CorrBlock(f1, f2, 4, 4, precision="tf32", storage="fp16")
        want = torch.relu(torch.nn.functional.conv2d(blk(c), w, bias))
        for impl in ("pipelined", "fused"):
            got = blk.lookup_convc1(c, w, bias, impl=impl)
            assert float((got - want).abs().max()) <= 2e-2 * float(want.abs().max())
    with pytest.raises(ValueError):
        CorrBlock(f1, f2, 4, 4, layout="rowmajor", storage="fp16")


def test_blocked_layout_with_autograd():
    """The backward never reads the forward volume: gradients are those of the row-major block."""
    from understanding_flow_robustness_b200 import CorrBlock, coords_grid
    torch.manual_seed(11)
    B, C, H, W = 1, 32, 16, 32
    f1 = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
    f2 = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
    c = coords_grid(B, H, W, "cuda") + 2.0 * torch.randn(B, 2, H, W, device="cuda")
    g = torch.randn(B, 4 * 81, H, W, device="cuda")
    res = []
    for layout in ("auto", "rowmajor"):
        # backward_precision="fp32": the exact kernels are deterministic (the tensor-core products land with atomics)
        blk = CorrBlock(f1, f2, 4, 4, layout=layout, backward_precision="fp32")
        out = blk(c)
        res.append((out.detach(), *torch.autograd.grad(out, (f1, f2), g)))
        assert blk._blocked == (3 if layout == "auto" else 0)
    for x, y in zip(*res):
        assert torch.equal(x, y)


def test_corrblock_memory_is_released_by_refcount():
    """The reference's pyramid dies with `corr_fn` (raft.py:150-156 keeps the only reference).  Ours must too:
    no reference cycle block -> handle -> grad_fn -> ctx -> block, and an alive lookup graph must not pin the
    1.25 GB forward volume (its backward never reads it)."""
    import gc

    from understanding_flow_robustness_b200 import CorrBlock, coords_grid
    gc.collect()
    gc.disable()
    try:
        torch.manual_seed(5)
        B, C, H, W = 1, 64, 32, 64
        f1 = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
        f2 = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
        c = coords_grid(B, H, W, "cuda") + torch.randn(B, 2, H, W, device="cuda")
        torch.cuda.synchronize()
        base = torch.cuda.memory_allocated()
        blk = CorrBlock(f1, f2, 4, 4)
        vol_bytes = sum(v.numel() * 4 for v in blk._levels)
        assert torch.cuda.memory_allocated() - base >= vol_bytes
        out = blk(c)
        del blk                                              # graph (out) still alive: the volume must go anyway
        assert torch.cuda.memory_allocated() - base < vol_bytes // 2
        loss = out.square().mean()
        loss.backward()
        assert f1.grad is not None and float(f1.grad.abs().max()) > 0
        del out, loss
        f1.grad = f2.grad = None
        torch.cuda.synchronize()
        assert torch.cuda.memory_allocated() == base         # no gc.collect() happened
    finally:
        gc.enable()
