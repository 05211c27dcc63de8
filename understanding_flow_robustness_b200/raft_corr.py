"""RAFT correlation blocks on the B200 kernels: `CorrBlock`, `AlternateCorrBlock`, `alt_cuda_corr`.

Mirrors models/raft/corr.py of the reference (same constructor / call signatures, same output
layout and channel order) and the pybind surface of models/alt_cuda_corr/correlation.cpp:51-54.

  CorrBlock(fmap1, fmap2, num_levels=4, radius=4, compute_spatial=False)(coords) -> (B, L*(2r+1)^2, H, W)
      reference: corr.py:26-106.  The all-pairs volume, its 1/sqrt(C) scale and the pooled pyramid
      come out of one tcgen05 kernel; every lookup is one gather kernel for all levels.
  AlternateCorrBlock(fmap1, fmap2, num_levels=4, radius=4)(coords)
      reference: corr.py:109-137.  Same result without materialising the volume.
  alt_cuda_corr.forward(fmap1, fmap2, coords, radius) -> [corr]
  alt_cuda_corr.backward(fmap1, fmap2, coords, corr_grad, radius) -> [fmap1_grad, fmap2_grad, coords_grad]

Differences from the reference, all deliberate:
  * the volume is contracted on the tensor cores: `precision="tf32x3"` (default) is the split-TF32
    contraction with the accuracy of the reference's fp32 matmul, `precision="tf32"` one TF32 pass (half
    the build time; bound in DESIGN.md), `precision="fp32"` the exact CUDA-core kernel; the environment
    variable B200CORR_VOLUME_PRECISION picks the default for callers that cannot pass arguments;
  * both blocks are differentiable w.r.t. the feature maps (the reference's AlternateCorrBlock is
    forward-only because nothing wraps alt_cuda_corr.backward); coordinates get no gradient, as in
    the reference (raft.py:188 detaches them, correlation_kernel.cu:307 returns zeros);
  * lookups accumulate their gradient in place into ONE dense gradient pyramid per block instead of
    allocating a dense volume gradient per grid_sample call.
"""
import ctypes
import math
import os
import types
import weakref

import torch
import torch.nn.functional as F

from . import _lib
from .spatial_correlation_sampler import spatial_correlation_sample

PRECISIONS = {"tf32": 0, "tf32x3": 1, "fp32": 2}
LOOKUP_MODES = {"grid_sample": 0, "direct": 1}


def coords_grid(batch, ht, wd, device=None):
    """models/raft/utils/utils.py:79-82 -- (B, 2, H, W), channel 0 = x."""
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)


def bilinear_sampler(img, coords, mask=False):
    """models/raft/utils/utils.py:62-76 (kept for callers that import it from the corr module)."""
    H, W = img.shape[-2:]
    xgrid, ygrid = coords.split([1, 1], dim=-1)
    xgrid = 2 * xgrid / (W - 1) - 1
    ygrid = 2 * ygrid / (H - 1) - 1
    grid = torch.cat([xgrid, ygrid], dim=-1)
    img = F.grid_sample(img, grid, align_corners=True)
    if mask:
        m = (xgrid > -1) & (ygrid > -1) & (xgrid < 1) & (ygrid < 1)
        return img, m.float()
    return img


def _require_cuda_f32(who, *ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError(f"{who}: CUDA tensors required (this build has no CPU path)")
        if t.dtype != torch.float32:
            raise RuntimeError(f"{who}: float32 tensors required, got {t.dtype}")
        if t.device != ts[0].device:
            raise RuntimeError(f"{who}: tensors must be on the same device")


def _level_shapes(H, W, num_levels):
    shapes = []
    for _ in range(num_levels):
        shapes.append((H, W))
        H, W = H // 2, W // 2
    return shapes


# ------------------------------------------------------------------------------------------------
# raw C-ABI calls
def allpairs_pyramid(fmap1, fmap2, num_levels=4, precision="tf32", blocked=False, storage="fp32"):
    """List of (B*H*W, 1, H_l, W_l) tensors: vol_0 = f1^T f2 / sqrt(C), vol_{l+1} = avg_pool2d(vol_l, 2, 2).

    blocked=True returns `(levels, mask)`: the levels whose bit is set in `mask` hold each slice as a grid of 8x8
    tiles of 64 consecutive floats (include/b200corr.h, "blocked volume layout") -- same tensor shapes, same
    values, another element order; `deblock()` gives the reference's row-major view of such a level.  mask is 0
    where the library has no blocked kernel for the problem.
    storage="fp16" (with blocked=True): the blocked levels are fp16 tensors (include/b200corr.h,
    b200corr_allpairs_pyramid_storage) -- half the bytes to write and to look up; returns `(levels, mask)` where the
    levels in `mask` have dtype float16.  Falls back to fp32 storage where there is no blocked kernel (mask 0)."""
    if storage not in ("fp32", "fp16"):
        raise ValueError("allpairs_pyramid: storage must be 'fp32' or 'fp16'")
    if storage == "fp16" and not blocked:
        raise ValueError("allpairs_pyramid: fp16 storage exists for the blocked layout only (blocked=True)")
    fmap1 = fmap1.contiguous()
    fmap2 = fmap2.contiguous()
    _require_cuda_f32("allpairs_pyramid", fmap1, fmap2)
    if fmap1.shape != fmap2.shape or fmap1.dim() != 4:
        raise RuntimeError("allpairs_pyramid: fmap1 and fmap2 must both be (B, C, H, W) of the same shape")
    B, C, H, W = fmap1.shape
    prec = PRECISIONS[precision]
    L = _lib.lib()
    with torch.cuda.device(fmap1.device):
        mask = L.b200corr_allpairs_blocked_levels(num_levels, H, W, prec) if blocked and B > 0 else 0
    shapes = _level_shapes(H, W, num_levels)
    for l in range(num_levels):
        if (mask >> l) & 1:
            shapes[l] = blocked_level_dims(l, H, W)   # padded to whole 8x8 tiles
    half_mask = mask if storage == "fp16" else 0
    levels = [torch.empty((B * H * W, 1, h, w), dtype=torch.float16 if (half_mask >> l) & 1 else torch.float32,
                          device=fmap1.device) for l, (h, w) in enumerate(shapes)]
    nbytes = L.b200corr_allpairs_workspace_bytes(B, C, H, W, prec)
    ws = torch.empty((max(nbytes, 1) + 127) // 128 * 32, dtype=torch.float32, device=fmap1.device)
    with torch.cuda.device(fmap1.device):
        code = L.b200corr_allpairs_pyramid_storage(_lib.ptr(fmap1), _lib.ptr(fmap2), _lib.ptr_array(levels), num_levels,
                                                   B, C, H, W, H, W, 1.0 / math.sqrt(C), prec, mask, half_mask,
                                                   _lib.ptr(ws), nbytes, _lib.current_stream(fmap1.device))
    _lib.check(code, "b200corr_allpairs_pyramid_storage")
    return (levels, mask) if blocked else levels


def blocked_level_dims(level, H, W):
    """(Hp, Wp) of a blocked level: the level's extent padded to whole 8x8 tiles (include/b200corr.h)."""
    hp, wp = ctypes.c_int(), ctypes.c_int()
    _lib.lib().b200corr_blocked_level_dims(level, H, W, ctypes.byref(hp), ctypes.byref(wp))
    return hp.value, wp.value


def deblock(level, h, w):
    """Row-major (Q, 1, h, w) copy of a level stored in the blocked layout (padded to (Hp, Wp) = level.shape[2:])."""
    Q, _, hp, wp = level.shape
    full = level.view(Q, hp // 8, wp // 8, 8, 8).permute(0, 1, 3, 2, 4).reshape(Q, 1, hp, wp)
    return full[:, :, :h, :w].float().contiguous()    # fp16-stored levels come back as fp32 (exact)


def _half_mask(levels, blocked_levels):
    """Bit i: levels[i] is an fp16 tensor (legal for blocked levels only)."""
    m = 0
    for i, v in enumerate(levels):
        if v.dtype == torch.float16:
            if not (blocked_levels >> i) & 1:
                raise RuntimeError("lookup: an fp16 level must be in the blocked layout")
            m |= 1 << i
    return m


def lookup_forward(levels, coords, radius, H, W, mode="grid_sample", first_level=0, blocked_levels=0):
    """`levels[i]` is pyramid level first_level + i: extent (H, W) >> (first_level + i), sampled at
    coords / 2^(first_level + i).  Bit i of `blocked_levels`: levels[i] is in the blocked layout; a blocked level
    may be an fp16 tensor (allpairs_pyramid(..., storage="fp16"))."""
    coords = coords.contiguous()
    half = _half_mask(levels, blocked_levels)
    _require_cuda_f32("lookup_forward", coords, *[v for i, v in enumerate(levels) if not (half >> i) & 1])
    for v in levels:
        if not v.is_cuda or v.device != coords.device:
            raise RuntimeError("lookup_forward: levels and coords must be CUDA tensors on the same device")
    B = coords.shape[0]
    n = (2 * radius + 1) ** 2
    out = torch.empty((B, len(levels) * n, H, W), dtype=torch.float32, device=coords.device)
    with torch.cuda.device(coords.device):
        code = _lib.lib().b200corr_lookup_forward_storage(_lib.ptr_array(levels), len(levels), first_level,
                                                          blocked_levels, half, _lib.ptr(coords), _lib.ptr(out), B, H, W,
                                                          radius, LOOKUP_MODES[mode], _lib.current_stream(coords.device))
    _lib.check(code, "b200corr_lookup_forward_storage")
    return out


_CONVC1_CACHE = {}


def prepare_convc1(weight, num_levels, radius):
    """Prepared (TF32-rounded, per-level, zero-padded) copy of a convc1 weight `(n_out, L*(2r+1)^2[, 1, 1])` for
    `lookup_convc1_forward`; cached per (storage, version), i.e. recomputed after an optimiser step."""
    w = weight.detach()
    n_out = w.shape[0]
    w2 = w.reshape(n_out, -1).contiguous().float()
    if w2.shape[1] != num_levels * (2 * radius + 1) ** 2:
        raise RuntimeError(f"prepare_convc1: weight has {w2.shape[1]} input channels, expected "
                           f"{num_levels} * {(2 * radius + 1) ** 2}")
    # cached per weight OBJECT (a weak reference guards against a recycled id / data pointer) and version counter
    key = (id(weight), num_levels, radius)
    hit = _CONVC1_CACHE.get(key)
    if hit is not None and hit[0]() is weight and hit[1] == weight._version and hit[2] == weight.data_ptr():
        return hit[3]
    _require_cuda_f32("prepare_convc1", w2)
    L = _lib.lib()
    nbytes = L.b200corr_lookup_convc1_weight_bytes(num_levels, radius, n_out)
    wp = torch.empty(nbytes // 4, dtype=torch.float32, device=w2.device)
    with torch.cuda.device(w2.device):
        code = L.b200corr_lookup_convc1_prepare(_lib.ptr(w2), _lib.ptr(wp), num_levels, radius, n_out,
                                                _lib.current_stream(w2.device))
    _lib.check(code, "b200corr_lookup_convc1_prepare")
    for k in [k for k, v in _CONVC1_CACHE.items() if v[0]() is None]:     # weights that are gone
        del _CONVC1_CACHE[k]
    _CONVC1_CACHE[key] = (weakref.ref(weight), weight._version, weight.data_ptr(), wp)
    return wp


def lookup_convc1_forward(levels, coords, wprep, bias, n_out, radius, H, W, mode="grid_sample", blocked_levels=0,
                          relu=True):
    """act(convc1(lookup(levels, coords))) without materialising the lookup: (B, n_out, H, W).
    models/raft/raft.py:189 + models/raft/update.py:104,111 as one kernel (include/b200corr.h)."""
    coords = coords.contiguous()
    _require_cuda_f32("lookup_convc1_forward", coords, wprep, *levels)
    if bias is not None:
        bias = bias.detach().contiguous().float()
        _require_cuda_f32("lookup_convc1_forward", coords, bias)
    B = coords.shape[0]
    out = torch.empty((B, n_out, H, W), dtype=torch.float32, device=coords.device)
    with torch.cuda.device(coords.device):
        code = _lib.lib().b200corr_lookup_convc1_forward(
            _lib.ptr_array(levels), len(levels), blocked_levels, _lib.ptr(coords), _lib.ptr(wprep),
            _lib.ptr(bias) if bias is not None else None, _lib.ptr(out), B, H, W, radius, LOOKUP_MODES[mode], n_out,
            1 if relu else 0, _lib.current_stream(coords.device))
    _lib.check(code, "b200corr_lookup_convc1_forward")
    return out


def conv1x1_forward(x, weight, bias=None, relu=False):
    """act(conv2d(x, weight (n_out, K, 1, 1), bias)) on the tensor cores (TF32), x read as it lies: (B, K, H, W) ->
    (B, n_out, H, W).  The 1x1 convolution behind the lookup (models/raft/update.py:104,111)."""
    x = x.contiguous()
    n_out = weight.shape[0]
    w2 = weight.detach().reshape(n_out, -1).contiguous().float()
    _require_cuda_f32("conv1x1_forward", x, w2)
    B, K, H, W = x.shape
    if w2.shape[1] != K:
        raise RuntimeError(f"conv1x1_forward: weight has {w2.shape[1]} input channels, x has {K}")
    if bias is not None:
        bias = bias.detach().contiguous().float()
    out = torch.empty((B, n_out, H, W), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        code = _lib.lib().b200corr_conv1x1_forward(_lib.ptr(x), _lib.ptr(w2), _lib.ptr(bias) if bias is not None else None,
                                                   _lib.ptr(out), B, K, n_out, H * W, 1 if relu else 0,
                                                   _lib.current_stream(x.device))
    _lib.check(code, "b200corr_conv1x1_forward")
    return out


def allpairs_volume_rect(fmap1, fmap2, precision="tf32x3"):
    """(B*H1*W1, 1, H2, W2) volume of f1 (queries) against a key map of another size (AlternateCorrBlock's
    pooled f2): vol[b,p1,y,x] = f1[b,:,p1] . f2[b,:,y,x] / sqrt(C).  Tensor-core precisions, W2 % 4 == 0."""
    fmap1 = fmap1.contiguous()
    fmap2 = fmap2.contiguous()
    _require_cuda_f32("allpairs_volume_rect", fmap1, fmap2)
    B, C, H1, W1 = fmap1.shape
    H2, W2 = fmap2.shape[-2:]
    if fmap2.shape[:2] != fmap1.shape[:2] or W2 % 4 != 0 or precision == "fp32":
        raise RuntimeError("allpairs_volume_rect: need equal batch / channels, W2 % 4 == 0 and a tensor-core precision")
    prec = PRECISIONS[precision]
    L = _lib.lib()
    vol = torch.empty((B * H1 * W1, 1, H2, W2), dtype=torch.float32, device=fmap1.device)
    nbytes = L.b200corr_allpairs_rect_workspace_bytes(B, C, H1, W1, H2, W2, prec)
    ws = torch.empty((max(nbytes, 1) + 127) // 128 * 32, dtype=torch.float32, device=fmap1.device)
    with torch.cuda.device(fmap1.device):
        code = L.b200corr_allpairs_pyramid_rect(_lib.ptr(fmap1), _lib.ptr(fmap2), _lib.ptr_array([vol]), 1, B, C, H1, W1,
                                                H2, W2, 1.0 / math.sqrt(C), prec, _lib.ptr(ws), nbytes,
                                                _lib.current_stream(fmap1.device))
    _lib.check(code, "b200corr_allpairs_pyramid_rect")
    return vol


def lookup_backward(grad_levels, coords, grad_out, radius, H, W, mode="grid_sample"):
    coords = coords.contiguous()
    grad_out = grad_out.contiguous()
    _require_cuda_f32("lookup_backward", coords, grad_out, *grad_levels)
    B = coords.shape[0]
    with torch.cuda.device(coords.device):
        code = _lib.lib().b200corr_lookup_backward(_lib.ptr_array(grad_levels), len(grad_levels),
                                                   _lib.ptr(coords), _lib.ptr(grad_out), B, H, W, radius,
                                                   LOOKUP_MODES[mode], _lib.current_stream(coords.device))
    _lib.check(code, "b200corr_lookup_backward")


def pyramid_backward(grad_levels, B, H, W):
    with torch.cuda.device(grad_levels[0].device):
        code = _lib.lib().b200corr_pyramid_backward(_lib.ptr_array(grad_levels), len(grad_levels), B, H, W,
                                                    _lib.current_stream(grad_levels[0].device))
    _lib.check(code, "b200corr_pyramid_backward")


def volume_backward(grad_levels, fmap1, fmap2, scale, precision):
    """dF1 = scale * fold(G) . F2^T, dF2 = scale * fold(G)^T . F1 from the UNFOLDED per-level gradients: what autograd
    derives for `torch.matmul(fmap1^T, fmap2)` + 3 x `avg_pool2d` (models/raft/corr.py:55-64,104).  The fold is
    moved onto the feature maps (average pooling commutes with the contraction), the two products run on the
    tensor cores in TF32 (precision "tf32" / "tf32x3") or exactly on the CUDA cores ("fp32")."""
    fmap1 = fmap1.contiguous()
    fmap2 = fmap2.contiguous()
    _require_cuda_f32("volume_backward", fmap1, fmap2, *grad_levels)
    B, C, H, W = fmap1.shape
    L = _lib.lib()
    g1 = torch.empty_like(fmap1)
    g2 = torch.empty_like(fmap2)
    nbytes = L.b200corr_volume_backward_workspace_bytes(len(grad_levels), B, C, H, W)
    ws = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=fmap1.device)
    with torch.cuda.device(fmap1.device):
        code = L.b200corr_volume_backward(_lib.ptr_array(grad_levels), len(grad_levels), _lib.ptr(fmap1), _lib.ptr(fmap2),
                                          _lib.ptr(g1), _lib.ptr(g2), B, C, H, W, scale, PRECISIONS[precision],
                                          _lib.ptr(ws), nbytes, _lib.current_stream(fmap1.device))
    _lib.check(code, "b200corr_volume_backward")
    return g1, g2


# ------------------------------------------------------------------------------------------------
# alt_cuda_corr drop-in (models/alt_cuda_corr/correlation.cpp:23-54)
def _alt_check(fmap1, fmap2, coords):
    for t in (fmap1, fmap2, coords):
        if not t.is_cuda:
            raise RuntimeError("alt_cuda_corr: must be a CUDA tensor")          # CHECK_CUDA
        if not t.is_contiguous():
            raise RuntimeError("alt_cuda_corr: must be contiguous")             # CHECK_CONTIGUOUS
    _require_cuda_f32("alt_cuda_corr", fmap1, fmap2, coords)
    if fmap1.dim() != 4 or fmap2.dim() != 4 or coords.dim() != 5 or coords.shape[-1] != 2:
        raise RuntimeError("alt_cuda_corr: expected fmap (B,H,W,C) and coords (B,N,H,W,2)")
    if fmap1.shape[0] != fmap2.shape[0] or fmap1.shape[3] != fmap2.shape[3]:
        raise RuntimeError("alt_cuda_corr: batch / channel mismatch between the feature maps")
    if coords.shape[0] != fmap1.shape[0] or tuple(coords.shape[2:4]) != tuple(fmap1.shape[1:3]):
        raise RuntimeError("alt_cuda_corr: coords must be (B, N, H1, W1, 2)")


def _alt_forward(fmap1, fmap2, coords, radius):
    _alt_check(fmap1, fmap2, coords)
    B, H1, W1, C = fmap1.shape
    _, H2, W2, _ = fmap2.shape
    N = coords.shape[1]
    rd = 2 * radius + 1
    corr = torch.empty((B, N, rd * rd, H1, W1), dtype=torch.float32, device=fmap1.device)
    with torch.cuda.device(fmap1.device):
        code = _lib.lib().b200corr_altcorr_forward(_lib.ptr(fmap1), _lib.ptr(fmap2), _lib.ptr(coords),
                                                   _lib.ptr(corr), B, N, H1, W1, H2, W2, C, radius,
                                                   _lib.current_stream(fmap1.device))
    _lib.check(code, "b200corr_altcorr_forward")
    return [corr]


def _alt_backward(fmap1, fmap2, coords, corr_grad, radius):
    _alt_check(fmap1, fmap2, coords)
    corr_grad = corr_grad.contiguous()
    B, H1, W1, C = fmap1.shape
    _, H2, W2, _ = fmap2.shape
    N = coords.shape[1]
    g1 = torch.empty_like(fmap1)
    g2 = torch.empty_like(fmap2)
    gc = torch.empty_like(coords)
    with torch.cuda.device(fmap1.device):
        code = _lib.lib().b200corr_altcorr_backward(_lib.ptr(fmap1), _lib.ptr(fmap2), _lib.ptr(coords),
                                                    _lib.ptr(corr_grad), _lib.ptr(g1), _lib.ptr(g2),
                                                    _lib.ptr(gc), B, N, H1, W1, H2, W2, C, radius,
                                                    _lib.current_stream(fmap1.device))
    _lib.check(code, "b200corr_altcorr_backward")
    return [g1, g2, gc]


alt_cuda_corr = types.ModuleType("alt_cuda_corr")
alt_cuda_corr.__doc__ = "B200 drop-in for the reference's alt_cuda_corr extension (forward / backward)."
alt_cuda_corr.forward = _alt_forward
alt_cuda_corr.backward = _alt_backward


class _AltCorrFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fmap1, fmap2, coords, radius):
        ctx.save_for_backward(fmap1, fmap2, coords)
        ctx.radius = radius
        return _alt_forward(fmap1, fmap2, coords, radius)[0]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad):
        fmap1, fmap2, coords = ctx.saved_tensors
        g1, g2, _ = _alt_backward(fmap1, fmap2, coords, grad, ctx.radius)
        return g1, g2, None, None


# ------------------------------------------------------------------------------------------------
class _GradState:
    """What the backward of a CorrBlock needs, and nothing that references the block, its pyramid or the
    autograd handle: the graph (handle -> grad_fn -> ctx) must not keep the 1.25 GB volume alive, nor form a
    cycle with the block (the reference's pyramid is freed by refcount when `corr_fn` goes out of scope)."""

    __slots__ = ("B", "H", "W", "num_levels", "radius", "lookup_mode", "precision", "backward_precision", "grad_levels")

    def __init__(self, B, H, W, num_levels, radius, lookup_mode, precision, backward_precision):
        self.B, self.H, self.W, self.num_levels = B, H, W, num_levels
        self.radius, self.lookup_mode, self.precision = radius, lookup_mode, precision
        self.backward_precision = backward_precision
        self.grad_levels = None


class _VolumeFunction(torch.autograd.Function):
    """Builds the pyramid into `block` and returns a 1-element handle that carries the autograd edge
    from the lookups back to the feature maps."""

    @staticmethod
    def forward(ctx, fmap1, fmap2, block):
        block._build(fmap1, fmap2)
        ctx.save_for_backward(fmap1, fmap2)
        ctx.state = block._state
        return fmap1.new_zeros(1)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, _grad_handle):
        fmap1, fmap2 = ctx.saved_tensors
        st = ctx.state
        B, C, H, W = fmap1.shape
        gl = st.grad_levels
        st.grad_levels = None
        if gl is None:
            return torch.zeros_like(fmap1), torch.zeros_like(fmap2), None
        g1, g2 = volume_backward(gl, fmap1, fmap2, 1.0 / math.sqrt(C), st.backward_precision)
        return g1, g2, None


class _LookupFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, handle, coords, block):
        ctx.state = block._state
        ctx.save_for_backward(coords)
        return lookup_forward(block._levels, coords, block.radius, block.H, block.W, block.lookup_mode,
                              blocked_levels=block._blocked)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        (coords,) = ctx.saved_tensors
        st = ctx.state
        if st.grad_levels is None:
            # the gradient pyramid is row-major whatever the layout of the forward volume (nothing in the
            # backward reads the forward volume)
            st.grad_levels = [grad_out.new_zeros((st.B * st.H * st.W, 1, h, w))
                              for (h, w) in _level_shapes(st.H, st.W, st.num_levels)]
        lookup_backward(st.grad_levels, coords, grad_out, st.radius, st.H, st.W, st.lookup_mode)
        return grad_out.new_zeros(1), None, None


class CorrBlock:
    """models/raft/corr.py:26-106.

    Note on `corr_pyramid` / `get_corr_pyramid()`: they hand out the raw pyramid buffers (for the blocked layout a
    de-blocked copy of the two fine levels), WITHOUT an autograd graph -- in the reference these are differentiable
    matmul / avg_pool2d outputs (raft.py:161-163 exposes them through return_feat_maps).  A loss built on them gets
    no gradient; gradients flow through the lookups (`__call__`), whose coordinates are detached as in raft.py:188."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, compute_spatial=False, precision=None,
                 lookup_mode="grid_sample", layout="auto", backward_precision=None, storage=None):
        """precision: "tf32x3" (default: split-TF32 on the tensor cores, the accuracy of the reference's fp32
        matmul), "tf32" (one TF32 pass, half the build time, |err| <= 2^-10 * sum|f1 f2| / sqrt(C)) or "fp32"
        (CUDA cores).  None reads the environment variable B200CORR_VOLUME_PRECISION, so a caller that cannot
        pass arguments (the unmodified models/raft/raft.py:150-156) can still choose.
        backward_precision: precision of the two gradient contractions dF = G . F: "tf32" (tensor cores, operands
        truncated to TF32: relative 2^-10 per product, the default for the tensor-core forward precisions -- what
        cuBLAS computes for the reference when TF32 is allowed) or "fp32" (exact, CUDA cores; the default for
        precision="fp32").  B200CORR_VOLUME_BACKWARD_PRECISION overrides the default.
        layout: "auto" keeps the two fine levels in the blocked layout (8x8 tiles, include/b200corr.h) where the
        library supports the problem -- the lookups read them 1.4x faster; `corr_pyramid` / `get_corr_pyramid()`
        still hand out the reference's row-major tensors (converted on first use).  "rowmajor": as the reference.
        storage: "fp32" (default) or "fp16" -- the two blocked levels (94 % of the volume) kept as fp16: each value
        rounded once from the fp32 accumulator (relative 2^-11, the size of the TF32 input rounding; saturates at
        +-65504), half the bytes for the build to write and every lookup to read, half the memory.  Opt-in (None
        reads B200CORR_VOLUME_STORAGE); needs layout="auto"; gradients are unaffected (the backward never reads
        the forward volume)."""
        if storage is None:
            storage = os.environ.get("B200CORR_VOLUME_STORAGE", "fp32")
        if storage not in ("fp32", "fp16"):
            raise ValueError("CorrBlock: storage must be 'fp32' or 'fp16'")
        if storage == "fp16" and layout != "auto":
            raise ValueError("CorrBlock: fp16 storage needs layout='auto' (it exists for the blocked levels)")
        self.storage = storage
        self.num_levels = num_levels
        self.radius = radius
        self.compute_spatial = compute_spatial
        if precision is None:
            precision = os.environ.get("B200CORR_VOLUME_PRECISION", "tf32x3")
        if precision not in PRECISIONS:
            raise ValueError(f"CorrBlock: precision must be one of {sorted(PRECISIONS)}")
        self.precision = precision
        self.lookup_mode = lookup_mode
        if layout not in ("auto", "rowmajor"):
            raise ValueError("CorrBlock: layout must be 'auto' or 'rowmajor'")
        self._want_blocked = layout == "auto"
        self._levels, self._blocked, self._rowmajor = [], 0, None
        self._state = None
        self._handle = None
        if self.compute_spatial:
            # corr.py:33-54: FlowNetC-style 21x21 (dilation 2) correlation instead of all pairs
            self.upsample = torch.nn.Upsample(scale_factor=2, mode="nearest")
            out_corr = spatial_correlation_sample(fmap1, fmap2, kernel_size=1, patch_size=21, stride=1,
                                                  padding=0, dilation_patch=2)
            batch, ph, pw, h, w = out_corr.size()
            self.spatial_corr = out_corr.view(batch, ph * pw, h, w) / fmap1.size(1)
            corr = out_corr.view(batch * ph * pw, 1, h, w)
            self._levels.append(corr)
            for _ in range(self.num_levels - 1):
                corr = F.avg_pool2d(corr, 2, stride=2)
                self._levels.append(corr)
        else:
            self.B, self.C, self.H, self.W = fmap1.shape
            if backward_precision is None:
                backward_precision = os.environ.get("B200CORR_VOLUME_BACKWARD_PRECISION",
                                                    "fp32" if precision == "fp32" else "tf32")
            if backward_precision not in ("tf32", "fp32"):
                raise ValueError("CorrBlock: backward_precision must be 'tf32' or 'fp32'")
            self._state = _GradState(self.B, self.H, self.W, num_levels, radius, lookup_mode, precision,
                                     backward_precision)
            needs_grad = torch.is_grad_enabled() and (fmap1.requires_grad or fmap2.requires_grad)
            if needs_grad:
                self._handle = _VolumeFunction.apply(fmap1.contiguous(), fmap2.contiguous(), self)
            else:
                self._build(fmap1.detach(), fmap2.detach())

    def _build(self, fmap1, fmap2):
        if self._want_blocked:
            self._levels, self._blocked = allpairs_pyramid(fmap1, fmap2, self.num_levels, self.precision, blocked=True,
                                                           storage=self.storage)
        else:
            self._levels, self._blocked = allpairs_pyramid(fmap1, fmap2, self.num_levels, self.precision), 0
        self._rowmajor = None

    @property
    def corr_pyramid(self):
        """The pyramid in the reference's layout, corr.py:66-67: (B*H*W, 1, H_l, W_l) row-major tensors."""
        if self._blocked == 0:
            return self._levels
        if self._rowmajor is None:
            shapes = _level_shapes(self.H, self.W, self.num_levels)
            self._rowmajor = [deblock(v, *shapes[i]) if (self._blocked >> i) & 1 else v
                              for i, v in enumerate(self._levels)]
        return self._rowmajor

    def get_corr_pyramid(self):
        return self.corr_pyramid

    def get_spatial_corr(self):
        return self.spatial_corr if self.compute_spatial else None

    def __call__(self, coords):
        if self.compute_spatial:
            # corr.py:88-93: the pooled sampler outputs are upsampled back, coords are ignored
            batch, _, h1, w1 = coords.shape
            out_pyramid = []
            for i in range(self.num_levels):
                corr = self.corr_pyramid[i]
                for _ in range(i):
                    corr = self.upsample(corr)
                out_pyramid.append(corr.view(batch, h1, w1, -1))
            return torch.cat(out_pyramid, dim=-1).permute(0, 3, 1, 2).contiguous().float()
        coords = coords.detach().float()
        if self._handle is not None and torch.is_grad_enabled():
            return _LookupFunction.apply(self._handle, coords, self)
        return lookup_forward(self._levels, coords, self.radius, self.H, self.W, self.lookup_mode,
                              blocked_levels=self._blocked)

    def lookup_convc1(self, coords, weight, bias=None, relu=True, impl=None):
        """`F.relu(convc1(self(coords)))` (models/raft/update.py:104,111 on top of raft.py:189) on the tensor cores.
        impl="fused": ONE kernel, the (B, L*(2r+1)^2, H, W) lookup result stays in shared memory / TMEM and never
        touches HBM; impl="pipelined" (default, the faster one at RAFT's sizes): the plain lookup kernel followed by
        a tcgen05 1x1-convolution kernel that reads the lookup result out of the L2.  B200CORR_LOOKUP_CONVC1
        overrides the default.  Inference path: under grad mode with differentiable features or weights the
        unfused reference chain runs instead, so gradients are always those of the reference."""
        coords = coords.detach().float()
        n_out = weight.shape[0]
        needs_grad = torch.is_grad_enabled() and (self._handle is not None or weight.requires_grad or
                                                  (bias is not None and bias.requires_grad))
        fused_ok = (not self.compute_spatial and not needs_grad and n_out % 32 == 0 and n_out <= 256 and
                    self.B > 0 and 1 <= self.radius <= 4)
        if not fused_ok:
            out = F.conv2d(self(coords), weight.reshape(n_out, -1, 1, 1), bias)
            return F.relu(out) if relu else out
        if impl is None:
            impl = os.environ.get("B200CORR_LOOKUP_CONVC1", "pipelined")
        nin = self.num_levels * (2 * self.radius + 1) ** 2
        if impl == "pipelined" and nin % 4 == 0 and (self.H * self.W) % 4 == 0:
            # two kernels: the plain lookup, then the convolution as a tcgen05 GEMM reading the lookup result from
            # the L2 (chunks of 4 samples: 39.8 MB at 48x160, well inside the 126 MB L2)
            outs = []
            for b0 in range(0, self.B, 4):
                sl = slice(b0 * self.H * self.W, min(self.B, b0 + 4) * self.H * self.W)
                corr = lookup_forward([v[sl] for v in self._levels], coords[b0:b0 + 4], self.radius, self.H, self.W,
                                      self.lookup_mode, blocked_levels=self._blocked)
                outs.append(conv1x1_forward(corr, weight, bias, relu))
            return outs[0] if len(outs) == 1 else torch.cat(outs, 0)
        if any(v.dtype != torch.float32 for v in self._levels):
            # the one-kernel variant reads fp32 volumes only: plain lookup (which reads fp16 tiles) + convolution
            out = conv1x1_forward(self(coords), weight, bias, relu) if nin % 4 == 0 and (self.H * self.W) % 4 == 0 else None
            if out is None:
                out = F.conv2d(self(coords), weight.reshape(n_out, -1, 1, 1), bias)
                out = F.relu(out) if relu else out
            return out
        wp = prepare_convc1(weight, self.num_levels, self.radius)
        return lookup_convc1_forward(self._levels, coords, wp, bias, n_out, self.radius, self.H, self.W,
                                     self.lookup_mode, self._blocked, relu)

    @staticmethod
    def corr(fmap1, fmap2, precision="tf32x3"):
        """corr.py:98-106 -> (B, H, W, 1, H, W)."""
        B, C, H, W = fmap1.shape
        return allpairs_pyramid(fmap1, fmap2, 1, precision)[0].view(B, H, W, 1, H, W)


class AlternateCorrBlock:
    """models/raft/corr.py:109-137.  The NHWC copies are made once here instead of on every call.

    The block exists to avoid the (H*W)^2 volume.  From the level on where the pooled key map has at most
    `dense_max_keys` pixels (levels >= 2 at 48x160: 1/16 + 1/64 of the full volume) a 10x10 window covers
    most of the map anyway, so those levels are served from a small dense volume built once on the tensor
    cores (split TF32: fp32-level accuracy) and the bilinear lookup kernel; the fine levels keep the
    on-the-fly alt_cuda_corr kernel.  Same values; inference only (with autograd every level uses
    alt_cuda_corr, which carries the gradient).  dense_max_keys=0 restores the reference's structure."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, dense_max_keys=1024):
        self.num_levels = num_levels
        self.radius = radius
        needs_grad = torch.is_grad_enabled() and (fmap1.requires_grad or fmap2.requires_grad)
        self.pyramid = [(fmap1, fmap2)]
        for _ in range(self.num_levels):
            fmap1 = F.avg_pool2d(fmap1, 2, stride=2)
            fmap2 = F.avg_pool2d(fmap2, 2, stride=2)
            self.pyramid.append((fmap1, fmap2))
        self._f1 = self.pyramid[0][0].permute(0, 2, 3, 1).contiguous().float()
        self._f2 = [self.pyramid[i][1].permute(0, 2, 3, 1).contiguous().float() for i in range(num_levels)]
        # first level served from a dense volume (every coarser level is smaller still)
        self._dense_from = num_levels
        self._dense = []
        B = self._f1.shape[0]
        if not needs_grad and dense_max_keys > 0 and B > 0 and 1 <= radius <= 4:
            for i in range(num_levels):
                h, w = self.pyramid[i][1].shape[-2:]
                if h * w <= dense_max_keys and all(self.pyramid[j][1].shape[-1] % 4 == 0 and
                                                   min(self.pyramid[j][1].shape[-2:]) >= 1
                                                   for j in range(i, num_levels)):
                    self._dense_from = i
                    break
            f1 = self.pyramid[0][0].float()
            self._dense = [allpairs_volume_rect(f1, self.pyramid[j][1].float()) for j in range(self._dense_from, num_levels)]

    def __call__(self, coords):
        coords_nhwc = coords.permute(0, 2, 3, 1)
        B, H, W, _ = coords_nhwc.shape
        dim = self.pyramid[0][0].shape[1]
        n = (2 * self.radius + 1) ** 2
        corr_list = []
        for i in range(self._dense_from):
            coords_i = (coords_nhwc / 2 ** i).reshape(B, 1, H, W, 2).contiguous().float().detach()
            corr = _AltCorrFunction.apply(self._f1, self._f2[i], coords_i, self.radius)
            corr_list.append(corr.squeeze(1))
        parts = []
        if corr_list:
            parts.append(torch.stack(corr_list, dim=1).reshape(B, len(corr_list) * n, H, W) / math.sqrt(dim))
        if self._dense:
            # the dense volumes already carry the 1/sqrt(C) factor; "direct" = alt_cuda_corr's coordinates
            parts.append(lookup_forward(self._dense, coords.detach().float(), self.radius, H, W, "direct",
                                        first_level=self._dense_from))
        if len(parts) == 1:
            return parts[0]
        return torch.cat(parts, dim=1) if parts else coords.new_zeros((B, 0, H, W))
