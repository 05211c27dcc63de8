"""PWC-Net's `warp()` as one kernel.

Reference: `PWCDCNet.warp(x, flo)`, models/PWCNet.py:164-204 -- mesh grid + flow, normalised to [-1, 1],
`grid_sample` of the features and of an all-ones tensor, `mask = (mask >= 0.0001)`, `output * mask`.  It feeds
`input2` of every correlation below the top pyramid level (:293-294, 308-309, 323-324, 338-339).
`warp(x, flo)` returns the same tensor and is differentiable w.r.t. both arguments (the thresholded mask passes
no gradient, as in the reference).
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib


def _check(who, *ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError(f"{who}: CUDA tensors only (this build has no CPU path)")
        if t.dtype != torch.float32:
            raise RuntimeError(f"{who}: float32 only")
        if t.device != ts[0].device:
            raise RuntimeError(f"{who}: inputs must be on the same device")


class WarpFunction(Function):
    @staticmethod
    def forward(ctx, x, flo):
        _check("warp", x, flo)
        if x.dim() != 4 or flo.dim() != 4 or flo.shape[1] != 2 or flo.shape[0] != x.shape[0] or flo.shape[2:] != x.shape[2:]:
            raise RuntimeError("warp: x must be (B, C, H, W) and flo (B, 2, H, W)")
        x = x.contiguous()
        flo = flo.contiguous()
        B, C, H, W = x.shape
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            code = _lib.lib().b200corr_warp_forward(_lib.ptr(x), _lib.ptr(flo), _lib.ptr(out), B, C, H, W,
                                                    _lib.current_stream(x.device))
        _lib.check(code, "b200corr_warp_forward")
        ctx.save_for_backward(x, flo)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        x, flo = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        B, C, H, W = x.shape
        gx = torch.empty_like(x)
        gf = torch.empty_like(flo)
        with torch.cuda.device(x.device):
            code = _lib.lib().b200corr_warp_backward(_lib.ptr(x), _lib.ptr(flo), _lib.ptr(grad_out), _lib.ptr(gx),
                                                     _lib.ptr(gf), B, C, H, W, _lib.current_stream(x.device))
        _lib.check(code, "b200corr_warp_backward")
        return gx, gf


def warp(x, flo):
    """Warp `x` (features of image 2) back to image 1 along the flow `flo` -- PWCNet.py:164-204."""
    return WarpFunction.apply(x, flo)
