"""Fused FlowNetC merge block.

The reference's FlowNetC family merges its two streams with (models/FlowNetC.py:133-147,
models/submodules.py:124-138; the same lines exist in FlowNetC_flexible_larger_field.py and the
FlowNet2 copies):

    out_corr   = correlate(out_conv3a, out_conv3b)        # sampler (patch 21, dilation_patch 2), view, / C
    out_corr   = self.corr_activation(out_corr)           # LeakyReLU(0.1)
    in_conv3_1 = torch.cat((out_conv_redir, out_corr), 1)

`correlate_merge(out_conv3a, out_conv3b, out_conv_redir)` returns the same `in_conv3_1` from one
correlation kernel that applies `/ C` and the activation to its register accumulators and stores into the
channel slice of the concat tensor; its backward recovers the LeakyReLU mask from the stored output.
Values are those of the composed reference ops on CUDA (torch divides by a Python scalar as a multiplication
by the fp32 reciprocal; for the power-of-two channel counts of the call sites that is the exact quotient).
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from . import backend


def _check(who, *ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError(f"{who}: CUDA tensors only (this build has no CPU path)")
        if t.dtype != torch.float32:
            raise RuntimeError(f"{who}: float32 only")
        if t.device != ts[0].device:
            raise RuntimeError(f"{who}: inputs must be on the same device")


def merge_supported(input1, patch_size=21, dilation_patch=2):
    B, C, H, W = input1.shape
    return bool(_lib.lib().b200corr_merge_supported(B, C, H, W, patch_size, dilation_patch))


def merge_forward(input1, input2, merged, c_off, patch_size=21, dilation_patch=2, negative_slope=0.1):
    """Writes leaky_relu(correlate(input1, input2), negative_slope) into merged[:, c_off : c_off + patch_size**2]."""
    _check("merge_forward", input1, input2, merged)
    if input1.dim() != 4 or input1.shape != input2.shape or not (input1.is_contiguous() and input2.is_contiguous()):
        raise RuntimeError("merge_forward: input1 and input2 must be contiguous (B, C, H, W) tensors of the same shape")
    B, C, H, W = input1.shape
    if merged.dim() != 4 or merged.shape[0] != B or tuple(merged.shape[2:]) != (H, W) or not merged.is_contiguous():
        raise RuntimeError("merge_forward: `merged` must be a contiguous (B, c_total, H, W) tensor")
    with torch.cuda.device(input1.device):
        code = _lib.lib().b200corr_merge_forward(_lib.ptr(input1), _lib.ptr(input2), _lib.ptr(merged), B, C, H, W,
                                                 patch_size, dilation_patch, merged.shape[1], c_off,
                                                 float(negative_slope), _lib.current_stream(input1.device))
    _lib.check(code, "b200corr_merge_forward")
    return merged


def merge_backward(input1, input2, merged, grad_merged, c_off, patch_size=21, dilation_patch=2, negative_slope=0.1):
    """(grad_input1, grad_input2) from the gradient of the concat tensor."""
    _check("merge_backward", input1, input2, merged, grad_merged)
    grad_merged = grad_merged.contiguous()
    B, C, H, W = input1.shape
    if grad_merged.shape != merged.shape:
        raise RuntimeError("merge_backward: grad_merged must have the shape of merged")
    L = _lib.lib()
    g1 = torch.empty_like(input1)
    g2 = torch.empty_like(input2)
    nbytes = L.b200corr_merge_backward_scratch_bytes(B, H, W, patch_size)
    scratch = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=input1.device)
    hyper = (1, 1, patch_size, patch_size, 0, 0, 1, 1, dilation_patch, dilation_patch, 1, 1)
    with torch.cuda.device(input1.device):
        plan = backend._backward_plan(input1.device, B, C, H, W, hyper, 0)
        ws, ws_bytes = (_lib.ptr(plan), plan.numel() * 4) if plan is not False else (None, 0)
        code = L.b200corr_merge_backward(_lib.ptr(input1), _lib.ptr(input2), _lib.ptr(merged), _lib.ptr(grad_merged),
                                         _lib.ptr(g1), _lib.ptr(g2), _lib.ptr(scratch), ws, ws_bytes, B, C, H, W,
                                         patch_size, dilation_patch, merged.shape[1], c_off, float(negative_slope),
                                         _lib.current_stream(input1.device))
    _lib.check(code, "b200corr_merge_backward")
    return g1, g2


class CorrelateMergeFunction(Function):
    """forward(input1, input2, patch_size, dilation_patch, negative_slope, n_before, *others): the concat of
    others[:n_before], the activated correlation and others[n_before:] along dim 1."""

    @staticmethod
    def forward(ctx, input1, input2, patch_size, dilation_patch, negative_slope, n_before, *others):
        _check("correlate_merge", input1, input2, *others)
        B, _, H, W = input1.shape
        for t in others:
            if t.dim() != 4 or t.shape[0] != B or tuple(t.shape[2:]) != (H, W):
                raise RuntimeError("correlate_merge: concatenated tensors must be (B, c, H, W) like the feature maps")
        widths = [t.shape[1] for t in others]
        c_off = sum(widths[:n_before])
        P2 = patch_size * patch_size
        merged = torch.empty((B, sum(widths) + P2, H, W), dtype=torch.float32, device=input1.device)
        at = 0
        for i, t in enumerate(others):
            if i == n_before:
                at += P2
            merged[:, at:at + widths[i]].copy_(t)
            at += widths[i]
        merge_forward(input1, input2, merged, c_off, patch_size, dilation_patch, negative_slope)
        ctx.save_for_backward(input1, input2, merged)
        ctx.hyper = (c_off, patch_size, dilation_patch, negative_slope, n_before, widths)
        return merged

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_merged):
        input1, input2, merged = ctx.saved_tensors
        c_off, patch_size, dilation_patch, negative_slope, n_before, widths = ctx.hyper
        g1 = g2 = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            g1, g2 = merge_backward(input1, input2, merged, grad_merged, c_off, patch_size, dilation_patch,
                                    negative_slope)
        g_others = []
        at = 0
        for i, w in enumerate(widths):
            if i == n_before:
                at += patch_size * patch_size
            g_others.append(grad_merged[:, at:at + w] if ctx.needs_input_grad[6 + i] else None)
            at += w
        return (g1, g2, None, None, None, None, *g_others)


def correlate_merge(input1, input2, redir=None, patch_size=21, dilation_patch=2, negative_slope=0.1, after=()):
    """cat((redir, leaky_relu(correlate(input1, input2), negative_slope), *after), 1).

    FlowNetC (FlowNetC.py:133-147): `correlate_merge(conv3a, conv3b, conv_redir(conv3a))`.
    PWC-Net (PWCNet.py:286-292, `x = cat((corr5, c15, up_flow6, up_feat6), 1)` after `leakyRELU(corr(c15, warp5))`):
    `correlate_merge(c15, warp5, None, 9, 1, 0.1, after=(c15, up_flow6, up_feat6))`."""
    before = () if redir is None else (redir,)
    return CorrelateMergeFunction.apply(input1.contiguous(), input2.contiguous(), patch_size, dilation_patch,
                                        negative_slope, len(before), *before, *after)
