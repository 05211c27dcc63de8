"""Drop-ins for the two native extensions of the reference's FlowNet2 family.

`channelnorm_cuda` and `resample2d_cuda` below have the functions of the reference's pybind modules
(models/channelnorm_package/channelnorm_cuda.cc:29-32, models/resample2d_package/resample2d_cuda.cc:28-31): same names,
same argument order, outputs written into the tensors the caller passes, return value 1.  Registered under those module
names (`shims.install_reference_shims`), the reference's own wrappers -- channelnorm.py, resample2d.py -- import them
unmodified.  `ChannelNorm` / `Resample2d` mirror those wrappers for direct use.
"""
import types

import torch
from torch.autograd import Function
from torch.nn.modules.module import Module

from . import _lib


def _check(who, *ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError(f"{who}: CUDA tensors only (this build has no CPU path)")
        if t.dtype != torch.float32:
            raise RuntimeError(f"{who}: float32 only (as the reference's kernels)")
        if not t.is_contiguous():
            raise RuntimeError(f"{who}: tensors must be contiguous")
        if t.device != ts[0].device:
            raise RuntimeError(f"{who}: tensors must be on the same device")


def _call(name, dev, *args):
    with torch.cuda.device(dev):
        code = getattr(_lib.lib(), name)(*args, _lib.current_stream(dev))
    _lib.check(code, name)
    return 1


def _cn_forward(input1, output, norm_deg):
    _check("channelnorm forward", input1, output)
    B, C, H, W = input1.shape
    if tuple(output.shape) != (B, 1, H, W):
        raise RuntimeError("channelnorm forward: output must be (B, 1, H, W)")
    return _call("b200corr_channelnorm_forward", input1.device, _lib.ptr(input1), _lib.ptr(output), B, C, H, W, int(norm_deg))


def _cn_backward(input1, output, gradOutput, gradInput1, norm_deg):
    gradOutput = gradOutput.contiguous()
    _check("channelnorm backward", input1, output, gradOutput, gradInput1)
    B, C, H, W = input1.shape
    if gradInput1.shape != input1.shape or tuple(gradOutput.shape) != (B, 1, H, W):
        raise RuntimeError("channelnorm backward: bad shapes")
    return _call("b200corr_channelnorm_backward", input1.device, _lib.ptr(input1), _lib.ptr(output), _lib.ptr(gradOutput),
                 _lib.ptr(gradInput1), B, C, H, W, int(norm_deg))


def _rs_forward(input1, input2, output, kernel_size, bilinear):
    _check("resample2d forward", input1, input2, output)
    B, C, H, W = input1.shape
    if tuple(input2.shape) != (B, 2, H, W) or output.shape != input1.shape:
        raise RuntimeError("resample2d forward: need input1 (B,C,H,W), input2 (B,2,H,W) and an output like input1")
    return _call("b200corr_resample2d_forward", input1.device, _lib.ptr(input1), _lib.ptr(input2), _lib.ptr(output),
                 B, C, H, W, int(kernel_size), int(bool(bilinear)))


def _rs_backward(input1, input2, gradOutput, gradInput1, gradInput2, kernel_size, bilinear):
    gradOutput = gradOutput.contiguous()
    _check("resample2d backward", input1, input2, gradOutput, gradInput1, gradInput2)
    B, C, H, W = input1.shape
    if gradOutput.shape != input1.shape or gradInput1.shape != input1.shape or gradInput2.shape != input2.shape:
        raise RuntimeError("resample2d backward: bad shapes")
    return _call("b200corr_resample2d_backward", input1.device, _lib.ptr(input1), _lib.ptr(input2), _lib.ptr(gradOutput),
                 _lib.ptr(gradInput1), _lib.ptr(gradInput2), B, C, H, W, int(kernel_size), int(bool(bilinear)))


channelnorm_cuda = types.SimpleNamespace(forward=_cn_forward, backward=_cn_backward)
resample2d_cuda = types.SimpleNamespace(forward=_rs_forward, backward=_rs_backward)


class ChannelNormFunction(Function):
    """models/channelnorm_package/channelnorm.py:6-32."""

    @staticmethod
    def forward(ctx, input1, norm_deg=2):
        input1 = input1.contiguous()
        b, _, h, w = input1.size()
        output = input1.new_empty((b, 1, h, w))
        channelnorm_cuda.forward(input1, output, norm_deg)
        ctx.save_for_backward(input1, output)
        ctx.norm_deg = norm_deg
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input1, output = ctx.saved_tensors
        grad_input1 = torch.empty_like(input1)
        channelnorm_cuda.backward(input1, output, grad_output, grad_input1, ctx.norm_deg)
        return grad_input1, None


class ChannelNorm(Module):
    def __init__(self, norm_deg=2):
        super().__init__()
        self.norm_deg = norm_deg

    def forward(self, input1):
        return ChannelNormFunction.apply(input1, self.norm_deg)


class Resample2dFunction(Function):
    """models/resample2d_package/resample2d.py:7-47."""

    @staticmethod
    def forward(ctx, input1, input2, kernel_size=1, bilinear=True):
        input1 = input1.contiguous()
        input2 = input2.contiguous()
        ctx.save_for_backward(input1, input2)
        ctx.kernel_size = kernel_size
        ctx.bilinear = bilinear
        output = torch.empty_like(input1)
        resample2d_cuda.forward(input1, input2, output, kernel_size, bilinear)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        grad_input1 = torch.empty_like(input1)
        grad_input2 = torch.empty_like(input2)
        resample2d_cuda.backward(input1, input2, grad_output, grad_input1, grad_input2, ctx.kernel_size, ctx.bilinear)
        return grad_input1, grad_input2, None, None


class Resample2d(Module):
    def __init__(self, kernel_size=1, bilinear=True):
        super().__init__()
        self.kernel_size = kernel_size
        self.bilinear = bilinear

    def forward(self, input1, input2):
        return Resample2dFunction.apply(input1.contiguous(), input2, self.kernel_size, self.bilinear)
