"""`spatial_correlation_sample` / `SpatialCorrelationSampler` -- the reference's operator API.

Mirrors models/Pytorch-Correlation-extension/Correlation_Module/spatial_correlation_sampler/
spatial_correlation_sampler.py:8-147: same function / class names, same keyword arguments (each an
int or an (H, W) pair), same 5-D output (B, patchH, patchW, oH, oW), a once-differentiable autograd
Function that saves (input1, input2) and returns two gradients plus six None.  The backend is the
C-ABI library instead of the reference's pybind extension.
"""
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from . import backend as correlation


def spatial_correlation_sample(input1, input2, kernel_size=1, patch_size=1, stride=1, padding=0,
                               dilation=1, dilation_patch=1):
    """Correlate every `kernel_size` window of input1 with the `patch_size` x `patch_size` grid of
    windows of input2 displaced by multiples of `dilation_patch`
    (reference docstring: spatial_correlation_sampler.py:18-40)."""
    return SpatialCorrelationSamplerFunction.apply(input1, input2, kernel_size, patch_size, stride,
                                                   padding, dilation, dilation_patch)


class SpatialCorrelationSamplerFunction(Function):
    @staticmethod
    def forward(ctx, input1, input2, kernel_size=1, patch_size=1, stride=1, padding=0, dilation=1,
                dilation_patch=1):
        ctx.save_for_backward(input1, input2)
        ctx.hyper = (*_pair(kernel_size), *_pair(patch_size), *_pair(padding), *_pair(dilation),
                     *_pair(dilation_patch), *_pair(stride))
        return correlation.forward(input1, input2, *ctx.hyper)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        grad_input1, grad_input2 = correlation.backward(input1, input2, grad_output, *ctx.hyper)
        return grad_input1, grad_input2, None, None, None, None, None, None


class SpatialCorrelationSampler(nn.Module):
    def __init__(self, kernel_size=1, patch_size=1, stride=1, padding=0, dilation=1, dilation_patch=1):
        super().__init__()
        self.kernel_size = kernel_size
        self.patch_size = patch_size
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.dilation_patch = dilation_patch

    def forward(self, input1, input2):
        return SpatialCorrelationSamplerFunction.apply(input1, input2, self.kernel_size, self.patch_size,
                                                       self.stride, self.padding, self.dilation,
                                                       self.dilation_patch)
