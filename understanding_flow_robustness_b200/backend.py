"""Drop-in for the reference's pybind module `spatial_correlation_sampler_backend`.

Reference: models/Pytorch-Correlation-extension/Correlation_Module/correlation_sampler.cpp:59-129
(`forward(input1, input2, kH, kW, patchH, patchW, padH, padW, dilationH, dilationW,
dilation_patchH, dilation_patchW, dH, dW)` and `backward(input1, input2, grad_output, <same 12>)`).
Same argument order and meaning; the checks the reference does with TORCH_CHECK
(CHECK_CUDA / CHECK_CONTIGUOUS / CHECK_SAME_DEVICE, :31-34,69-75,101-107) raise RuntimeError here
too, plus shape/dtype validation the reference leaves out.  CUDA tensors only: the reference's CPU
branch (:76-86) has no counterpart in this package.
"""
import collections

import torch

from . import _lib

_DTYPES = {torch.float32: 0, torch.float64: 1, torch.float16: 2, torch.bfloat16: 3}


def _check_inputs(who, *tensors):
    first = tensors[0]
    for t in tensors:
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{who}: expected torch.Tensor, got {type(t)}")
        if not t.is_cuda:
            raise RuntimeError(f"{who}: input must be a CUDA tensor (this build has no CPU path)")
        if not t.is_contiguous():
            raise RuntimeError(f"{who}: input must be contiguous")
        if t.device != first.device:
            raise RuntimeError(f"{who}: inputs must be on the same device")
        if t.dtype != first.dtype:
            raise RuntimeError(f"{who}: inputs must have the same dtype")
    if first.dtype not in _DTYPES:
        raise RuntimeError(f"{who}: unsupported dtype {first.dtype} (float32 / float64 / float16 / bfloat16)")
    return _DTYPES[first.dtype]


def output_size(size, pad, kernel, dilation, stride):
    """correlation_cuda_kernel.cu:249-253 / correlation.cpp:90-94"""
    return (size + 2 * pad - ((kernel - 1) * dilation + 1)) // stride + 1


def forward(input1, input2, kH, kW, patchH, patchW, padH, padW, dilationH, dilationW,
            dilation_patchH, dilation_patchW, dH, dW, out=None):
    dt = _check_inputs("correlation forward", input1, input2)
    if input1.dim() != 4 or input1.shape != input2.shape:
        raise RuntimeError("correlation forward: input1 and input2 must both be (B, C, H, W) of the same shape")
    B, C, H, W = input1.shape
    oH = output_size(H, padH, kH, dilationH, dH)
    oW = output_size(W, padW, kW, dilationW, dW)
    if oH < 0 or oW < 0:
        raise RuntimeError("correlation forward: kernel does not fit the padded input")
    if out is None:
        out = torch.empty((B, patchH, patchW, oH, oW), dtype=input1.dtype, device=input1.device)
    elif (tuple(out.shape) != (B, patchH, patchW, oH, oW) or out.dtype != input1.dtype or out.device != input1.device
          or not out.is_contiguous()):
        raise RuntimeError("correlation forward: `out` must be a contiguous (B, patchH, patchW, oH, oW) tensor like the inputs")
    with torch.cuda.device(input1.device):
        code = _lib.lib().b200corr_sampler_forward(
            _lib.ptr(input1), _lib.ptr(input2), _lib.ptr(out), None, 0, B, C, H, W, kH, kW, patchH,
            patchW, padH, padW, dilationH, dilationW, dilation_patchH, dilation_patchW, dH, dW, dt,
            _lib.current_stream(input1.device))
    _lib.check(code, "b200corr_sampler_forward")
    return out


_PLANS = collections.OrderedDict()      # (device, shape, hyper, dtype) -> device plan, least recently used first
_PLAN_CAPACITY = 64


def _backward_plan(device, B, C, H, W, hyper, dt):
    """Device copy of the library's backward schedule, built once per (device, problem shape).

    The upload goes through a pinned buffer and is complete before this returns (the plan is then read by
    kernels on whatever stream a later call uses); eviction is least-recently-used and a plan is only dropped
    after the device has finished everything queued so far.  A cold cache inside a CUDA graph capture is an
    error: warm the shape up once before capturing."""
    key = (device, B, C, H, W, hyper, dt)
    plan = _PLANS.get(key)
    if plan is not None:
        _PLANS.move_to_end(key)
        return plan
    L = _lib.lib()
    nbytes = L.b200corr_sampler_backward_workspace_bytes(B, C, H, W, *hyper, dt)
    if nbytes == 0:
        plan = False
    else:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("correlation backward: first call for this shape inside a CUDA graph capture; run the "
                               "backward once before capturing (the schedule is uploaded on first use)")
        host = torch.empty(nbytes // 4, dtype=torch.int32).pin_memory()
        _lib.check(L.b200corr_sampler_backward_plan(B, C, H, W, *hyper, dt, host.data_ptr(), nbytes),
                   "b200corr_sampler_backward_plan")
        plan = host.to(device, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
    if len(_PLANS) >= _PLAN_CAPACITY:
        torch.cuda.synchronize(device)              # nothing in flight still reads the evicted plan
        _PLANS.popitem(last=False)
    _PLANS[key] = plan
    return plan


def backward(input1, input2, grad_output, kH, kW, patchH, patchW, padH, padW, dilationH, dilationW,
             dilation_patchH, dilation_patchW, dH, dW, out=None):
    grad_output = grad_output.contiguous()
    dt = _check_inputs("correlation backward", input1, input2, grad_output)
    B, C, H, W = input1.shape
    oH = output_size(H, padH, kH, dilationH, dH)
    oW = output_size(W, padW, kW, dilationW, dW)
    if tuple(grad_output.shape) != (B, patchH, patchW, oH, oW):
        raise RuntimeError(f"correlation backward: grad_output shape {tuple(grad_output.shape)} != "
                           f"{(B, patchH, patchW, oH, oW)}")
    if out is None:
        g1 = torch.empty_like(input1)
        g2 = torch.empty_like(input2)
    else:
        g1, g2 = out
        for g in (g1, g2):
            if g.shape != input1.shape or g.dtype != input1.dtype or g.device != input1.device or not g.is_contiguous():
                raise RuntimeError("correlation backward: `out` must be two contiguous tensors like input1")
    hyper = (kH, kW, patchH, patchW, padH, padW, dilationH, dilationW, dilation_patchH, dilation_patchW, dH, dW)
    with torch.cuda.device(input1.device):
        plan = _backward_plan(input1.device, B, C, H, W, hyper, dt)
        ws, ws_bytes = (_lib.ptr(plan), plan.numel() * 4) if plan is not False else (None, 0)
        code = _lib.lib().b200corr_sampler_backward(
            _lib.ptr(input1), _lib.ptr(input2), _lib.ptr(grad_output), _lib.ptr(g1), _lib.ptr(g2),
            ws, ws_bytes, B, C, H, W, kH, kW, patchH, patchW, padH, padW, dilationH, dilationW,
            dilation_patchH, dilation_patchW, dH, dW, dt, _lib.current_stream(input1.device))
    _lib.check(code, "b200corr_sampler_backward")
    return [g1, g2]
