"""ctypes front-end of oracle/sampler_oracle.c (CPU restatement of the reference sampler).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  numpy in, numpy out.
Reference: models/Pytorch-Correlation-extension/Correlation_Module/correlation.cpp:75-178 and the
Python wrapper spatial_correlation_sampler/spatial_correlation_sampler.py:8-116 (argument order,
`_pair` handling of the six hyper-parameters).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsampler_oracle.so")
_lib = None


def build():
    src = os.path.join(_HERE, "sampler_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _pair(v):
    return (int(v), int(v)) if not isinstance(v, (tuple, list)) else (int(v[0]), int(v[1]))


def out_size(size, pad, k, dil, stride):
    """correlation.cpp:90-94"""
    return (size + 2 * pad - ((k - 1) * dil + 1)) // stride + 1


def _suffix(dtype):
    return {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}[np.dtype(dtype)]


def forward(in1, in2, kernel_size=1, patch_size=1, stride=1, padding=0, dilation=1, dilation_patch=1):
    """out[B, patchH, patchW, oH, oW]; SURVEY.md section 8(0) S1."""
    in1 = np.ascontiguousarray(in1)
    in2 = np.ascontiguousarray(in2, dtype=in1.dtype)
    B, C, H, W = in1.shape
    kH, kW = _pair(kernel_size)
    pH, pW = _pair(patch_size)
    sH, sW = _pair(stride)
    padH, padW = _pair(padding)
    dH, dW = _pair(dilation)
    dpH, dpW = _pair(dilation_patch)
    oH, oW = out_size(H, padH, kH, dH, sH), out_size(W, padW, kW, dW, sW)
    out = np.empty((B, pH, pW, oH, oW), dtype=in1.dtype)
    fn = getattr(_load(), "sampler_oracle_forward_" + _suffix(in1.dtype))
    fn(in1.ctypes.data_as(ctypes.c_void_p), in2.ctypes.data_as(ctypes.c_void_p),
       out.ctypes.data_as(ctypes.c_void_p),
       *map(ctypes.c_int, (B, C, H, W, kH, kW, pH, pW, padH, padW, dH, dW, dpH, dpW, sH, sW)))
    return out


def backward(in1, in2, grad_out, kernel_size=1, patch_size=1, stride=1, padding=0, dilation=1,
             dilation_patch=1):
    """(grad_in1, grad_in2), each [B, C, H, W]."""
    in1 = np.ascontiguousarray(in1)
    in2 = np.ascontiguousarray(in2, dtype=in1.dtype)
    grad_out = np.ascontiguousarray(grad_out, dtype=in1.dtype)
    B, C, H, W = in1.shape
    kH, kW = _pair(kernel_size)
    pH, pW = _pair(patch_size)
    sH, sW = _pair(stride)
    padH, padW = _pair(padding)
    dH, dW = _pair(dilation)
    dpH, dpW = _pair(dilation_patch)
    oH, oW = grad_out.shape[3], grad_out.shape[4]
    assert grad_out.shape == (B, pH, pW, oH, oW)
    g1 = np.empty_like(in1)
    g2 = np.empty_like(in2)
    fn = getattr(_load(), "sampler_oracle_backward_" + _suffix(in1.dtype))
    fn(in1.ctypes.data_as(ctypes.c_void_p), in2.ctypes.data_as(ctypes.c_void_p),
       grad_out.ctypes.data_as(ctypes.c_void_p), g1.ctypes.data_as(ctypes.c_void_p),
       g2.ctypes.data_as(ctypes.c_void_p),
       *map(ctypes.c_int, (B, C, H, W, oH, oW, kH, kW, pH, pW, padH, padW, dH, dW, dpH, dpW, sH, sW)))
    return g1, g2
