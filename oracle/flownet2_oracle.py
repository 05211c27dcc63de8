"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's FlowNet2 natives.

channelnorm: models/channelnorm_package/channelnorm_kernel.cu:19-96.  resample2d (kernel_size 1):
models/resample2d_package/resample2d_kernel.cu:17-195.  The reference ships these as CUDA-only extensions: resample2d
compiles unmodified for sm_100a (oracle/_ref/ref_resample2d_cuda, checked against this file on the GPU box);
channelnorm does not compile against torch 2.11 (parity unpinned by the reference itself: pinned against the formula
in the kernel source).  Only tests/ may import this module.
"""
import numpy as np


def channelnorm_forward(x):
    return np.sqrt((x.astype(np.float64) ** 2).sum(1, keepdims=True)).astype(np.float32)


def channelnorm_backward(x, out, gout):
    return (gout.astype(np.float64) * x.astype(np.float64) / (out.astype(np.float64) + 1e-9)).astype(np.float32)


def _geometry(flow, H, W):
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    xf = (xs[None] + flow[:, 0]).astype(np.float32)
    yf = (ys[None] + flow[:, 1]).astype(np.float32)
    fx, fy = np.floor(xf), np.floor(yf)
    xL = np.clip(fx.astype(np.int64), 0, W - 1)
    xR = np.clip((fx + 1).astype(np.int64), 0, W - 1)
    yT = np.clip(fy.astype(np.int64), 0, H - 1)
    yB = np.clip((fy + 1).astype(np.int64), 0, H - 1)
    return xf, yf, fx, fy, xL, xR, yT, yB


def resample2d_forward(x, flow, bilinear=True):
    B, C, H, W = x.shape
    xf, yf, fx, fy, xL, xR, yT, yB = _geometry(flow, H, W)
    bi = np.arange(B)[:, None, None]
    out = np.zeros_like(x)
    if not bilinear:
        xN = np.clip(np.floor(xf + np.float32(0.5)).astype(np.int64), 0, W - 1)
        yN = np.clip(np.floor(yf + np.float32(0.5)).astype(np.int64), 0, H - 1)
        for c in range(C):
            out[:, c] = x[bi, c, yN, xN]
        return out
    a = (xf - fx).astype(np.float64)
    b = (yf - fy).astype(np.float64)
    for c in range(C):
        v = np.zeros((B, H, W), np.float32)
        for w, yy, xx in (((1 - a) * (1 - b), yT, xL), (a * (1 - b), yT, xR), ((1 - a) * b, yB, xL), (a * b, yB, xR)):
            v = v + (w * x[bi, c, yy, xx].astype(np.float64)).astype(np.float32)
        out[:, c] = v
    return out


def resample2d_backward(x, flow, gout):
    """(grad_input1, grad_input2); the image gradient uses `xf - int(xf)` (truncation) as the reference does (:96-97)."""
    B, C, H, W = x.shape
    xf, yf, fx, fy, xL, xR, yT, yB = _geometry(flow, H, W)
    a = (xf - np.trunc(xf)).astype(np.float64)
    b = (yf - np.trunc(yf)).astype(np.float64)
    g1 = np.zeros(x.shape, np.float64)
    bi = np.broadcast_to(np.arange(B)[:, None, None], (B, H, W))
    for c in range(C):
        g = gout[:, c].astype(np.float64)
        for w, yy, xx in (((1 - a) * (1 - b), yT, xL), (a * (1 - b), yT, xR), ((1 - a) * b, yB, xL), (a * b, yB, xR)):
            np.add.at(g1[:, c], (bi, yy, xx), w * g)
    gx = (1.0 - (xf - fx)).astype(np.float64)    # gamma of the odd (dy) channel
    gy = (1.0 - (yf - fy)).astype(np.float64)    # gamma of the even (dx) channel
    g2 = np.zeros((B, 2, H, W), np.float64)
    for c in range(C):
        g = gout[:, c].astype(np.float64)
        v = x[:, c].astype(np.float64)
        TL, TR, BL, BR = v[bi, yT, xL], v[bi, yT, xR], v[bi, yB, xL], v[bi, yB, xR]
        g2[:, 0] += gy * g * (TR - TL) + (1 - gy) * g * (BR - BL)
        g2[:, 1] += gx * g * (BL - TL) + (1 - gx) * g * (BR - TR)
    return g1.astype(np.float32), g2.astype(np.float32)
