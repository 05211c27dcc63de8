"""Universal-patch placement on the GPU.

Reference: `circle_transform` (patch_attacks/utils_patch.py:257-358: brightness jitter + clip, `patch * mask`,
scipy zoom / rotate of order 1, paste at a random location) and the composition
`adv = (1 - mask_canvas) * img + patch_canvas` for both frames (patch_attacks/main.py:537-542), all on the host
with numpy once per image pair.  `compose_adversarial(img1, img2, patch, mask, placements)` does the same for a
batch of pairs in one kernel and is differentiable w.r.t. the canonical patch (one gather kernel, no atomics);
the images are constants of the attack and get no gradient.

    placements : (n, 5) = (scale, angle [rad], centre x, centre y, brightness offset) per pair
    returns    : adv1, adv2 in the memory format of img1

`attack.compose_torch` is the same function written with affine_grid / grid_sample; tests compare both.
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib


def _strides(t):
    n, c, h, w = t.stride()
    return n, c, h, w


def _dense_nchw_or_nhwc(t):
    return t.is_contiguous() or t.is_contiguous(memory_format=torch.channels_last)


class ComposeFunction(Function):
    @staticmethod
    def forward(ctx, img1, img2, patch, mask, placements):
        for t in (img1, img2, patch, mask, placements):
            if not t.is_cuda:
                raise RuntimeError("compose_adversarial: CUDA tensors only (this build has no CPU path)")
            if t.dtype != torch.float32:
                raise RuntimeError("compose_adversarial: float32 only")
            if t.device != img1.device:
                raise RuntimeError("compose_adversarial: inputs must be on the same device")
        if img1.dim() != 4 or img1.shape[1] != 3 or img2.shape != img1.shape:
            raise RuntimeError("compose_adversarial: img1 / img2 must be (n, 3, H, W) of the same shape")
        if not _dense_nchw_or_nhwc(img1):
            img1 = img1.contiguous()
        if img2.stride() != img1.stride():
            img2 = img2.contiguous(memory_format=torch.channels_last if not img1.is_contiguous() else torch.contiguous_format)
            if img2.stride() != img1.stride():            # degenerate sizes: fall back to one explicit layout
                img1, img2 = img1.contiguous(), img2.contiguous()
        n, _, H, W = img1.shape
        p = patch.shape[-1]
        if patch.numel() != 3 * p * p or mask.numel() != p * p or patch.shape[-2] != p:
            raise RuntimeError("compose_adversarial: patch must be (1, 3, p, p) and mask (1, 1, p, p)")
        if tuple(placements.shape) != (n, 5):
            raise RuntimeError("compose_adversarial: placements must be (n, 5)")
        patch_c, mask_c, pl = patch.contiguous(), mask.contiguous(), placements.contiguous()
        adv1, adv2 = torch.empty_like(img1), torch.empty_like(img2)      # preserve_format: same strides
        assert adv1.stride() == img1.stride() and adv2.stride() == img1.stride()
        with torch.cuda.device(img1.device):
            code = _lib.lib().b200corr_patch_compose_forward(
                _lib.ptr(img1), _lib.ptr(img2), _lib.ptr(patch_c), _lib.ptr(mask_c), _lib.ptr(pl), _lib.ptr(adv1),
                _lib.ptr(adv2), n, H, W, p, *_strides(img1), _lib.current_stream(img1.device))
        _lib.check(code, "b200corr_patch_compose_forward")
        ctx.save_for_backward(img1, img2, patch_c, mask_c, pl)
        ctx.patch_shape = patch.shape
        return adv1, adv2

    @staticmethod
    @once_differentiable
    def backward(ctx, g1, g2):
        img1, img2, patch, mask, pl = ctx.saved_tensors
        n, _, H, W = img1.shape
        p = patch.shape[-1]
        if n == 0:                       # no pairs: no contribution (strides of empty tensors carry no layout)
            return None, None, torch.zeros(ctx.patch_shape, dtype=torch.float32, device=patch.device), None, None
        fmt = torch.contiguous_format if img1.is_contiguous() else torch.channels_last
        g1 = g1.contiguous(memory_format=fmt)
        g2 = g2.contiguous(memory_format=fmt)
        if g1.stride() != img1.stride() or g2.stride() != img1.stride():
            raise RuntimeError("compose_adversarial backward: gradient layout differs from the images'")
        L = _lib.lib()
        nbytes = L.b200corr_patch_compose_backward_scratch_bytes(n, p)
        scratch = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=img1.device)
        gp = torch.empty(ctx.patch_shape, dtype=torch.float32, device=img1.device)
        with torch.cuda.device(img1.device):
            code = L.b200corr_patch_compose_backward(
                _lib.ptr(img1), _lib.ptr(img2), _lib.ptr(patch), _lib.ptr(mask), _lib.ptr(pl), _lib.ptr(g1), _lib.ptr(g2),
                _lib.ptr(gp), _lib.ptr(scratch), nbytes, n, H, W, p, *_strides(img1), _lib.current_stream(img1.device))
        _lib.check(code, "b200corr_patch_compose_backward")
        return None, None, gp, None, None


def compose_adversarial(img1, img2, patch, mask, placements):
    """(adv1, adv2): the placed universal patch composited into both frames, clamped to [0, 1]."""
    return ComposeFunction.apply(img1, img2, patch, mask, placements)
