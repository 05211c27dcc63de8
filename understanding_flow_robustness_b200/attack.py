"""Image-pair-sharded universal attacks: the batched restatement of the reference's attack loops.

Reference loops (both batch-1, single GPU, sequential):
  patch_attacks/main.py:345-613  -- universal adversarial patch: clean flow, random placement
      (utils_patch.py:257-358: brightness jitter, zoom, rotation, random location), up to max_count
      steps of  loss = (1-a) * mean(1 - cos(flow_adv, -flow_clean)) + a * L1(patch, patch_init),
      patch -= clamp(0.5 * lr * (dL/d adv_tgt + dL/d adv_ref), -2, 2), clamp to [0, 1].
  global_attacks/universal_perturbation.py:354-530 -- universal additive perturbation (2,3,H,W):
      n_step x { flow loss vs target, backward, sign (or raw) gradient step, clamp to +-eps }.

Sharding: every operator on the path is independent per image pair, so rank r of
G takes pairs r::G of the global batch; the ONLY cross-pair quantity is the gradient of the universal
patch / perturbation, summed with one all-reduce (NCCL over NVLink on the GPU box, 120 KB for a
100x100 patch) before the identical clamp-and-step on every rank.

Deliberate deviation, documented in DESIGN.md: the reference updates the patch after every single
pair and folds the placement jitter back into the stored patch; this module takes one mini-batch
step per iteration with differentiable placement (affine_grid / grid_sample), so the gradient lands
on the canonical patch.  Parity for the sharded loop is therefore "G-rank gradient == 1-rank gradient
of the same global batch", which tests/test_attack_cpu.py asserts with world_size 2 (gloo).
"""
import math
from dataclasses import dataclass

import torch
import torch.distributed as dist
import torch.nn.functional as F


@dataclass
class PatchAttackConfig:
    lr: float = 1e3            # patch_attacks/main.py --lr default region (step is clamped anyway)
    alpha: float = 0.0         # weight of the L1(patch, patch_init) regulariser (main.py:564-571)
    max_count: int = 2         # inner steps per iteration (main.py:546,610)
    clamp: float = 2.0         # main.py:581-583
    max_rotation_deg: float = 5.0    # utils_patch.py:290-296: 10 * (U - 0.5)
    max_scale_jitter: float = 0.025  # utils_patch.py:282-285: 1 + 0.05 * (U - 0.5)


def shard_slice(global_batch, rank, world):
    """Indices of the image pairs owned by `rank` (pairs r::G)."""
    return list(range(rank, global_batch, world))


def circle_mask(p, device=None):
    """utils_patch.py:236-254: disc of radius p/2 - 2 inside a p x p square."""
    ys, xs = torch.meshgrid(torch.arange(p, device=device), torch.arange(p, device=device), indexing="ij")
    c = (p - 1) / 2.0
    return (((xs - c) ** 2 + (ys - c) ** 2) <= (p / 2.0 - 2) ** 2).float()[None, None]


def sample_placements(n, H, W, p, cfg, generator=None, device=None):
    """Per-pair (scale, angle, cx, cy); utils_patch.py:270-356 restated as parameters."""
    u = torch.rand(n, 4, generator=generator, device=device)
    scale = 1.0 + 2 * cfg.max_scale_jitter * (u[:, 0] - 0.5)
    angle = math.radians(2 * cfg.max_rotation_deg) * (u[:, 1] - 0.5)
    m = p / 2.0 + 2
    cx = m + u[:, 2] * (W - 2 * m)
    cy = m + u[:, 3] * (H - 2 * m)
    return torch.stack([scale, angle, cx, cy], dim=1)


def place(patch, mask, placements, H, W):
    """Differentiable paste of the canonical (1,3,p,p) patch and its mask into (n,3,H,W) canvases."""
    n = placements.shape[0]
    p = patch.shape[-1]
    s, a, cx, cy = placements.unbind(1)
    # output pixel (x, y) -> patch coordinate: R(-a) * (x - cx, y - cy) / s, normalised to [-1, 1]
    cos, sin = torch.cos(a) / s, torch.sin(a) / s
    sx, sy = (W - 1) / (p - 1), (H - 1) / (p - 1)
    tx = (-(cx - (W - 1) / 2.0) * cos - (cy - (H - 1) / 2.0) * sin) * 2 / (p - 1)
    ty = ((cx - (W - 1) / 2.0) * sin - (cy - (H - 1) / 2.0) * cos) * 2 / (p - 1)
    theta = torch.stack([torch.stack([cos * sx, sin * sy, tx], 1), torch.stack([-sin * sx, cos * sy, ty], 1)], 1)
    grid = F.affine_grid(theta, (n, 3, H, W), align_corners=True)
    both = torch.cat([patch * mask, mask], 1).expand(n, -1, -1, -1)
    out = F.grid_sample(both, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    return out[:, :3], out[:, 3:4]


def cosine_flow_loss(flow, target):
    """patch_attacks/main.py:564-567: mean(1 - cos(flow_adv, target)) over pixels and pairs (sum form:
    the caller divides by the GLOBAL number of pairs so that shards add up)."""
    return (1.0 - F.cosine_similarity(flow, target, dim=1)).mean(dim=(1, 2)).sum()


def _allreduce_sum(t, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def patch_gradient(flow_fn, img1, img2, patch, mask, patch_init, placements, target, global_pairs, alpha):
    """Local contribution to d(loss)/d(patch) for this rank's pairs (loss normalised by global_pairs)."""
    patch = patch.detach().requires_grad_(True)
    H, W = img1.shape[-2:]
    canvas, m = place(patch, mask, placements, H, W)
    adv1 = (1 - m) * img1 + canvas                      # main.py:537-542 (canvas already carries the mask)
    adv2 = (1 - m) * img2 + canvas
    flow = flow_fn(adv1.clamp(0, 1), adv2.clamp(0, 1))
    loss = (1 - alpha) * cosine_flow_loss(flow, target) / global_pairs
    if alpha > 0:
        loss = loss + alpha * (mask * (patch - patch_init)).abs().mean() * (img1.shape[0] / global_pairs)
    (g,) = torch.autograd.grad(loss, patch)
    return g, loss.detach()


def patch_attack_iteration(flow_fn, img1, img2, patch, mask, patch_init, cfg, global_pairs, generator=None,
                           group=None):
    """One universal-patch iteration over this rank's shard; returns (new_patch, mean loss).

    Every rank must call it with the same `patch`; after the all-reduce every rank holds the same
    update, exactly as a single process iterating over the whole global batch would."""
    n, _, H, W = img1.shape
    with torch.no_grad():
        target = -flow_fn(img1, img2)                   # main.py:371,395
    placements = sample_placements(n, H, W, patch.shape[-1], cfg, generator, img1.device)
    loss = None
    for _ in range(cfg.max_count):
        g, loss = patch_gradient(flow_fn, img1, img2, patch, mask, patch_init, placements, target, global_pairs,
                                 cfg.alpha)
        packed = torch.cat([g.reshape(-1), loss.reshape(1)])
        _allreduce_sum(packed, group)                   # the one collective of the path
        g, loss = packed[:-1].view_as(patch), packed[-1]
        step = (0.5 * cfg.lr * g).clamp(-cfg.clamp, cfg.clamp)   # main.py:575-583
        patch = (patch - step).clamp(0, 1)              # main.py:585-600
    return patch, loss


def universal_perturbation_iteration(flow_fn, img1, img2, delta, eps, step_size, n_step, global_pairs,
                                     sign=True, group=None):
    """universal_perturbation.py:452-530 over this rank's shard: delta is (1, 2, 3, H, W)."""
    with torch.no_grad():
        target = -flow_fn(img1, img2)                   # :372-380
    loss = None
    for _ in range(n_step):
        d = delta.detach().requires_grad_(True)
        flow = flow_fn((img1 + d[:, 0]).clamp(0, 1), (img2 + d[:, 1]).clamp(0, 1))
        loss = cosine_flow_loss(flow, target) / global_pairs
        (g,) = torch.autograd.grad(loss, d)
        packed = torch.cat([g.reshape(-1), loss.detach().reshape(1)])
        _allreduce_sum(packed, group)
        g, loss = packed[:-1].view_as(delta), packed[-1]
        upd = g.sign() if sign else g                   # :477-488
        delta = (delta - step_size * upd).clamp(-eps, eps)   # :503-520
    return delta, loss
