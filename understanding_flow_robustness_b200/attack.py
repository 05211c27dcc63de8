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
    max_brightness: float = 0.05     # utils_patch.py:271-272: patch + U * 0.1 - 0.05, then clip


def shard_slice(global_batch, rank, world):
    """Indices of the image pairs owned by `rank` (pairs r::G)."""
    return list(range(rank, global_batch, world))


def circle_mask(p, device=None):
    """utils_patch.py:236-254: disc of radius p/2 - 2 inside a p x p square."""
    ys, xs = torch.meshgrid(torch.arange(p, device=device), torch.arange(p, device=device), indexing="ij")
    c = (p - 1) / 2.0
    return (((xs - c) ** 2 + (ys - c) ** 2) <= (p / 2.0 - 2) ** 2).float()[None, None]


def sample_placements(n, H, W, p, cfg, generator=None, device=None):
    """Per-pair (scale, angle, cx, cy, brightness offset); utils_patch.py:270-356 restated as parameters."""
    u = torch.rand(n, 5, generator=generator, device=device)
    scale = 1.0 + 2 * cfg.max_scale_jitter * (u[:, 0] - 0.5)
    angle = math.radians(2 * cfg.max_rotation_deg) * (u[:, 1] - 0.5)
    m = p / 2.0 + 2
    cx = m + u[:, 2] * (W - 2 * m)
    cy = m + u[:, 3] * (H - 2 * m)
    bright = 2 * cfg.max_brightness * (u[:, 4] - 0.5)
    return torch.stack([scale, angle, cx, cy, bright], dim=1)


def place(patch, mask, placements, H, W):
    """Differentiable paste of the canonical (1,3,p,p) patch and its mask into (n,3,H,W) canvases (torch ops:
    affine_grid + grid_sample).  placements: (n, 4) or (n, 5) -- the fifth column is the brightness offset."""
    n = placements.shape[0]
    p = patch.shape[-1]
    s, a, cx, cy = placements[:, 0], placements[:, 1], placements[:, 2], placements[:, 3]
    # output pixel (x, y) -> patch coordinate: R(-a) * (x - cx, y - cy) / s, normalised to [-1, 1]
    cos, sin = torch.cos(a) / s, torch.sin(a) / s
    sx, sy = (W - 1) / (p - 1), (H - 1) / (p - 1)
    tx = (-(cx - (W - 1) / 2.0) * cos - (cy - (H - 1) / 2.0) * sin) * 2 / (p - 1)
    ty = ((cx - (W - 1) / 2.0) * sin - (cy - (H - 1) / 2.0) * cos) * 2 / (p - 1)
    theta = torch.stack([torch.stack([cos * sx, sin * sy, tx], 1), torch.stack([-sin * sx, cos * sy, ty], 1)], 1)
    grid = F.affine_grid(theta, (n, 3, H, W), align_corners=True)
    q = patch.expand(n, -1, -1, -1)
    if placements.shape[1] > 4:
        q = (q + placements[:, 4].view(n, 1, 1, 1)).clamp(0, 1)      # utils_patch.py:271-273
    both = torch.cat([q * mask, mask.expand(n, -1, -1, -1)], 1)
    out = F.grid_sample(both, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    return out[:, :3], out[:, 3:4]


def compose_torch(img1, img2, patch, mask, placements):
    """main.py:537-542 on top of `place`: the torch-op formulation of `patch_transform.compose_adversarial`
    (any device / dtype; used by the CPU host-logic tests and as the kernel's parity target)."""
    canvas, m = place(patch, mask, placements, img1.shape[-2], img1.shape[-1])
    return ((1 - m) * img1 + canvas).clamp(0, 1), ((1 - m) * img2 + canvas).clamp(0, 1)


def compose_cuda(img1, img2, patch, mask, placements):
    """The same composition as one CUDA kernel (csrc/patch_transform.cu); raises on anything but CUDA fp32."""
    from .patch_transform import compose_adversarial

    return compose_adversarial(img1, img2, patch, mask, placements)


def cosine_flow_loss(flow, target):
    """patch_attacks/main.py:564-567: mean(1 - cos(flow_adv, target)) over pixels and pairs (sum form:
    the caller divides by the GLOBAL number of pairs so that shards add up)."""
    return (1.0 - F.cosine_similarity(flow, target, dim=1)).mean(dim=(1, 2)).sum()


def _allreduce_sum(t, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def patch_gradient(flow_fn, img1, img2, patch, mask, patch_init, placements, target, global_pairs, alpha,
                   compose_fn=compose_cuda):
    """Local contribution to d(loss)/d(patch) for this rank's pairs (loss normalised by global_pairs).
    Returns one packed tensor [gradient (3 p p), loss (1)] -- the buffer the all-reduce sums."""
    patch = patch.detach().requires_grad_(True)
    adv1, adv2 = compose_fn(img1, img2, patch, mask, placements)          # main.py:537-542
    flow = flow_fn(adv1, adv2)
    loss = (1 - alpha) * cosine_flow_loss(flow, target) / global_pairs
    if alpha > 0:
        loss = loss + alpha * (mask * (patch - patch_init)).abs().mean() * (img1.shape[0] / global_pairs)
    (g,) = torch.autograd.grad(loss, patch)
    return torch.cat([g.reshape(-1), loss.detach().reshape(1)])


def apply_patch_step(patch, packed, cfg):
    """main.py:575-600 on the all-reduced packed gradient: clamped step, clamp to the image range."""
    g = packed[:-1].view_as(patch)
    step = (0.5 * cfg.lr * g).clamp(-cfg.clamp, cfg.clamp)
    return (patch - step).clamp(0, 1), packed[-1]


def patch_attack_iteration(flow_fn, img1, img2, patch, mask, patch_init, cfg, global_pairs, generator=None,
                           group=None, compose_fn=compose_cuda):
    """One universal-patch iteration over this rank's shard; returns (new_patch, mean loss).

    Every rank must call it with the same `patch`; after the all-reduce every rank holds the same
    update, exactly as a single process iterating over the whole global batch would."""
    n, _, H, W = img1.shape
    with torch.no_grad():
        target = -flow_fn(img1, img2)                   # main.py:371,395
    placements = sample_placements(n, H, W, patch.shape[-1], cfg, generator, img1.device).to(img1.dtype)
    loss = None
    for _ in range(cfg.max_count):
        packed = patch_gradient(flow_fn, img1, img2, patch, mask, patch_init, placements, target, global_pairs,
                                cfg.alpha, compose_fn)
        _allreduce_sum(packed, group)                   # the one collective of the path: 3 p p + 1 floats
        patch, loss = apply_patch_step(patch, packed, cfg)
    return patch, loss


def perturbation_gradient(flow_fn, img1, img2, delta, target, global_pairs):
    """Local contribution to d(loss)/d(delta), delta (1, 2, 3, H, W); packed [gradient, loss]."""
    d = delta.detach().requires_grad_(True)
    flow = flow_fn((img1 + d[:, 0]).clamp(0, 1), (img2 + d[:, 1]).clamp(0, 1))      # :226-236 add + clamp
    loss = cosine_flow_loss(flow, target) / global_pairs
    (g,) = torch.autograd.grad(loss, d)
    return torch.cat([g.reshape(-1), loss.detach().reshape(1)])


def apply_perturbation_step(delta, packed, eps, step_size, sign=True):
    """universal_perturbation.py:477-520: (sign of the) gradient step, clamp to the +-eps ball."""
    g = packed[:-1].view_as(delta)
    upd = g.sign() if sign else g
    return (delta - step_size * upd).clamp(-eps, eps), packed[-1]


def universal_perturbation_iteration(flow_fn, img1, img2, delta, eps, step_size, n_step, global_pairs,
                                     sign=True, group=None):
    """universal_perturbation.py:452-530 over this rank's shard: delta is (1, 2, 3, H, W); the (2,3,H,W)
    gradient (3.9 MB at 256x640) is all-reduced before the sign step, every one of the n_step steps."""
    with torch.no_grad():
        target = -flow_fn(img1, img2)                   # :372-380
    loss = None
    for _ in range(n_step):
        packed = perturbation_gradient(flow_fn, img1, img2, delta, target, global_pairs)
        _allreduce_sum(packed, group)
        delta, loss = apply_perturbation_step(delta, packed, eps, step_size, sign)
    return delta, loss


class GraphedGradient:
    """A gradient function of the loops above captured once into a CUDA graph and replayed.

    `fn(*static_inputs) -> packed` must be a pure function of the tensors in `static_inputs` (they are updated in
    place before each replay).  The collective stays outside the graph: replay, then all-reduce the result.
    Per-rank batches of 8 pairs leave the GPU waiting on ~600 eager launches per step; a replay has none."""

    def __init__(self, fn, static_inputs, warmup=2):
        self.fn, self.inputs = fn, list(static_inputs)
        dev = self.inputs[0].device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):                     # cuDNN autotune, allocator, backward plans: before capture
                self.fn(*self.inputs)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self.fn(*self.inputs)

    def __call__(self, *new_inputs):
        for dst, src in zip(self.inputs, new_inputs):
            if src is not None and src is not dst:
                dst.copy_(src)
        self.graph.replay()
        return self.out


def nccl_value_check(flow_fn, device, rank, world, global_pairs, H, W, p, tol=None, group=None, dtype=torch.float32):
    """SURVEY section 4 item 6 as a function: every rank computes the patch gradient and the universal-perturbation
    gradient of its shard (pairs r::G of ONE seeded global batch) and all-reduces them (NCCL); rank 0 also computes
    both gradients of the whole batch in a single autograd call.  Returns the relative max-norm differences.
    Library convolutions run in fp32 (no TF32) for the duration of the check.

    dtype float64 (flow_fn in double; the sampler's fp64 kernels, torch-op placement) pins the plumbing to 1e-9;
    in float32 cuDNN picks other algorithms -- other summation orders -- for a shard than for the whole batch, and
    the gradients are sums of nearly cancelling terms: default tolerance 1e-2 there."""
    if tol is None:
        tol = 1e-9 if dtype == torch.float64 else 1e-2
    compose_fn = compose_cuda if dtype == torch.float32 else compose_torch
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        g = torch.Generator(device=device).manual_seed(1234)       # same stream of numbers on every rank
        i1 = torch.rand(global_pairs, 3, H, W, device=device, generator=g).to(dtype)
        i2 = torch.rand(global_pairs, 3, H, W, device=device, generator=g).to(dtype)
        patch = torch.rand(1, 3, p, p, device=device, generator=g).to(dtype)
        mask = circle_mask(p, device).to(dtype)
        cfg = PatchAttackConfig()
        pl = sample_placements(global_pairs, H, W, p, cfg, g, device).to(dtype)
        delta = (0.01 * torch.randn(1, 2, 3, H, W, device=device, generator=g)).to(dtype)
        idx = shard_slice(global_pairs, rank, world)
        with torch.no_grad():
            tgt = -flow_fn(i1[idx], i2[idx])
        gp = patch_gradient(flow_fn, i1[idx], i2[idx], patch, mask, patch, pl[idx], tgt, global_pairs, 0.0, compose_fn)
        gd = perturbation_gradient(flow_fn, i1[idx], i2[idx], delta, tgt, global_pairs)
        _allreduce_sum(gp, group)
        _allreduce_sum(gd, group)
        res = {"world": world, "global_pairs": global_pairs, "image": [H, W], "patch": p, "tol": tol,
               "dtype": str(dtype).replace("torch.", ""),
               "allreduce_bytes": {"patch": gp.numel() * 4, "perturbation": gd.numel() * 4}}
        if rank == 0:
            with torch.no_grad():
                tgt_all = -flow_fn(i1, i2)
            wp = patch_gradient(flow_fn, i1, i2, patch, mask, patch, pl, tgt_all, global_pairs, 0.0, compose_fn)
            wd = perturbation_gradient(flow_fn, i1, i2, delta, tgt_all, global_pairs)

            def rel(a, b):
                return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
            res.update(patch_grad_rel_err=rel(gp[:-1], wp[:-1]), perturbation_grad_rel_err=rel(gd[:-1], wd[:-1]),
                       loss_rel_err=max(abs(float(gp[-1]) / float(wp[-1]) - 1), abs(float(gd[-1]) / float(wd[-1]) - 1)))
            res["ok"] = bool(max(res["patch_grad_rel_err"], res["perturbation_grad_rel_err"], res["loss_rel_err"]) <= tol)
        return res
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
