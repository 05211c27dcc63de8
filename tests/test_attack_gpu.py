"""Attack loops on the real operators (cuda:0): FlowNetC body + patch attack + universal perturbation,
the GPU patch placement kernel against its torch formulation, CUDA-graph replay of the gradient step, and
(where the box has >= 2 GPUs) the NCCL value check: sharded all-reduced gradient == single-process gradient."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def _fp32_library_math():
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


def _net(fused=False):
    from understanding_flow_robustness_b200.harness import FlowNetCHarness
    net = FlowNetCHarness(fused_merge=fused).cuda().eval()
    for q in net.parameters():
        q.requires_grad_(False)
    return net


def test_flownetc_harness_forward_backward_uses_the_cuda_sampler():
    from understanding_flow_robustness_b200 import _lib
    torch.manual_seed(0)
    net = _net()
    a = torch.rand(2, 3, 128, 192, device="cuda", requires_grad=True)
    b = torch.rand(2, 3, 128, 192, device="cuda")
    n0 = _lib.lib().b200corr_launch_count()
    flow = net(a, b)
    assert flow.shape == (2, 2, 128, 192)
    flow.square().mean().backward()
    assert _lib.lib().b200corr_launch_count() - n0 >= 3          # fwd + 2 bwd kernels of the sampler
    assert torch.isfinite(a.grad).all() and float(a.grad.abs().max()) > 0


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("n,H,W,p", [(3, 64, 96, 20), (2, 128, 192, 33), (1, 40, 40, 2)])
def test_patch_compose_kernel_equals_the_torch_formulation(n, H, W, p, channels_last):
    """csrc/patch_transform.cu against affine_grid + grid_sample (attack.compose_torch): values <= 1e-4
    (coordinate round trip through the normalised grid), gradient w.r.t. the patch <= 1e-4 relative."""
    from understanding_flow_robustness_b200 import attack
    g = torch.Generator(device="cuda").manual_seed(n * 1000 + p)
    i1 = torch.rand(n, 3, H, W, device="cuda", generator=g)
    i2 = torch.rand(n, 3, H, W, device="cuda", generator=g)
    if channels_last:
        i1, i2 = (t.contiguous(memory_format=torch.channels_last) for t in (i1, i2))
    patch = torch.rand(1, 3, p, p, device="cuda", generator=g)
    patch[0, :, 0, 0] = torch.tensor([0.0, 1.0, 0.98], device="cuda")       # the brightness clip is exercised
    mask = attack.circle_mask(p, "cuda") if p > 8 else torch.ones(1, 1, p, p, device="cuda")
    cfg = attack.PatchAttackConfig(max_rotation_deg=20.0, max_scale_jitter=0.2)
    pl = attack.sample_placements(n, H, W, p, cfg, g, "cuda")
    pl[0, 2:4] = torch.tensor([1.5, H - 2.25], device="cuda")              # mostly outside the image
    w1 = torch.randn(n, 3, H, W, device="cuda", generator=g)
    w2 = torch.randn(n, 3, H, W, device="cuda", generator=g)
    res = []
    for fn in (attack.compose_cuda, attack.compose_torch):
        pt = patch.clone().requires_grad_(True)
        a1, a2 = fn(i1, i2, pt, mask, pl)
        (gp,) = torch.autograd.grad((a1 * w1).sum() + (a2 * w2).sum(), pt)
        res.append((a1.detach(), a2.detach(), gp))
    (k1, k2, kg), (t1, t2, tg) = res
    assert k1.stride() == i1.stride()
    assert float((k1 - t1).abs().max()) <= 1e-4 and float((k2 - t2).abs().max()) <= 1e-4
    assert float(k1.min()) >= 0.0 and float(k1.max()) <= 1.0
    # pixels whose pre-clamp value sits within rounding of 0 or 1 may flip the clamp predicate: compare in aggregate
    assert float((kg - tg).abs().max()) <= 1e-3 * float(tg.abs().max()) + 1e-4
    # deterministic: the gather backward has no atomics
    pt = patch.clone().requires_grad_(True)
    a1, a2 = attack.compose_cuda(i1, i2, pt, mask, pl)
    (gp2,) = torch.autograd.grad((a1 * w1).sum() + (a2 * w2).sum(), pt)
    assert torch.equal(gp2, kg)


def test_sharded_patch_gradient_equals_whole_batch_gradient():
    """SURVEY 8e parity: sum of per-shard gradients == gradient of the whole batch (<= 1e-4 rel, fp32)."""
    from understanding_flow_robustness_b200 import attack
    torch.manual_seed(1)
    net = _net()
    n, H, W, p = 4, 128, 192, 32
    i1 = torch.rand(n, 3, H, W, device="cuda")
    i2 = torch.rand(n, 3, H, W, device="cuda")
    patch = torch.rand(1, 3, p, p, device="cuda")
    mask = attack.circle_mask(p, "cuda")
    cfg = attack.PatchAttackConfig()
    pl = attack.sample_placements(n, H, W, p, cfg, None, "cuda")
    with torch.no_grad():
        tgt = -net(i1, i2)
    whole = attack.patch_gradient(net, i1, i2, patch, mask, patch.clone(), pl, tgt, n, 0.0)
    parts = torch.zeros_like(whole)
    for r in range(2):
        idx = attack.shard_slice(n, r, 2)
        parts += attack.patch_gradient(net, i1[idx], i2[idx], patch, mask, patch.clone(), pl[idx], tgt[idx], n, 0.0)
    assert float((parts[:-1] - whole[:-1]).abs().max()) <= 1e-4 * float(whole[:-1].abs().max())
    assert abs(float(parts[-1]) - float(whole[-1])) <= 1e-5 * abs(float(whole[-1]))
    # the kernel composition and the torch-op composition drive the same gradient through the network
    ref = attack.patch_gradient(net, i1, i2, patch, mask, patch.clone(), pl, tgt, n, 0.0, attack.compose_torch)
    assert float((ref[:-1] - whole[:-1]).abs().max()) <= 2e-3 * float(ref[:-1].abs().max())
    new, loss = attack.patch_attack_iteration(net, i1, i2, patch, mask, patch.clone(), cfg, n)
    assert new.shape == patch.shape and torch.isfinite(loss)
    assert float(new.min()) >= 0.0 and float(new.max()) <= 1.0


def test_universal_perturbation_at_256x640_sharded_equals_whole_batch():
    """global_attacks/universal_perturbation.py:452-530 on the GPU at the reference's training resolution:
    the (2,3,256,640) gradient (3.9 MB, the all-reduce payload) of two shards sums to the whole batch's,
    and the iteration keeps delta inside the eps ball."""
    from understanding_flow_robustness_b200 import attack
    torch.manual_seed(2)
    net = _net(fused=True)
    n, H, W = 4, 256, 640
    i1 = torch.rand(n, 3, H, W, device="cuda")
    i2 = torch.rand(n, 3, H, W, device="cuda")
    delta = 0.01 * torch.randn(1, 2, 3, H, W, device="cuda")
    with torch.no_grad():
        tgt = -net(i1, i2)
    whole = attack.perturbation_gradient(net, i1, i2, delta, tgt, n)
    assert whole.numel() == 2 * 3 * H * W + 1 and whole.numel() * 4 == 3932164
    parts = torch.zeros_like(whole)
    for r in range(2):
        idx = attack.shard_slice(n, r, 2)
        parts += attack.perturbation_gradient(net, i1[idx], i2[idx], delta, tgt[idx], n)
    # fp32: cuDNN picks other algorithms (summation orders) for batch 2 than for batch 4 and the cosine-loss gradient
    # of a random-init net is a sum of nearly cancelling terms -> 1e-2 here; the fp64 run below pins the arithmetic
    assert float((parts[:-1] - whole[:-1]).abs().max()) <= 1e-2 * float(whole[:-1].abs().max())
    assert abs(float(parts[-1]) - float(whole[-1])) <= 1e-5 * abs(float(whole[-1]))
    net64 = _net().double()
    a, b, d64 = i1[:, :, :64, :128].double(), i2[:, :, :64, :128].double(), delta[..., :64, :128].double()
    with torch.no_grad():
        t64 = -net64(a, b)
    whole64 = attack.perturbation_gradient(net64, a, b, d64, t64, n)
    parts64 = sum(attack.perturbation_gradient(net64, a[i], b[i], d64, t64[i], n)
                  for i in (attack.shard_slice(n, r, 2) for r in range(2)))
    assert float((parts64 - whole64).abs().max()) <= 1e-9 * float(whole64.abs().max())
    eps, lr = 0.02, 0.005
    new, loss = attack.universal_perturbation_iteration(net, i1, i2, torch.zeros_like(delta), eps, lr, 3, n)
    assert new.shape == delta.shape and torch.isfinite(loss)
    assert float(new.abs().max()) <= eps + 1e-7 and float(new.abs().max()) > 0
    # three signed steps of 0.005 from zero: every entry is a multiple of the step inside the ball
    assert float(((new / lr).round() * lr - new).abs().max()) <= 1e-6


def test_graph_replay_of_the_gradient_step_equals_eager():
    """attack.GraphedGradient: the captured step (compose kernel, conv stack, sampler fwd+bwd, loss, autograd)
    replayed with new patch / placements gives the eager result."""
    from understanding_flow_robustness_b200 import attack
    torch.manual_seed(3)
    net = _net(fused=True)
    n, H, W, p = 2, 128, 192, 24
    i1 = torch.rand(n, 3, H, W, device="cuda")
    i2 = torch.rand(n, 3, H, W, device="cuda")
    mask = attack.circle_mask(p, "cuda")
    cfg = attack.PatchAttackConfig()
    with torch.no_grad():
        tgt = -net(i1, i2)
    init = torch.rand(1, 3, p, p, device="cuda")

    def grad_fn(patch, pl):
        return attack.patch_gradient(net, i1, i2, patch, mask, init, pl, tgt, n, 0.0)

    s_patch = init.clone()
    s_pl = attack.sample_placements(n, H, W, p, cfg, None, "cuda")
    graphed = attack.GraphedGradient(grad_fn, [s_patch, s_pl])
    for k in range(3):
        patch = torch.rand(1, 3, p, p, device="cuda")
        pl = attack.sample_placements(n, H, W, p, cfg, None, "cuda")
        got = graphed(patch, pl).clone()
        want = grad_fn(patch, pl)
        assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max()), k


@pytest.mark.parametrize("world", [2])
def test_nccl_sharded_gradients_equal_single_process(world):
    """`torchrun --nproc N` -- the all-reduced patch and perturbation gradients of N
    ranks equal the single-process gradients of the same global batch (<= 1e-4 relative)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29631",
                        os.path.join(ROOT, "tests", "nccl_value_check.py")],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "nccl value check ok" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
