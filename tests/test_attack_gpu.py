"""Attack loop on the real operator: FlowNetC harness + patch attack on cuda:0."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_flownetc_harness_forward_backward_uses_the_cuda_sampler():
    from understanding_flow_robustness_b200 import _lib
    from understanding_flow_robustness_b200.harness import FlowNetCHarness
    torch.manual_seed(0)
    net = FlowNetCHarness().cuda().eval()
    a = torch.rand(2, 3, 128, 192, device="cuda", requires_grad=True)
    b = torch.rand(2, 3, 128, 192, device="cuda")
    n0 = _lib.lib().b200corr_launch_count()
    flow = net(a, b)
    assert flow.shape == (2, 2, 128, 192)
    flow.square().mean().backward()
    assert _lib.lib().b200corr_launch_count() - n0 >= 3          # fwd + 2 bwd kernels of the sampler
    assert torch.isfinite(a.grad).all() and float(a.grad.abs().max()) > 0


def test_sharded_patch_gradient_equals_whole_batch_gradient():
    """SURVEY 8e parity: sum of per-shard gradients == gradient of the whole batch (<= 1e-4 rel, fp32)."""
    from understanding_flow_robustness_b200 import attack
    from understanding_flow_robustness_b200.harness import FlowNetCHarness
    torch.manual_seed(1)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = FlowNetCHarness().cuda().eval()
    for q in net.parameters():
        q.requires_grad_(False)
    n, H, W, p = 4, 128, 192, 32
    i1 = torch.rand(n, 3, H, W, device="cuda")
    i2 = torch.rand(n, 3, H, W, device="cuda")
    patch = torch.rand(1, 3, p, p, device="cuda")
    mask = attack.circle_mask(p, "cuda")
    cfg = attack.PatchAttackConfig()
    pl = attack.sample_placements(n, H, W, p, cfg, None, "cuda")
    with torch.no_grad():
        tgt = -net(i1, i2)
    g_all, l_all = attack.patch_gradient(net, i1, i2, patch, mask, patch.clone(), pl, tgt, n, 0.0)
    g_sum, l_sum = torch.zeros_like(g_all), 0.0
    for r in range(2):
        idx = attack.shard_slice(n, r, 2)
        g, l = attack.patch_gradient(net, i1[idx], i2[idx], patch, mask, patch.clone(), pl[idx], tgt[idx], n, 0.0)
        g_sum += g
        l_sum += float(l)
    assert float((g_sum - g_all).abs().max()) <= 1e-4 * float(g_all.abs().max())
    assert abs(l_sum - float(l_all)) <= 1e-5 * abs(float(l_all))
    new, loss = attack.patch_attack_iteration(net, i1, i2, patch, mask, patch.clone(), cfg, n)
    assert new.shape == patch.shape and torch.isfinite(loss)
