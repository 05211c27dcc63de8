"""Host logic of the pair-sharded attack loops on CPU: world_size-2 gloo run == single process.

The flow network here is a tiny torch-only stand-in (the correlation operator has no CPU path); the
attack code is agnostic to `flow_fn`.  The GPU variant with the FlowNetC harness is in
tests/test_attack_gpu.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from understanding_flow_robustness_b200 import attack


class TinyFlow(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = torch.nn.Conv2d(6, 8, 3, 1, 1)
        self.b = torch.nn.Conv2d(8, 2, 3, 1, 1)

    def forward(self, x1, x2):
        return self.b(torch.tanh(self.a(torch.cat([x1, x2], 1))))


def _data(n=6, H=24, W=32):
    g = torch.Generator().manual_seed(1)
    return torch.rand(n, 3, H, W, generator=g), torch.rand(n, 3, H, W, generator=g)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _single(kind):
    net = TinyFlow().double()
    i1, i2 = (t.double() for t in _data())
    cfg = attack.PatchAttackConfig(lr=50.0, max_count=2)
    if kind == "patch":
        p = 10
        patch = torch.rand(1, 3, p, p, generator=torch.Generator().manual_seed(2)).double()
        mask = attack.circle_mask(p).double()
        pl = attack.sample_placements(6, 24, 32, p, cfg, torch.Generator().manual_seed(3)).double()
        with torch.no_grad():
            tgt = -net(i1, i2)
        packed = attack.patch_gradient(net, i1, i2, patch, mask, patch.clone(), pl, tgt, 6, 0.1, attack.compose_torch)
        return packed[:-1].view_as(patch), packed[-1]
    delta = torch.zeros(1, 2, 3, 24, 32).double()
    return attack.universal_perturbation_iteration(net, i1, i2, delta, 0.05, 0.01, 3, 6)


def _worker(rank, world, port, kind, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    net = TinyFlow().double()
    i1, i2 = (t.double() for t in _data())
    idx = attack.shard_slice(6, rank, world)
    cfg = attack.PatchAttackConfig(lr=50.0, max_count=2)
    if kind == "patch":
        p = 10
        patch = torch.rand(1, 3, p, p, generator=torch.Generator().manual_seed(2)).double()
        mask = attack.circle_mask(p).double()
        pl = attack.sample_placements(6, 24, 32, p, cfg, torch.Generator().manual_seed(3)).double()[idx]
        with torch.no_grad():
            tgt = -net(i1[idx], i2[idx])
        packed = attack.patch_gradient(net, i1[idx], i2[idx], patch, mask, patch.clone(), pl, tgt, 6, 0.1,
                                       attack.compose_torch)
        dist.all_reduce(packed)
        out = (packed[:-1].view_as(patch), packed[-1])
    else:
        delta = torch.zeros(1, 2, 3, 24, 32).double()
        out = attack.universal_perturbation_iteration(net, i1[idx], i2[idx], delta, 0.05, 0.01, 3, 6)
    if rank == 0:
        q.put((out[0].numpy(), float(out[1])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["patch", "perturbation"])
def test_sharded_equals_single_process(kind):
    ref_g, ref_loss = _single(kind)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, q)) for r in range(2)]
    for p in procs:
        p.start()
    g, loss = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.from_numpy(g)
    assert float((g - ref_g).abs().max()) <= 1e-9 * max(1.0, float(ref_g.abs().max()))
    assert abs(loss - float(ref_loss)) <= 1e-9


def test_shard_slice_partitions_the_batch():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in attack.shard_slice(64, r, world))
        assert seen == list(range(64))


def test_placement_is_identity_for_centered_unit_patch():
    p, H, W = 9, 9, 9
    patch = torch.rand(1, 3, p, p)
    mask = torch.ones(1, 1, p, p)
    pl = torch.tensor([[1.0, 0.0, (W - 1) / 2.0, (H - 1) / 2.0]])
    canvas, m = attack.place(patch, mask, pl, H, W)
    assert torch.allclose(canvas, patch, atol=1e-5) and torch.allclose(m, mask, atol=1e-5)
    # translated by +2 px in x on a larger canvas
    canvas, m = attack.place(patch, mask, torch.tensor([[1.0, 0.0, 10.0, 8.0]]), 17, 21)
    assert torch.allclose(canvas[0, :, 4:13, 6:15], patch[0], atol=1e-5)
    assert float(m.sum()) == pytest.approx(81.0, abs=1e-3)


def test_brightness_column_shifts_and_clips_the_patch():
    p, H, W = 9, 9, 9
    patch = torch.rand(1, 3, p, p)
    mask = torch.ones(1, 1, p, p)
    pl = torch.tensor([[1.0, 0.0, (W - 1) / 2.0, (H - 1) / 2.0, 0.05], [1.0, 0.0, 4.0, 4.0, -2.0]])
    canvas, _ = attack.place(patch, mask, pl, H, W)
    assert torch.allclose(canvas[0], (patch[0] + 0.05).clamp(0, 1), atol=1e-5)
    assert float(canvas[1].abs().max()) <= 1e-6                 # clipped at zero (utils_patch.py:272)


def test_the_default_composition_refuses_cpu_tensors():
    """No silent CPU path: the loops default to the CUDA kernel and say so when handed host tensors."""
    net = TinyFlow()
    i1, i2 = _data(2)
    patch = torch.rand(1, 3, 10, 10)
    cfg = attack.PatchAttackConfig()
    with pytest.raises(RuntimeError):
        attack.patch_attack_iteration(net, i1, i2, patch, attack.circle_mask(10), patch.clone(), cfg, 2)


def test_patch_iteration_runs_and_keeps_range():
    net = TinyFlow()
    i1, i2 = _data(4)
    p = 10
    patch = torch.rand(1, 3, p, p)
    cfg = attack.PatchAttackConfig(lr=100.0, max_count=2)
    new, loss = attack.patch_attack_iteration(net, i1, i2, patch, attack.circle_mask(p), patch.clone(), cfg, 4,
                                              torch.Generator().manual_seed(0), compose_fn=attack.compose_torch)
    assert new.shape == patch.shape and float(new.min()) >= 0.0 and float(new.max()) <= 1.0
    assert torch.isfinite(loss)
