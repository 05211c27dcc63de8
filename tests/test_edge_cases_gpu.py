"""Edge cases the reference handles implicitly (empty batch, one-pixel maps, a single channel, patches
larger than the image, ragged sizes) through every entry point; checked against the CPU oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def test_empty_batch_everywhere():
    from understanding_flow_robustness_b200 import CorrBlock, AlternateCorrBlock, coords_grid, raft_corr
    from understanding_flow_robustness_b200 import spatial_correlation_sample
    a = torch.zeros(0, 8, 6, 8, device="cuda", requires_grad=True)
    b = torch.zeros(0, 8, 6, 8, device="cuda", requires_grad=True)
    out = spatial_correlation_sample(a, b, kernel_size=1, patch_size=21, dilation_patch=2)
    assert tuple(out.shape) == (0, 21, 21, 6, 8)
    out.sum().backward()
    assert tuple(a.grad.shape) == (0, 8, 6, 8)
    f = torch.zeros(0, 8, 8, 8, device="cuda")
    pyr = raft_corr.allpairs_pyramid(f, f, 3, "tf32")
    assert [tuple(p.shape) for p in pyr] == [(0, 1, 8, 8), (0, 1, 4, 4), (0, 1, 2, 2)]
    c = torch.zeros(0, 2, 8, 8, device="cuda")
    assert tuple(raft_corr.lookup_forward(pyr, c, 2, 8, 8).shape) == (0, 3 * 25, 8, 8)
    assert tuple(CorrBlock(f, f, 3, 2)(c).shape) == (0, 75, 8, 8)
    assert tuple(AlternateCorrBlock(f, f, 3, 2)(c).shape) == (0, 75, 8, 8)


@pytest.mark.parametrize("shape", [(1, 1, 1, 4), (2, 3, 1, 8), (1, 5, 3, 4), (1, 129, 2, 12), (3, 2, 7, 16)],
                         ids=lambda s: "x".join(map(str, s)))
def test_sampler_tiny_and_oversized_patch(shape):
    """FlowNetC structure on maps much smaller than the 41-pixel displacement range (most of the
    21x21 planes are entirely out of bounds), one channel, one row; register-blocked and generic path."""
    from oracle import sampler_oracle
    from understanding_flow_robustness_b200 import spatial_correlation_sample
    rng = np.random.default_rng(sum(shape))
    in1 = rng.standard_normal(shape).astype(np.float32)
    in2 = rng.standard_normal(shape).astype(np.float32)
    a = torch.from_numpy(in1).cuda().requires_grad_()
    b = torch.from_numpy(in2).cuda().requires_grad_()
    out = spatial_correlation_sample(a, b, kernel_size=1, patch_size=21, dilation_patch=2)
    g = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    out.backward(torch.from_numpy(g).cuda())
    ref = sampler_oracle.forward(in1, in2, 1, 21, 1, 0, 1, 2)
    r1, r2 = sampler_oracle.backward(in1, in2, g, 1, 21, 1, 0, 1, 2)
    assert _rel(out.detach().cpu().numpy(), ref) <= 1e-5
    assert _rel(a.grad.cpu().numpy(), r1) <= 1e-5 and _rel(b.grad.cpu().numpy(), r2) <= 1e-5


@pytest.mark.parametrize("hw", [(1, 4), (2, 8), (3, 5), (9, 4)], ids=lambda s: "x".join(map(str, s)))
def test_raft_tiny_maps(hw):
    """All-pairs + pyramid + lookup on maps smaller than one tile / one lookup window."""
    import math
    from oracle import raft_oracle
    from understanding_flow_robustness_b200 import coords_grid, raft_corr
    H, W = hw
    torch.manual_seed(H * 10 + W)
    B, C, L, r = 2, 5, 1, 4
    f1 = torch.randn(B, C, H, W, device="cuda")
    f2 = torch.randn(B, C, H, W, device="cuda")
    exact = torch.einsum("bcm,bcn->bmn", f1.double().view(B, C, -1), f2.double().view(B, C, -1)) / math.sqrt(C)
    for prec, tol in (("fp32", 1e-5), ("tf32", 2e-3)):
        vol = raft_corr.allpairs_pyramid(f1, f2, L, prec)[0].view(B, H * W, H * W).double()
        assert float((vol - exact).abs().max()) <= tol * float(exact.abs().max()) + 1e-6
    pyr = raft_corr.allpairs_pyramid(f1, f2, L, "fp32")
    coords = coords_grid(B, H, W, "cuda") + 1.5 * torch.randn(B, 2, H, W, device="cuda")
    out = raft_corr.lookup_forward(pyr, coords, r, H, W, "direct").cpu().numpy()
    ora = raft_oracle.lookup([p.cpu().numpy() for p in pyr], coords.cpu().numpy(), r, roundtrip=False)
    assert np.abs(out - ora).max() <= 2e-6 * np.abs(ora).max() + 1e-7


def test_cuda_graph_capture_of_the_operators():
    """Every operator launches on the current stream with caller-owned memory only, so a step can be
    captured into a CUDA graph and replayed on new data (what bench.py times)."""
    from understanding_flow_robustness_b200 import backend, coords_grid, raft_corr
    torch.manual_seed(0)
    q = (1, 1, 21, 21, 0, 0, 1, 1, 2, 2, 1, 1)
    a = torch.randn(2, 128, 12, 32, device="cuda")
    b = torch.randn(2, 128, 12, 32, device="cuda")
    g = torch.randn(2, 21, 21, 12, 32, device="cuda")
    f1 = torch.randn(1, 64, 16, 32, device="cuda")
    f2 = torch.randn(1, 64, 16, 32, device="cuda")
    c = coords_grid(1, 16, 32, "cuda") + 2.0 * torch.randn(1, 2, 16, 32, device="cuda")
    with torch.no_grad():
        backend.forward(a, b, *q); backend.backward(a, b, g, *q)          # warm-up: attributes, plan cache
        raft_corr.lookup_forward(raft_corr.allpairs_pyramid(f1, f2, 3, "tf32"), c, 3, 16, 32)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = backend.forward(a, b, *q)
            g1, g2 = backend.backward(a, b, g, *q)
            pyr = raft_corr.allpairs_pyramid(f1, f2, 3, "tf32")
            look = raft_corr.lookup_forward(pyr, c, 3, 16, 32)
        for t in (a, b, g, f1, f2):
            t.normal_()                                                   # new data in the captured buffers
        graph.replay()
        torch.cuda.synchronize()
        got = [t.clone() for t in (out, g1, g2, look)]
        ref_out = backend.forward(a, b, *q)
        r1, r2 = backend.backward(a, b, g, *q)
        ref_look = raft_corr.lookup_forward(raft_corr.allpairs_pyramid(f1, f2, 3, "tf32"), c, 3, 16, 32)
    for x, y in zip(got, (ref_out, r1, r2, ref_look)):
        assert torch.equal(x, y)


def test_caller_owned_output_buffers_and_host_pipeline():
    """backend.forward/backward(out=...) write into caller-owned tensors (what the C ABI does anyway);
    the host-buffer pipeline built on it reproduces the plain operator."""
    from understanding_flow_robustness_b200 import backend
    from understanding_flow_robustness_b200.host_pipeline import SamplerHostPipeline
    torch.manual_seed(1)
    q = (1, 1, 21, 21, 0, 0, 1, 1, 2, 2, 1, 1)
    shape = (2, 32, 8, 16)
    a, b = torch.randn(*shape, device="cuda"), torch.randn(*shape, device="cuda")
    g = torch.randn(2, 21, 21, 8, 16, device="cuda")
    ref_out = backend.forward(a, b, *q)
    r1, r2 = backend.backward(a, b, g, *q)
    out, g1, g2 = torch.empty_like(ref_out), torch.empty_like(a), torch.empty_like(b)
    assert backend.forward(a, b, *q, out=out) is out
    backend.backward(a, b, g, *q, out=(g1, g2))
    assert torch.equal(out, ref_out) and torch.equal(g1, r1) and torch.equal(g2, r2)
    with pytest.raises(RuntimeError):
        backend.forward(a, b, *q, out=torch.empty(2, 21, 21, 8, 8, device="cuda"))
    with pytest.raises(RuntimeError):
        backend.backward(a, b, g, *q, out=(g1, torch.empty(1, device="cuda")))
    pipe = SamplerHostPipeline(shape, q, torch.device("cuda"))
    hosts = []
    for i in range(5):   # more batches than slots
        h = [torch.randn(*shape).pin_memory(), torch.randn(*shape).pin_memory(), torch.randn(2, 21, 21, 8, 16).pin_memory(),
             torch.empty(2, 21, 21, 8, 16).pin_memory(), torch.empty(*shape).pin_memory(), torch.empty(*shape).pin_memory()]
        pipe.submit(*h)
        hosts.append(h)
    pipe.synchronize()
    for h in hosts:
        o = backend.forward(h[0].cuda(), h[1].cuda(), *q)
        x1, x2 = backend.backward(h[0].cuda(), h[1].cuda(), h[2].cuda(), *q)
        assert torch.equal(h[3].cuda(), o) and torch.equal(h[4].cuda(), x1) and torch.equal(h[5].cuda(), x2)


def test_round2_entry_points_empty_batch_errors_and_graph_capture():
    """The round-2 entry points (volume backward, lookup -> convc1 in both variants, 1x1 conv, patch placement):
    empty batch, loud failure on CPU tensors, CUDA-graph capture + replay gives the eager values."""
    import math

    import torch.nn.functional as F
    from understanding_flow_robustness_b200 import CorrBlock, attack, coords_grid, raft_corr
    # ---- empty batch
    f0 = torch.zeros(0, 16, 8, 8, device="cuda")
    g0 = [torch.zeros(0, 1, 8 >> l, 8 >> l, device="cuda") for l in range(3)]
    d1, d2 = raft_corr.volume_backward(g0, f0, f0, 0.25, "tf32")
    assert tuple(d1.shape) == tuple(d2.shape) == (0, 16, 8, 8)
    w = torch.randn(32, 3 * 25, device="cuda")
    c0 = torch.zeros(0, 2, 8, 8, device="cuda")
    assert tuple(CorrBlock(f0, f0, 3, 2).lookup_convc1(c0, w, None).shape) == (0, 32, 8, 8)
    assert tuple(raft_corr.conv1x1_forward(torch.zeros(0, 8, 4, 4, device="cuda"), torch.randn(32, 8, device="cuda")).shape) == (0, 32, 4, 4)
    i0 = torch.zeros(0, 3, 16, 16, device="cuda")
    pt = torch.rand(1, 3, 8, 8, device="cuda", requires_grad=True)
    a1, a2 = attack.compose_cuda(i0, i0, pt, torch.ones(1, 1, 8, 8, device="cuda"), torch.zeros(0, 5, device="cuda"))
    (a1.sum() + a2.sum()).backward()
    assert tuple(a1.shape) == (0, 3, 16, 16) and float(pt.grad.abs().max()) == 0.0
    # ---- CPU tensors are refused (no fallback)
    with pytest.raises(RuntimeError):
        raft_corr.conv1x1_forward(torch.zeros(1, 8, 4, 4), torch.zeros(32, 8))
    with pytest.raises(RuntimeError):
        raft_corr.volume_backward([torch.zeros(64, 1, 8, 8)], torch.zeros(1, 16, 8, 8), torch.zeros(1, 16, 8, 8), 0.25, "fp32")
    with pytest.raises(RuntimeError):
        attack.compose_cuda(torch.zeros(1, 3, 16, 16), torch.zeros(1, 3, 16, 16), torch.zeros(1, 3, 8, 8), torch.ones(1, 1, 8, 8),
                            torch.zeros(1, 5))
    with pytest.raises(RuntimeError):        # unsupported channel count of the fused kernel's C ABI: loud, not silent
        raft_corr.lookup_convc1_forward([torch.zeros(64, 1, 8, 8, device="cuda")], torch.zeros(1, 2, 8, 8, device="cuda"),
                                        torch.zeros(1 * 24 * 32, device="cuda"), None, 24, 2, 8, 8)
    # ---- graph capture
    torch.manual_seed(0)
    B, C, H, W = 1, 32, 16, 32
    f1 = torch.randn(B, C, H, W, device="cuda")
    f2 = torch.randn(B, C, H, W, device="cuda")
    conv = torch.nn.Conv2d(4 * 81, 128, 1).cuda()
    glv = [torch.randn(B * H * W, 1, H >> l, W >> l, device="cuda") for l in range(4)]
    coords = coords_grid(B, H, W, "cuda") + torch.randn(B, 2, H, W, device="cuda")
    with torch.no_grad():
        blk = CorrBlock(f1, f2, 4, 4, precision="tf32")

        def work():
            return (blk.lookup_convc1(coords, conv.weight, conv.bias, impl="fused"),
                    blk.lookup_convc1(coords, conv.weight, conv.bias, impl="pipelined"),
                    *raft_corr.volume_backward(glv, f1, f2, 1 / math.sqrt(C), "fp32"))
        eager = [t.clone() for t in work()]
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            outs = work()
        coords.add_(0.0)
        g.replay()
        torch.cuda.synchronize()
        for a, b in zip(outs, eager):
            assert torch.equal(a, b)
        # fused vs pipelined differ only by rounding vs truncation of the operands
        assert float((eager[0] - eager[1]).abs().max()) <= 2.0 ** -8 * float(F.conv2d(blk(coords).abs(), conv.weight.abs()).max())


@pytest.mark.parametrize("B,K,N,H,W", [(1, 324, 256, 48, 160), (2, 196, 96, 13, 20), (3, 8, 32, 5, 4), (1, 100, 64, 9, 28),
                                       (1, 324, 256, 7, 36), (2, 40, 128, 16, 17 * 4)])
def test_conv1x1_tensor_core_kernel_vs_torch(B, K, N, H, W):
    """b200corr_conv1x1_forward (tcgen05, MN-major operand, tile pairs, ragged K / HW / N) against F.conv2d of the
    TF32-truncated operands."""
    import torch.nn.functional as F
    from understanding_flow_robustness_b200 import raft_corr
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        g = torch.Generator(device="cuda").manual_seed(K + N)
        x = torch.randn(B, K, H, W, device="cuda", generator=g)
        w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
        bias = torch.randn(N, device="cuda", generator=g)
        tr = lambda t: (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
        for relu in (False, True):
            got = raft_corr.conv1x1_forward(x, w, bias, relu=relu)
            ref = F.conv2d(tr(x), tr(w).view(N, K, 1, 1), bias)
            ref = F.relu(ref) if relu else ref
            assert got.shape == ref.shape
            assert float((got - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    finally:
        torch.backends.cudnn.allow_tf32 = prev
