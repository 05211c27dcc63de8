"""RAFT lookup fused with the motion encoder's 1x1 convolution (models/raft/raft.py:189 +
models/raft/update.py:104,111) against the unfused chain lookup -> F.conv2d -> relu.

Tolerance: samples and weights are rounded to TF32 inside the fused kernel, accumulation is fp32, so against the
fp32 convolution |err| <= 2^-10 * sum_k |w_k * corr_k| (+1e-5 accumulation slack); against the same convolution
with TF32-rounded operands (what cuDNN computes under torch's default allow_tf32) <= 1e-4 relative."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _tf32(x, trunc=False):
    # round-to-nearest (ties away) to 10 explicit mantissa bits: cvt.rna.tf32.f32 -- or plain truncation, what the
    # tensor core does with an fp32 operand it is handed as is
    i = x.contiguous().view(torch.int32)
    return ((i if trunc else i + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.fixture(autouse=True)
def _fp32_library_math():
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize("B,C,H,W,L,r,n_out,layout,sigma", [
    (1, 32, 16, 32, 4, 4, 256, "auto", 2.0),        # basic model: 324 -> 256, blocked fine levels
    (2, 16, 24, 40, 4, 4, 256, "rowmajor", 3.0),    # two samples, partial last tile (960 queries = 7.5 tiles)
    (1, 16, 16, 24, 4, 3, 96, "auto", 2.0),         # small model: 196 -> 96 (update.py:79), one accumulator half
    (1, 8, 13, 20, 3, 2, 64, "auto", 30.0),         # odd height, scalar / vec4 paths, windows off the map
    (1, 8, 8, 16, 2, 1, 32, "rowmajor", 1.0),
    (4, 64, 48, 160, 4, 4, 256, "auto", 3.0),       # BASELINE config 3 at full size
    (5, 16, 16, 32, 4, 4, 256, "auto", 2.0),        # more than one chunk of 4 samples in the pipelined variant
])
@pytest.mark.parametrize("impl", ["fused", "pipelined"])
def test_fused_lookup_convc1_equals_unfused_chain(B, C, H, W, L, r, n_out, layout, sigma, impl):
    from understanding_flow_robustness_b200 import CorrBlock, _lib, coords_grid
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + H + n_out)
    f1 = torch.randn(B, C, H, W, device="cuda", generator=g)
    f2 = torch.randn(B, C, H, W, device="cuda", generator=g)
    nin = L * (2 * r + 1) ** 2
    conv = torch.nn.Conv2d(nin, n_out, 1).cuda()
    coords = coords_grid(B, H, W, "cuda") + sigma * torch.randn(B, 2, H, W, device="cuda", generator=g)
    with torch.no_grad():
        blk = CorrBlock(f1, f2, L, r, precision="tf32", layout=layout)
        corr = blk(coords)
        want = F.relu(conv(corr))
        n0 = _lib.lib().b200corr_launch_count()
        got = blk.lookup_convc1(coords, conv.weight, conv.bias, impl=impl)
        assert _lib.lib().b200corr_launch_count() - n0 >= 1          # prepare (first call) + the kernel(s)
        assert got.shape == want.shape == (B, n_out, H, W)
        # documented bound against the fp32 convolution: operands rounded (fused: 2 x 2^-11) or truncated
        # (pipelined: the tensor core reads the fp32 tensors as they lie, 2 x 2^-10)
        tr = impl == "pipelined" and nin % 4 == 0 and (H * W) % 4 == 0      # else the fused kernel runs
        bound = (2.0 ** -9 if tr else 2.0 ** -10) * F.conv2d(corr.abs(), conv.weight.abs()) + 1e-5
        assert bool(((got - want).abs() <= bound).all()), float(((got - want).abs() - bound).max())
        # and tight against the convolution of the TF32 operands
        ref = F.relu(F.conv2d(_tf32(corr, tr), _tf32(conv.weight, tr), conv.bias))
        assert float((got - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
        # no ReLU, no bias
        got2 = blk.lookup_convc1(coords, conv.weight, None, relu=False, impl=impl)
        ref2 = F.conv2d(_tf32(corr, tr), _tf32(conv.weight, tr))
        assert float((got2 - ref2).abs().max()) <= 1e-4 * float(ref2.abs().max())
        # a second call with other coordinates reuses the prepared weights
        c2 = coords + 0.37
        ref3 = F.relu(F.conv2d(_tf32(blk(c2), tr), _tf32(conv.weight, tr), conv.bias))
        got3 = blk.lookup_convc1(c2, conv.weight, conv.bias, impl=impl)
        assert float((got3 - ref3).abs().max()) <= 1e-4 * float(ref3.abs().max())


def test_fused_path_falls_back_to_the_reference_chain_under_autograd():
    """Gradients are those of the reference ops: with differentiable features the unfused chain runs."""
    from understanding_flow_robustness_b200 import CorrBlock, coords_grid
    torch.manual_seed(0)
    f1 = torch.randn(1, 16, 16, 32, device="cuda", requires_grad=True)
    f2 = torch.randn(1, 16, 16, 32, device="cuda", requires_grad=True)
    conv = torch.nn.Conv2d(324, 256, 1).cuda()
    coords = coords_grid(1, 16, 32, "cuda") + torch.randn(1, 2, 16, 32, device="cuda")
    blk = CorrBlock(f1, f2, 4, 4)
    out = blk.lookup_convc1(coords, conv.weight, conv.bias)
    out.square().mean().backward()
    assert f1.grad is not None and float(f1.grad.abs().max()) > 0 and conv.weight.grad is not None


def test_unmodified_raft_update_block_with_the_fused_motion_encoder():
    """The reference's BasicMotionEncoder (update.py:94-121) with its first two lines replaced by the fused call:
    same motion features (TF32 bound) inside the unmodified RAFT forward."""
    from understanding_flow_robustness_b200.harness import reference_models
    if not reference_models.available():
        pytest.skip("reference model files not staged")
    from understanding_flow_robustness_b200.harness.raft_fused import fuse_motion_encoder
    torch.manual_seed(4)
    net = reference_models.reference_raft(iters=3).cuda().eval()
    i1 = 255 * torch.rand(1, 3, 128, 256, device="cuda")
    i2 = 255 * torch.rand(1, 3, 128, 256, device="cuda")
    with torch.no_grad():
        want = net(i1, i2)[-1]
        with fuse_motion_encoder(net) as calls:
            got = net(i1, i2)[-1]
        assert calls["fused"] == 3
        again = net(i1, i2)[-1]                       # the patch is undone on exit
    assert torch.equal(again, want)
    assert float((got - want).abs().max()) <= 5e-3 * float(want.abs().max())
