"""The reference's own UNMODIFIED model bodies (models/FlowNetC.py, models/raft/raft.py) run forward +
backward on this package's operators through `install_reference_shims()` and are compared with the SAME
bodies, same weights, same inputs on the reference's own operators:

  FlowNetC : `correlate()` on the reference's sampler wrapper + its CUDA kernels compiled unmodified for
             sm_100a (oracle/_ref/ref_sampler_cuda)                          models/FlowNetC.py:134-139
  RAFT     : the reference's torch `CorrBlock` (matmul / avg_pool2d / grid_sample, fp32 matmul)
                                                                             models/raft/raft.py:150-159,189

Tolerances: flow and input gradient <= 1e-4 relative (max-norm) for FlowNetC (fp32 sampler, 1e-5 per
call); RAFT with the split-TF32 volume (`tf32x3`, the shim default) <= 1e-3 after 4 recurrent updates,
and within the documented TF32 drift (<= 5e-2) with `precision="tf32"`.
The model files come from baseline/_ref/ (verbatim copies staged by baseline/stage_reference.py).
"""
import importlib.util
import os
import sys

import pytest
import torch

from understanding_flow_robustness_b200.harness import FlowNetCHarness, reference_models

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not reference_models.available(), reason="reference model files not staged")]


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.fixture(autouse=True)
def _fp32_library_math():
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


def _reference_sampler_on_reference_cuda_kernels():
    """The reference's Python wrapper (spatial_correlation_sampler.py) bound to its own CUDA extension."""
    from oracle import build_ref_cuda

    if not os.path.exists(build_ref_cuda.so_path("ref_sampler_cuda")):
        pytest.skip("oracle/_ref/ref_sampler_cuda not built")
    backend = build_ref_cuda.load_module("ref_sampler_cuda")
    wrapper = os.path.join(reference_models.reference_root(), "models", "Pytorch-Correlation-extension",
                           "Correlation_Module", "spatial_correlation_sampler", "spatial_correlation_sampler.py")
    keep = sys.modules.get("spatial_correlation_sampler_backend")
    sys.modules["spatial_correlation_sampler_backend"] = backend     # what `import ... as correlation` binds (:1)
    try:
        spec = importlib.util.spec_from_file_location("_reference_scs_wrapper", wrapper)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if keep is not None:
            sys.modules["spatial_correlation_sampler_backend"] = keep
    assert mod.correlation is backend
    return mod.spatial_correlation_sample


def _flownetc_fwd_bwd(net, i1, i2):
    a = i1.clone().requires_grad_(True)
    b = i2.clone().requires_grad_(True)
    flow = net(a, b)
    (flow * torch.linspace(0.5, 1.5, flow.numel(), device=flow.device).view_as(flow)).mean().backward()
    return flow.detach(), a.grad, b.grad


def test_unmodified_flownetc_on_our_sampler_equals_it_on_the_reference_cuda_sampler():
    from understanding_flow_robustness_b200 import _lib
    import understanding_flow_robustness_b200 as b200

    ref_fn = _reference_sampler_on_reference_cuda_kernels()
    torch.manual_seed(0)
    net = reference_models.reference_flownetc().cuda().eval()
    sub = reference_models.import_reference("submodules")
    assert sub.spatial_correlation_sample is b200.spatial_correlation_sample
    i1 = torch.rand(2, 3, 192, 256, device="cuda")
    i2 = torch.rand(2, 3, 192, 256, device="cuda")
    n0 = _lib.lib().b200corr_launch_count()
    ours = _flownetc_fwd_bwd(net, i1, i2)
    assert _lib.lib().b200corr_launch_count() - n0 >= 3           # forward + two backward kernels went through the C ABI
    sub.spatial_correlation_sample = ref_fn                       # correlate() looks the name up in its module globals
    try:
        n1 = _lib.lib().b200corr_launch_count()
        ref = _flownetc_fwd_bwd(net, i1, i2)
        assert _lib.lib().b200corr_launch_count() == n1           # ... and this run did not
    finally:
        sub.spatial_correlation_sample = b200.spatial_correlation_sample
    assert ours[0].shape == (2, 2, 192, 256)
    for o, r, what in zip(ours, ref, ("flow", "grad img1", "grad img2")):
        assert torch.isfinite(o).all() and float(r.abs().max()) > 0
        assert _rel(o, r) <= 1e-4, (what, _rel(o, r))


@pytest.mark.parametrize("fused", [False, True])
def test_harness_equals_the_reference_body_with_the_same_weights(fused):
    """harness/flownetc.py (the fused-merge variant of the network) against models/FlowNetC.py."""
    torch.manual_seed(1)
    ref = reference_models.reference_flownetc().cuda().eval()
    ours = FlowNetCHarness(fused_merge=fused).cuda().eval()
    ours.load_state_dict(ref.state_dict())
    i1 = torch.rand(2, 3, 128, 192, device="cuda")
    i2 = torch.rand(2, 3, 128, 192, device="cuda")
    o, r = _flownetc_fwd_bwd(ours, i1, i2), _flownetc_fwd_bwd(ref, i1, i2)
    for x, y, what in zip(o, r, ("flow", "grad img1", "grad img2")):
        assert _rel(x, y) <= 1e-5, (what, fused, _rel(x, y))


def _reference_torch_corrblock():
    """models/raft/corr.py itself (NOT the shim), imported next to the shimmed module."""
    reference_models.import_reference("raft.raft")
    path = os.path.join(reference_models.reference_root(), "models", "raft", "corr.py")
    spec = importlib.util.spec_from_file_location("models.raft._reference_corr", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.CorrBlock


def _raft_fwd_bwd(net, i1, i2):
    a = i1.clone().requires_grad_(True)
    b = i2.clone().requires_grad_(True)
    preds = net(a, b)
    sum(p.abs().mean() for p in preds).backward()
    return preds[-1].detach(), a.grad, b.grad


@pytest.mark.parametrize("precision,tol_flow,tol_grad", [("tf32x3", 1e-3, 5e-3), ("fp32", 1e-3, 5e-3), ("tf32", 5e-2, 2e-1)])
def test_unmodified_raft_on_our_corrblock_equals_it_on_the_reference_torch_corrblock(precision, tol_flow, tol_grad,
                                                                                      monkeypatch):
    import understanding_flow_robustness_b200 as b200
    from understanding_flow_robustness_b200 import _lib

    monkeypatch.setenv("B200CORR_VOLUME_PRECISION", precision)
    RefCorrBlock = _reference_torch_corrblock()
    raft_mod = reference_models.import_reference("raft.raft")
    assert raft_mod.CorrBlock is b200.CorrBlock
    torch.manual_seed(2)
    net = reference_models.reference_raft(iters=4).cuda().eval()
    i1 = 255 * torch.rand(1, 3, 128, 256, device="cuda")
    i2 = 255 * torch.rand(1, 3, 128, 256, device="cuda")
    n0 = _lib.lib().b200corr_launch_count()
    ours = _raft_fwd_bwd(net, i1, i2)
    assert _lib.lib().b200corr_launch_count() - n0 >= 1 + 4 + 4   # volume build, 4 lookups, 4 lookup backwards at least
    raft_mod.CorrBlock = RefCorrBlock
    try:
        ref = _raft_fwd_bwd(net, i1, i2)
    finally:
        raft_mod.CorrBlock = b200.CorrBlock
    assert ours[0].shape == (1, 2, 128, 256)
    errs = [_rel(o, r) for o, r in zip(ours, ref)]
    assert errs[0] <= tol_flow and max(errs[1:]) <= tol_grad, (precision, errs)


def test_unmodified_raft_with_alternate_corr_block_matches_the_dense_block():
    """args.alternate_corr (raft.py:147-148): the alt path through our alt_cuda_corr drop-in, forward only."""
    torch.manual_seed(3)
    net = reference_models.reference_raft(iters=3).cuda().eval()
    i1 = 255 * torch.rand(1, 3, 128, 256, device="cuda")
    i2 = 255 * torch.rand(1, 3, 128, 256, device="cuda")
    with torch.no_grad():
        dense = net(i1, i2)[-1]
        net.args.alternate_corr = True
        alt = net(i1, i2)[-1]
    assert _rel(alt, dense) <= 1e-3
