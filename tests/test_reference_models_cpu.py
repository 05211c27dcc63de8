"""CPU-side checks of the harness against the reference's own model files (no compute through our
operators here: there is no GPU).  Skipped where neither /root/reference nor baseline/_ref exists."""
import pytest
import torch

from understanding_flow_robustness_b200.harness import FlowNetCHarness, reference_models

pytestmark = pytest.mark.skipif(not reference_models.available(), reason="reference model files not present")


def test_harness_has_the_reference_flownetc_parameters():
    """models/FlowNetC.py:11-64: same parameter names and shapes, so state_dicts are interchangeable."""
    ref = reference_models.reference_flownetc()
    ours = FlowNetCHarness()
    rs, os_ = ref.state_dict(), ours.state_dict()
    assert sorted(rs.keys()) == sorted(os_.keys())
    for k in rs:
        assert rs[k].shape == os_[k].shape, k
    ours.load_state_dict(rs)          # strict
    assert sum(p.numel() for p in ours.parameters()) == 39175298      # "Parameter count", FlowNetC.py:8


def test_reference_files_resolve_to_this_package():
    import understanding_flow_robustness_b200 as b200
    sub = reference_models.import_reference("submodules")
    assert sub.spatial_correlation_sample is b200.spatial_correlation_sample
    raft = reference_models.import_reference("raft.raft")
    assert raft.CorrBlock is b200.CorrBlock and raft.AlternateCorrBlock is b200.AlternateCorrBlock
    net = reference_models.reference_raft(iters=2)
    assert sum(p.numel() for p in net.parameters()) > 5_000_000


def test_cpu_tensors_are_refused_not_silently_computed():
    """There is no CPU fallback: the reference body on CPU tensors must fail loudly in our operator."""
    net = reference_models.reference_flownetc().eval()
    net.normalize_correctly = lambda im: im - 0.4        # the reference's version calls .cuda()
    x = torch.rand(1, 3, 64, 64)
    with pytest.raises(RuntimeError):
        net(x, x)
