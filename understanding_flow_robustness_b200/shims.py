"""Make the reference's model files import this implementation without modification.

The reference imports `from spatial_correlation_sampler import spatial_correlation_sample`
(models/submodules.py:5-16, models/FlowNetC*.py, models/PWCNet.py, models/raft/corr.py:12-13),
`import spatial_correlation_sampler_backend` (spatial_correlation_sampler.py:1) and
`import alt_cuda_corr` (models/raft/corr.py:6-10), each under try/except ImportError; RAFT itself does
`from .corr import AlternateCorrBlock, CorrBlock` (models/raft/raft.py:5), a relative import that the
import system resolves through sys.modules["<package>.corr"] first.
"""
import sys

from . import backend, raft_corr
from . import spatial_correlation_sampler as scs


def install_reference_shims(raft_package="models.raft"):
    """Register this implementation under the reference's module names.  `raft_package` is the dotted
    name the reference's RAFT package is imported under (None: leave models/raft/corr.py alone)."""
    sys.modules["spatial_correlation_sampler_backend"] = backend
    sys.modules["spatial_correlation_sampler"] = scs
    sys.modules["alt_cuda_corr"] = raft_corr.alt_cuda_corr
    # FlowNet2's natives (models/channelnorm_package/channelnorm.py:1, models/resample2d_package/resample2d.py:1)
    from . import flownet2_natives
    sys.modules["channelnorm_cuda"] = flownet2_natives.channelnorm_cuda
    sys.modules["resample2d_cuda"] = flownet2_natives.resample2d_cuda
    if raft_package:
        sys.modules[raft_package + ".corr"] = raft_corr
    return scs, raft_corr.alt_cuda_corr
