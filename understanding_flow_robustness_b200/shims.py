"""Make the reference's model files import this implementation without modification.

The reference imports `from spatial_correlation_sampler import spatial_correlation_sample`
(models/submodules.py:5-16, models/FlowNetC*.py, models/PWCNet.py, models/raft/corr.py:12-13),
`import spatial_correlation_sampler_backend` (spatial_correlation_sampler.py:1) and
`import alt_cuda_corr` (models/raft/corr.py:6-10), each under try/except ImportError.
"""
import sys

from . import backend, raft_corr
from . import spatial_correlation_sampler as scs


def install_reference_shims():
    sys.modules["spatial_correlation_sampler_backend"] = backend
    sys.modules["spatial_correlation_sampler"] = scs
    sys.modules["alt_cuda_corr"] = raft_corr.alt_cuda_corr
    return scs, raft_corr.alt_cuda_corr
