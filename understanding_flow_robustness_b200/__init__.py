"""B200-native (sm_100a) cost-volume correlation hot path of lmb-freiburg/understanding_flow_robustness.

Public surface (same names / signatures as the reference's operator API):
  spatial_correlation_sample, SpatialCorrelationSampler, SpatialCorrelationSamplerFunction
      <- models/Pytorch-Correlation-extension/Correlation_Module/spatial_correlation_sampler/
  CorrBlock, AlternateCorrBlock, alt_cuda_corr            <- models/raft/corr.py, models/alt_cuda_corr/
  ChannelNorm, Resample2d   FlowNet2's native ops  <- models/channelnorm_package, models/resample2d_package
  warp              PWC-Net's warp() (grid_sample of the map and of a ones mask, threshold, multiply) as one kernel
      <- models/PWCNet.py:164-204
  correlate_merge   correlate() -> LeakyReLU -> cat of the FlowNetC merge block as one kernel
      <- models/submodules.py:124-138 + models/FlowNetC.py:133-147
  install_reference_shims()  registers `spatial_correlation_sampler`, `spatial_correlation_sampler_backend`
      and `alt_cuda_corr` in sys.modules so the reference's model files import this implementation.

All compute goes through the C-ABI library libb200corr.so (include/b200corr.h); there is no CPU
or PyTorch fallback: a missing library or a non-CUDA tensor raises.
"""
from .spatial_correlation_sampler import (  # noqa: F401
    SpatialCorrelationSampler,
    SpatialCorrelationSamplerFunction,
    spatial_correlation_sample,
)
from . import backend  # noqa: F401


def __getattr__(name):
    # RAFT pieces are imported lazily so `import understanding_flow_robustness_b200` stays cheap
    if name in ("CorrBlock", "AlternateCorrBlock", "alt_cuda_corr", "bilinear_sampler", "coords_grid"):
        from . import raft_corr

        return getattr(raft_corr, name)
    if name in ("correlate_merge", "CorrelateMergeFunction"):
        from . import merge_block

        return getattr(merge_block, name)
    if name in ("ChannelNorm", "Resample2d", "ChannelNormFunction", "Resample2dFunction"):
        from . import flownet2_natives

        return getattr(flownet2_natives, name)
    if name == "warp":
        from .pwc_warp import warp

        return warp
    if name == "install_reference_shims":
        from .shims import install_reference_shims

        return install_reference_shims
    raise AttributeError(name)
