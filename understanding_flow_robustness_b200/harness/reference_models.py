"""The reference's own, UNMODIFIED model bodies running on this package's operators.

`install_reference_shims()` registers this package under the module names the reference imports
(`spatial_correlation_sampler`, `alt_cuda_corr`, `models.raft.corr`); this module then imports the
reference's `models/FlowNetC.py` and `models/raft/raft.py` from a reference tree -- `/root/reference`
where it exists, else the verbatim copies `baseline/stage_reference.py` put under `baseline/_ref/`
(git-ignored, shipped to the GPU box) -- WITHOUT going through `models/__init__.py` (which imports
every model family and their third-party dependencies).

    net = reference_flownetc()                     # models/FlowNetC.py:11-197, correlate() -> our sampler
    net = reference_raft(iters=12)                 # models/raft/raft.py:25-233, CorrBlock -> ours

Nothing here computes anything: the conv stacks are the reference's, the correlation operators are
this package's CUDA kernels.  A missing tree raises (no restated network is substituted silently).
"""
import argparse
import importlib
import os
import sys
import types
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
CANDIDATES = ("/root/reference", os.path.join(_ROOT, "baseline", "_ref"))


def reference_root():
    """First tree that holds the reference's models/FlowNetC.py, or None."""
    for root in CANDIDATES:
        if os.path.isfile(os.path.join(root, "models", "FlowNetC.py")):
            return root
    return None


def available():
    return reference_root() is not None


def _models_package():
    root = reference_root()
    if root is None:
        raise RuntimeError("reference model files not found (looked in %s); run baseline/stage_reference.py where "
                           "/root/reference exists" % (CANDIDATES,))
    pkg = sys.modules.get("models")
    want = os.path.join(root, "models")
    if pkg is None or want not in list(getattr(pkg, "__path__", [])):
        pkg = types.ModuleType("models")
        pkg.__path__ = [want]
        sys.modules["models"] = pkg
    return pkg


def import_reference(name):
    """Import `models.<name>` from the reference tree with this package's operators shimmed in."""
    from ..shims import install_reference_shims

    install_reference_shims()
    _models_package()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module("models." + name)


def reference_flownetc(**kw):
    """models/FlowNetC.py:11 `FlowNetC(batchNorm=False, div_flow=20)`, random init as at :53-62."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")      # init.uniform / init.xavier_uniform deprecation
        return import_reference("FlowNetC").FlowNetC(**kw)


def raft_args(**over):
    """The attributes models/raft/raft.py:25-95 and models/raft/update.py:95-121 read from `args`."""
    a = argparse.Namespace(small=False, flowNetCEnc=False, no_separate_context=False, fnorm="instance", cnorm="batch",
                           iters=12, corr_levels=4, corr_radius=4, mixed_precision=False, dropout=0,
                           alternate_corr=False, compute_spatial=False, update_no_motion_downsampling=False)
    for k, v in over.items():
        setattr(a, k, v)
    return a


def reference_raft(**over):
    """models/raft/raft.py:25 `RAFT(args)`; keyword overrides go into the args namespace."""
    return import_reference("raft.raft").RAFT(raft_args(**over))
