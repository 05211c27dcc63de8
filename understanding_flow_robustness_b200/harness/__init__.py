"""Harness around the correlation hot path for BASELINE configs 2 and 4.

Not part of the drop-in surface.  `reference_models` imports the reference's UNMODIFIED model bodies
(models/FlowNetC.py, models/raft/*) on top of this package's operators -- from /root/reference, or from
the verbatim staged copies under baseline/_ref/ on the GPU box.  `FlowNetCHarness` is the same FlowNetC
(same parameter names, a reference state_dict loads unchanged) with the merge block optionally on the
fused kernel.
"""
from .flownetc import FlowNetCHarness, correlate  # noqa: F401
from . import reference_models  # noqa: F401
