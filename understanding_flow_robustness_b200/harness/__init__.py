"""Benchmark / attack harness around the correlation hot path (plain torch layers + our operators).

Not part of the drop-in surface: the reference's model bodies (models/FlowNetC.py, models/raft/*)
run unmodified on top of the operators via `install_reference_shims()`.  These restatements exist
because the reference tree is not available on the benchmark box.
"""
from .flownetc import FlowNetCHarness, correlate  # noqa: F401
