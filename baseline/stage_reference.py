"""Stage the UNMODIFIED reference model bodies into baseline/_ref/ (git-ignored; travels to the GPU box).

The reference is a script collection without a package (`pip install /root/reference` has nothing to
install), so its "install" is a verbatim copy of the handful of Python files the
drop-in demonstration needs.  Nothing here is product source and nothing is committed: baseline/_ref/
is listed in .gitignore.  The files are used
  * by tests/test_reference_models_gpu.py: the reference's own FlowNetC / RAFT bodies run forward +
    backward with this package's operators shimmed in, against the same bodies on the reference's own
    compiled sm_100a extensions / torch CorrBlock;
  * by bench.py's attack rows: the network under attack is the reference's FlowNetC.py, not a restatement.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = "/root/reference"

FILES = [
    "models/FlowNetC.py",
    "models/submodules.py",
    "models/raft/__init__.py",
    "models/raft/raft.py",
    "models/raft/corr.py",
    "models/raft/extractor.py",
    "models/raft/update.py",
    "models/raft/utils/__init__.py",
    "models/raft/utils/utils.py",
    # the reference's Python wrapper of its own sampler extension (bound to the compiled reference backend in tests)
    "models/Pytorch-Correlation-extension/Correlation_Module/spatial_correlation_sampler/__init__.py",
    "models/Pytorch-Correlation-extension/Correlation_Module/spatial_correlation_sampler/spatial_correlation_sampler.py",
]


def stage(force=False):
    """Copy the files if the reference tree is present; returns the staged root or None."""
    if not os.path.isdir(REF):
        return OUT if os.path.isdir(os.path.join(OUT, "models")) else None
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(OUT, rel)
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
    return OUT


if __name__ == "__main__":
    p = stage(force="--force" in sys.argv)
    print("staged reference model bodies:", p)
    sys.exit(0 if p else 1)
