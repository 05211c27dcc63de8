"""Build the UNMODIFIED reference CPU implementation of the sampler into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this.

The reference's CPU path (models/Pytorch-Correlation-extension/Correlation_Module/
correlation.cpp + correlation_sampler.cpp) is compiled from the sources where they
lie under /root/reference -- no source is copied into this repo.  The output
(`oracle/_ref/spatial_correlation_sampler_backend.so`) is git-ignored but travels to
the GPU box with the gpurun snapshot, where it serves as
  * the validation target for oracle/sampler_oracle.c (tests/test_oracle_cpu.py), and
  * the `cpu_baseline` / `--impl reference` arm of bench.py ("kind": "reference").

Recipe follows SURVEY.md section 8(c): the reference's own setup.py needs $CC and a CUDA
build, so we call torch.utils.cpp_extension.load on the two CPU sources directly
(without -DUSE_CUDA the pybind module binds correlation_cpp_forward/backward,
correlation_sampler.cpp:131-136).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = "/root/reference/models/Pytorch-Correlation-extension/Correlation_Module"
NAME = "spatial_correlation_sampler_backend"


def so_path():
    return os.path.join(OUT, NAME + ".so")


def build(verbose=False):
    """Compile if the reference tree is present; returns path of the .so or None."""
    if os.path.exists(so_path()):
        return so_path()
    if not os.path.isdir(REF):
        return None  # GPU box: only the prebuilt file is used
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("CXX", "/usr/bin/g++")  # the image's default g++ wrapper lacks libgomp.spec
    from torch.utils.cpp_extension import load

    load(
        name=NAME,
        sources=[os.path.join(REF, "correlation.cpp"), os.path.join(REF, "correlation_sampler.cpp")],
        extra_cflags=["-fopenmp", "-O3"],
        extra_include_paths=["/usr/local/cuda/include"],
        extra_ldflags=["-L/usr/lib/gcc/x86_64-linux-gnu/13", "-lgomp"],
        build_directory=OUT,
        verbose=verbose,
    )
    return so_path() if os.path.exists(so_path()) else None


def load_backend():
    """Import the prebuilt reference backend module (forward/backward)."""
    p = so_path()
    if not os.path.exists(p):
        raise FileNotFoundError(p)
    import importlib.util

    import torch  # noqa: F401  (must be imported before the extension)

    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(verbose=True)
    print("reference CPU backend:", p)
    sys.exit(0 if p else 1)
