"""Cross-compile the UNMODIFIED reference CUDA extensions for sm_100a into oracle/_ref/.

TEST / BENCH INFRASTRUCTURE ONLY.  Sources are compiled where they lie under /root/reference
(nothing is copied).  Outputs:
  oracle/_ref/ref_sampler_cuda/ref_sampler_cuda.so  <- correlation.cpp + correlation_sampler.cpp +
                                                       correlation_cuda_kernel.cu  (-DUSE_CUDA)
  oracle/_ref/ref_alt_cuda_corr/ref_alt_cuda_corr.so <- alt_cuda_corr/correlation.cpp +
                                                        correlation_kernel.cu
  oracle/_ref/ref_resample2d_cuda/ref_resample2d_cuda.so <- resample2d_package/resample2d_cuda.cc + resample2d_kernel.cu
They travel to the GPU box with the gpurun snapshot and are used there as
  * the "kernel to beat" timed beside ours by bench.py (never as the product path), and
  * a live pin of oracle/raft_oracle.alt_corr_* against the real alt_cuda_corr (tests -m gpu).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = "/root/reference/models"


def _load(name, sources, cflags, cuda_cflags):
    os.environ.setdefault("CXX", "/usr/bin/g++")
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils.cpp_extension import load

    d = os.path.join(OUT, name)
    os.makedirs(d, exist_ok=True)
    load(name=name, sources=sources, extra_cflags=cflags, extra_cuda_cflags=cuda_cflags,
         extra_ldflags=["-L/usr/lib/gcc/x86_64-linux-gnu/13", "-lgomp"],
         build_directory=d, verbose=True, is_python_module=False)
    return os.path.join(d, name + ".so")


def so_path(name):
    return os.path.join(OUT, name, name + ".so")


def build(which=("ref_sampler_cuda", "ref_alt_cuda_corr", "ref_resample2d_cuda")):
    if not os.path.isdir(REF):
        return
    cm = os.path.join(REF, "Pytorch-Correlation-extension/Correlation_Module")
    if "ref_sampler_cuda" in which and not os.path.exists(so_path("ref_sampler_cuda")):
        _load("ref_sampler_cuda",
              [f"{cm}/correlation.cpp", f"{cm}/correlation_sampler.cpp", f"{cm}/correlation_cuda_kernel.cu"],
              ["-fopenmp", "-O3", "-DUSE_CUDA"], ["-DUSE_CUDA", "-O3"])
    if "ref_alt_cuda_corr" in which and not os.path.exists(so_path("ref_alt_cuda_corr")):
        ac = os.path.join(REF, "alt_cuda_corr")
        _load("ref_alt_cuda_corr", [f"{ac}/correlation.cpp", f"{ac}/correlation_kernel.cu"],
              ["-O3"], ["-O3"])
    # FlowNet2's resample2d compiles unmodified against torch 2.11; channelnorm does not (its AT_DISPATCH takes the
    # removed `tensor.type()` form, channelnorm_kernel.cu:111,152), so that op is pinned against its formula only
    if "ref_resample2d_cuda" in which and not os.path.exists(so_path("ref_resample2d_cuda")):
        rs = os.path.join(REF, "resample2d_package")
        _load("ref_resample2d_cuda", [f"{rs}/resample2d_cuda.cc", f"{rs}/resample2d_kernel.cu"], ["-O3"], ["-O3"])


def load_module(name):
    """Import a prebuilt reference CUDA extension as a Python module (GPU box)."""
    import importlib.util

    import torch  # noqa: F401

    spec = importlib.util.spec_from_file_location(name, so_path(name))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    build(tuple(sys.argv[1:]) or ("ref_sampler_cuda", "ref_alt_cuda_corr", "ref_resample2d_cuda"))
    for n in ("ref_sampler_cuda", "ref_alt_cuda_corr", "ref_resample2d_cuda"):
        print(n, os.path.exists(so_path(n)))
