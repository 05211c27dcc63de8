"""In-tree build of libb200corr.so (sm_100a only) with plain nvcc -- no torch headers involved.

    python -m understanding_flow_robustness_b200.build [--force] [--verbose]

The library is a C-ABI shared object (include/b200corr.h); the Python host layer binds it with
ctypes.  The .so lands next to this file so that it travels to the GPU box with the repo snapshot.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
SUFFIX = os.environ.get("B200_LIB_SUFFIX", "")          # tuning variants: separate objects + library
OBJ = OBJ + SUFFIX
LIB = os.path.join(HERE, f"libb200corr{SUFFIX}.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-I", os.path.join(ROOT, "include"),
    "-I", CSRC,
] + os.environ.get("B200_EXTRA_NVCC", "").split()


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    ms = [os.path.getmtime(os.path.join(ROOT, "include", "b200corr.h"))]
    for f in os.listdir(CSRC):
        if f.endswith((".cuh", ".h")):
            ms.append(os.path.getmtime(os.path.join(CSRC, f)))
    return max(ms)


def _compile(src, verbose, extra):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    cmd = [NVCC, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    if verbose and (r.stdout or r.stderr):
        print(r.stdout, r.stderr, flush=True)
    return obj


def build(force=False, verbose=False, ptxas_info=False):
    """Compile every csrc/*.cu for sm_100a and link libb200corr.so.  Returns the library path."""
    os.makedirs(OBJ, exist_ok=True)
    hdr = _headers_mtime()
    extra = ["-Xptxas", "-v"] if ptxas_info else []
    todo, objs = [], []
    for src in _sources():
        obj = os.path.join(OBJ, src[:-3] + ".o")
        objs.append(obj)
        sm = max(os.path.getmtime(os.path.join(CSRC, src)), hdr)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < sm:
            todo.append(src)
    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            list(ex.map(lambda s: _compile(s, verbose or ptxas_info, extra), todo))
    if todo or not os.path.exists(LIB):
        cmd = [NVCC, "-arch=sm_100a", "-shared", "-o", LIB, *objs, "-lcudart"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv,
              ptxas_info="--ptxas" in sys.argv)
    print(p)
