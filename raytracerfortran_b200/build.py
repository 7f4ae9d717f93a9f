"""Compile libraytrace_b200.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libraytrace_b200.so")
SOURCES = [os.path.join(HERE, "csrc", f) for f in ("rt_kernels.cu", "rt_api.cu")]
DEPS = SOURCES + [os.path.join(HERE, "csrc", "rt_internal.h"),
                  os.path.join(ROOT, "include", "raytrace_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only; no other arch, no PTX fallback
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                                  # belt and braces: the kernels use *_rn intrinsics
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
    "-shared",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libraytrace_b200.so cannot be built")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build_library(force=False, verbose=False, out=None, defines=()):
    """Build (if stale) and return the path of the shared library.  `out` / `defines` build an
    experimental copy (profiling A/B runs) without touching the product library."""
    if out is not None or force or stale():
        cmd = [nvcc_path()] + NVCC_FLAGS + [f"-D{d}" for d in defines] \
            + (["-Xptxas", "-v"] if verbose else []) + ["-o", out or LIB] + SOURCES
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        subprocess.check_call(cmd, env=env)
    return out or LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
