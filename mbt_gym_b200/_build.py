"""Build libmbt_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
LIB_PATH = os.path.join(_HERE, "libmbt_b200.so")

SOURCES = ["mbt_capi.cu"]
DEPS = ["mbt_capi.cu", "mbt_kernels.cuh", "mbt_step_core.cuh", "mbt_host_params.h", "mbt_variants.h"]
HEADERS = ["mbt_b200.h", "mbt_math.h", "mbt_philox.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",            # no implicit FMA contraction: float results must match the oracle bit-for-bit
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
    "-shared",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libmbt_b200.so cannot be built (there is no CPU fallback)")
    return exe


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    paths = [os.path.join(CSRC, f) for f in DEPS] + [os.path.join(INCLUDE, f) for f in HEADERS]
    return any(os.path.getmtime(p) > built for p in paths)


def build_library(force=False, verbose=False, extra_flags=()):
    """Compile mbt_gym_b200/libmbt_b200.so; returns its path."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, *extra_flags, "-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(" ".join(cmd))
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libmbt_b200.so:\n" + res.stderr[-4000:])
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build_library(force=True, verbose=True, extra_flags=tuple(sys.argv[1:])))
