"""Build the in-tree CUDA library (sm_100a) with nvcc.  No torch involved: the library is a
plain C-ABI shared object (include/namaste_b200.h) that links the CUDA runtime only."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnamaste_b200.so")
SOURCES = ["api.cu"]
HEADERS = ["transit_math.cuh", "lnprob_kernels.cuh", "eval_pipeline.cuh", "sampler.cuh", "gp_kernels.cuh", "flatten_kernels.cuh", "summary_kernels.cuh", os.path.join("..", "..", "include", "namaste_b200.h")]
NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    # the transit arithmetic mirrors the reference operation by operation; FMAs are
    # written explicitly (fma()) only where a single rounding is provably harmless
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build namaste_b200/libnamaste_b200.so")
    return exe


def _stale():
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > built for d in deps if os.path.exists(d))


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu -> namaste_b200/libnamaste_b200.so (skipped when up to date)."""
    if not force and not _stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
