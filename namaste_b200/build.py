"""Build the in-tree CUDA library (sm_100a) with nvcc.  No torch involved: the library is a
plain C-ABI shared object (include/namaste_b200.h) that links the CUDA runtime only."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnamaste_b200.so")
# api.cu holds the C-ABI entry points; inst_s*.cu instantiate the pipeline for one supersampling factor each
# (see csrc/api_internal.cuh), so the translation units compile in parallel
SOURCES = ["api.cu", "inst_s0.cu", "inst_s1.cu", "inst_s10.cu", "inst_s11.cu", "inst_s15.cu"]
HEADERS = ["api_internal.cuh", "transit_math.cuh", "lnprob_kernels.cuh", "eval_pipeline.cuh", "sampler.cuh", "gp_kernels.cuh",
           "flatten_kernels.cuh", "summary_kernels.cuh", "split_kernels.cuh",
           os.path.join("..", "..", "include", "namaste_b200.h")]
OBJDIR = os.path.join(HERE, "build")
NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    # the transit arithmetic mirrors the reference operation by operation; FMAs are
    # written explicitly (fma()) only where a single rounding is provably harmless
    "-fmad=false",
    "-Xcompiler", "-fPIC",
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


def build_library(force=False, verbose=False, extra_flags=(), out=None, sources=None):
    """Compile csrc/*.cu -> namaste_b200/libnamaste_b200.so (skipped when up to date).
    `extra_flags` / `out` build an experimental variant beside it (A/B runs through NMB_LIB)."""
    variant = out is not None or bool(extra_flags)
    if not variant and not force and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    lib = out or LIB
    objdir = OBJDIR if not variant else os.path.join(OBJDIR, os.path.splitext(os.path.basename(lib))[0])
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", "-o", obj, os.path.join(CSRC, src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, res.stdout, res.stderr))
        if verbose:
            sys.stderr.write(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = lib + ".tmp"
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs,
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, lib)
    if variant:
        shutil.rmtree(objdir, ignore_errors=True)   # keep the tree small: the snapshot sent to the GPU box is size-limited
    return lib


if __name__ == "__main__":
    # python -m namaste_b200.build [--force] [-v] [--out path.so -DNAME=VALUE ...]
    flags = [a for a in sys.argv[1:] if a.startswith("-D")]
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, extra_flags=flags, out=out))
