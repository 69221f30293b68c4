"""Builds libspx.so (hand-written sm_100a CUDA + the C ABI of include/spx.h) in-tree with nvcc.

The built library lives next to this file so that it travels to the GPU box with the repo snapshot.
nvcc cross-compiles without a GPU, so this is also the CPU-side "does it build" check.
"""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libspx.so")
SOURCES = ["spx_engine.cu", "spx_tower.cu", "spx_tttnet.cu", "spx_replay.cu", "spx_train.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libspx.so cannot be built (there is no CPU fallback)")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(_HERE), "include", "spx.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a into libspx.so.  -fmad=false keeps the fp64 PUCT arithmetic free
    of FMA contraction (bit-exact with the reference's separate IEEE operations)."""
    if not force and not _stale():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [_nvcc(), *ARCH, "-O3", "-lineinfo", "-std=c++17", "-fmad=false", "-Xcompiler", "-fPIC", "-shared",
           "-o", LIB, *srcs, "-lcuda" if False else "-lcudart"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
