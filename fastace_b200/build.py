"""In-tree build of libfastace_b200.so for sm_100a (explicit nvcc, no JIT cache).

The built library lives next to the sources (fastace_b200/libfastace_b200.so); it is
git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfastace_b200.so")
SOURCES = [os.path.join(CSRC, "fastace_capi.cu"), os.path.join(CSRC, "fastace_host.cpp"), os.path.join(CSRC, "legacy_capi.cpp")]
def _deps():
    """Every file the library is compiled from: all of csrc/ and the public headers (a stale
    .so must never travel to the GPU box, so nothing is listed by hand)."""
    import glob
    inc = os.path.join(HERE, "..", "include")
    return sorted(set(SOURCES + glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                      glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cpp")) +
                      glob.glob(os.path.join(inc, "*.h")) + glob.glob(os.path.join(inc, "*.hpp")) +
                      [os.path.abspath(__file__)]))


NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--shared", "-Xcompiler", "-fPIC",
    "-cudart", "static",
    # keep the reference's fp64 operation order: no FMA contraction of a*b+c in step code
    "--fmad=false",
]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=False, extra_flags=(), out=None):
    """`out`: write a variant build (profiling defines in `extra_flags`) somewhere else than the product library."""
    if out is None and not force and not needs_build():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra_flags) + ["-o", out or LIB] + SOURCES + ["-ldl"]
    if verbose:
        print(" ".join(cmd))
    out_path = out or LIB
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + out.stdout + out.stderr)
    if verbose and (out.stdout or out.stderr):
        print(out.stdout + out.stderr)
    return out_path


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True,
          extra_flags=["-Xptxas", "-v"] if "--ptxas" in sys.argv else [])
    print(LIB)
