"""TEST INFRASTRUCTURE ONLY — builds tests/emu/libfastace_emu.so: the step kernels' CUDA source compiled for the
CPU under the SIMT emulator warp_emu.h (g++, no nvcc, no GPU)."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libfastace_emu.so")
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"


def deps():
    return ([os.path.join(HERE, "emu_step.cpp"), os.path.join(HERE, "warp_emu.h"), os.path.abspath(__file__)] +
            glob.glob(os.path.join(ROOT, "fastace_b200", "csrc", "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h")))


def needs_build():
    return not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps())


def build(force=False):
    if not force and not needs_build():
        return LIB
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-attributes", "-Wno-unknown-pragmas",
           "-x", "c++", "-I", CUDA_INC, "-I", HERE, "-o", LIB, os.path.join(HERE, "emu_step.cpp")]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("emulator build failed:\n" + out.stdout + out.stderr[-6000:])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
