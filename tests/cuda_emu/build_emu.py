"""TEST INFRASTRUCTURE: build the kernel sources against the CPU SIMT emulator (see cuda_emu.h).

Output: tests/cuda_emu/_build/libpomfret_gpu_emu.so — loaded only by tests marked `emu`.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build")


def build(verbose=False, opt="-O1"):
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, "libpomfret_gpu_emu.so")
    gpu = os.path.join(ROOT, "pomfret_b200", "csrc", "gpu")
    hts = os.path.join(ROOT, "pomfret_b200", "csrc", "hts")
    srcs = [os.path.join(gpu, "engine.cu")]
    deps = [os.path.join(gpu, f) for f in os.listdir(gpu)] + [os.path.join(HERE, "cuda_emu.h"), os.path.join(HERE, "cuda_emu.cpp")]
    if os.path.exists(so) and all(os.path.getmtime(so) > os.path.getmtime(d) for d in deps):
        return so
    cmd = ["g++", "-std=c++17", opt, "-g", "-fPIC", "-shared", "-DPOMFRET_CUDA_EMU", "-Wall", "-Wno-unknown-pragmas",
           "-Wno-unused-function", "-Wno-unused-variable", "-Wno-sign-compare",
           "-I", HERE, "-I", gpu, "-I", os.path.join(ROOT, "include"), "-I", hts,
           "-x", "c++", srcs[0], "-x", "c++", os.path.join(HERE, "cuda_emu.cpp"),
           "-x", "c", os.path.join(hts, "fisher.c"), "-lm", "-pthread", "-o", so]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return so


if __name__ == "__main__":
    print(build(verbose=True))
