"""In-tree build of the native code (nvcc cross-compiles sm_100a without a GPU).

    pomfret_b200/lib/libpomfret_gpu.so    CUDA kernels + C ABI            (nvcc, sm_100a, -lineinfo)
    pomfret_b200/lib/libpomfret_host.so   host front end as a library     (g++)
    pomfret_b200/bin/pomfret              host front end CLI               (g++)
    pomfret_b200/bin/pomfret-synth        synthetic BAM/VCF generator      (g++)

Built files are git-ignored but travel to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib")
BIN = os.path.join(PKG, "bin")
OBJ = os.path.join(ROOT, "build")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--fmad=false"]


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps if os.path.exists(d))


def _files(d, exts):
    out = []
    for base, _, names in os.walk(d):
        for n in names:
            if n.endswith(exts):
                out.append(os.path.join(base, n))
    return sorted(out)


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA engine cannot be built (there is no CPU fallback)")


def build_hts_objects(verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hts = os.path.join(CSRC, "hts")
    objs = []
    for src in _files(hts, (".c",)):
        obj = os.path.join(OBJ, "hts_" + os.path.basename(src)[:-2] + ".o")
        if not _newer(obj, [src] + _files(os.path.join(hts, "htslib"), (".h",))):
            _run(["gcc", "-O2", "-g", "-fPIC", "-Wall", "-I", hts, "-c", src, "-o", obj], verbose)
        objs.append(obj)
    return objs


def build_gpu(verbose=False, extra_flags=()):
    os.makedirs(LIB, exist_ok=True)
    gpu = os.path.join(CSRC, "gpu")
    hts = os.path.join(CSRC, "hts")
    out = os.path.join(LIB, "libpomfret_gpu.so")
    deps = _files(gpu, (".cu", ".cuh", ".h", ".inc")) + [os.path.join(ROOT, "include", "pomfret_gpu.h"),
                                                        os.path.join(hts, "fisher.c")]
    if _newer(out, deps) and not extra_flags:
        return out
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra_flags) + [
        "-shared", "-I", os.path.join(ROOT, "include"), "-I", gpu, "-I", hts, "-o", out,
        os.path.join(gpu, "engine.cu"), os.path.join(hts, "fisher.c")]
    _run(cmd, verbose)
    return out


def build_host(verbose=False):
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(BIN, exist_ok=True)
    objs = build_hts_objects(verbose)
    hts = os.path.join(CSRC, "hts")
    host = os.path.join(CSRC, "host")
    synth = os.path.join(CSRC, "synth")
    inc = ["-I", os.path.join(ROOT, "include"), "-I", hts, "-I", host, "-I", synth]
    cxx = ["g++", "-O2", "-g", "-std=c++17", "-Wall", "-fPIC"]
    host_srcs = [s for s in _files(host, (".cpp",)) if not s.endswith("_main.cpp")]
    hdrs = _files(host, (".h",)) + _files(os.path.join(hts, "htslib"), (".h",)) + [os.path.join(ROOT, "include", "pomfret_gpu.h")]
    lib = os.path.join(LIB, "libpomfret_host.so")
    if not _newer(lib, host_srcs + hdrs + objs):
        _run(cxx + ["-shared"] + inc + ["-o", lib] + host_srcs + objs + ["-lz", "-ldl", "-pthread"], verbose)
    exe = os.path.join(BIN, "pomfret")
    main = os.path.join(host, "pomfret_main.cpp")
    if os.path.exists(main) and not _newer(exe, [main, lib]):
        _run(cxx + inc + ["-o", exe, main] + host_srcs + objs + ["-lz", "-ldl", "-pthread"], verbose)
    sy = os.path.join(BIN, "pomfret-synth")
    sy_srcs = [os.path.join(synth, "synth.cpp"), os.path.join(synth, "synth_main.cpp")]
    if not _newer(sy, sy_srcs + [os.path.join(synth, "synth.h")] + objs):
        _run(cxx + inc + ["-o", sy] + sy_srcs + objs + ["-lz", "-pthread"], verbose)
    sylib = os.path.join(LIB, "libpomfret_synth.so")
    if not _newer(sylib, sy_srcs[:1] + [os.path.join(synth, "synth.h")] + objs):
        _run(cxx + ["-shared"] + inc + ["-o", sylib, sy_srcs[0]] + objs + ["-lz", "-pthread"], verbose)
    return lib


def build_oracle(verbose=False):
    """Build the checkers (oracle port; the reference harness only where /root/reference exists)."""
    _run(["make", "-C", os.path.join(ROOT, "oracle"), "port"] + ([] if verbose else ["-s"]), verbose)
    if os.path.isdir("/root/reference"):
        _run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"] + ([] if verbose else ["-s"]), verbose)


def build_all(verbose=False):
    build_gpu(verbose)
    build_host(verbose)
    build_oracle(verbose)


if __name__ == "__main__":
    build_all(verbose=True)
    print("ok")
