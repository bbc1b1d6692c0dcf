"""The C ABI: every symbol declared in include/*.h is exported by libpomfret_gpu.so, and the library
refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import glob
import os
import re

import pytest

import pomfret_b200 as pb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU_SO = os.path.join(pb.LIB_DIR, "libpomfret_gpu.so")


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names.update(re.findall(r"\b(pomfret_gpu_[a-z_0-9]+)\s*\(", text))
    return sorted(names)


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("pomfret_gpu_init", "pomfret_gpu_batch_add_read", "pomfret_gpu_decode", "pomfret_gpu_haptag",
                 "pomfret_gpu_pileup", "pomfret_gpu_join", "pomfret_gpu_batch_collect"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(GPU_SO), "libpomfret_gpu.so missing: run __graft_entry__.build()"
    lib = C.CDLL(GPU_SO)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    gpu = pb.load_gpu()
    assert gpu.device_count() == 0
    ctx = C.c_void_p()
    rc = gpu.lib.pomfret_gpu_init(C.byref(ctx), None, 0, 1)
    assert rc == -1 and "no CPU fallback" in gpu.strerror(rc)


def test_product_does_not_reference_the_oracle():
    """Nothing under pomfret_b200/ may include, link or import oracle/ (the oracle is test infrastructure)."""
    bad = []
    for base, _, names in os.walk(os.path.join(ROOT, "pomfret_b200")):
        for n in names:
            if n.endswith((".so", ".o", ".pyc")) or "/bin" in base:
                continue
            text = open(os.path.join(base, n), errors="ignore").read()
            if re.search(r"oracle/|oracle_bindings|liboracle|port_window|libpomfret_ref", text):
                bad.append(os.path.join(base, n))
    assert not bad, bad
