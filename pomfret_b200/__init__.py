"""pomfret_b200 — B200-native engine for Pomfret's `methphase` hot path.

The product is native code:

* ``lib/libpomfret_gpu.so``   hand-written sm_100a CUDA kernels behind the ``pomfret_gpu_*`` C ABI
  (``include/pomfret_gpu.h``);
* ``lib/libpomfret_host.so`` / ``bin/pomfret``   the host front end (BAM/VCF/GTF I/O, CLI, interval
  bookkeeping, writers) that drives the C ABI.

This Python package is only a thin ctypes binding used by the tests and by ``bench.py``; there is no
Python (or CPU) implementation of the path — loading fails loudly if the CUDA library is missing.
"""
from ._ffi import (Config, ReadDesc, WindowResult, Timing, Variant, BgzfBlock, BgzfStream, IngestFilter, SlicedRecord, GpuLib, HostLib, load_gpu, load_host,
                   make_config, PACKAGE_DIR, LIB_DIR)

__all__ = ["Config", "ReadDesc", "WindowResult", "Timing", "Variant", "BgzfBlock", "BgzfStream", "IngestFilter", "SlicedRecord", "GpuLib", "HostLib", "load_gpu",
           "load_host", "make_config", "PACKAGE_DIR", "LIB_DIR"]
