"""haptag_kernel probe (GPU box): the -u read haplotagger over every primary record of an untagged 30x sample.
usage: python profiles/haptag_probe.py [contigs] [contig_mb] [replicas]"""
import ctypes as C, os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, conftest
import pomfret_b200 as pb
from pomfret_b200 import _ffi
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 2
mb = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
rep = int(sys.argv[3]) if len(sys.argv) > 3 else 8
tmp = tempfile.mkdtemp()
data = conftest.run_synth(os.path.join(tmp, "u"), bench.synth_args(nc, mb, 30, 130, tagged=False))
host = pb.load_host(); gpu = pb.load_gpu()
lib = host.lib
lib.pomfret_host_contig_load.restype = C.c_void_p
lib.pomfret_host_contig_load.argtypes = [C.c_void_p, C.c_char_p]
lib.pomfret_host_load_variants.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
hb = host.bam_open(data["bam"])
ctx = gpu.init([0])
b = gpu.batch_begin(ctx)
dsz = C.sizeof(_ffi.ReadDesc)
for name, _ in bench.HG38[:nc]:
    w = lib.pomfret_host_contig_load(hb, name.encode())
    n = host.window_n(w)
    cap = 1 << 18
    vars_ = (pb.Variant * cap)()
    vb = np.zeros(cap * 4, np.uint8)
    nbv = C.c_int()
    nk = lib.pomfret_host_load_variants(data["vcf"].encode(), name.encode(), vars_, cap, vb.ctypes.data, cap * 4, C.byref(nbv))
    known = np.frombuffer(vars_, dtype=np.uint8, count=nk * C.sizeof(pb.Variant)).copy()
    var_pos = np.array([vars_[i].pos for i in range(nk)], dtype=np.uint32)
    descs = host.window_descs(w)
    starts = np.array([_ffi.ReadDesc.from_address(descs + i * dsz).pos for i in range(n)], dtype=np.uint32)
    kf = np.searchsorted(var_pos, starts, side="left").astype(np.uint32)
    b.reset()
    for _ in range(rep):
        b.add_reads(descs, n)
    b.submit()
    kfr = np.tile(kf, rep)
    for it in range(3):
        if it:
            b.rewind()
        b.haptag(known, nk, vb[:nbv.value], kfr)
        tags, st = b.collect_haptags()
    t = b.timing()
    print("%s: %d records x %d, %d known variants, haptag %.3f ms, %.1f GB/s algorithmic, %.2f M reads/s, tags %s"
          % (name, n, rep, nk, t.haptag_ms, t.haptag_bytes / t.haptag_ms / 1e6, n * rep / t.haptag_ms / 1e3,
             dict(zip(*np.unique(tags, return_counts=True)))))
    host.window_free(w)
b.end()
