"""Read haplotagging kernel (row c): per-read tags and vote counts against the oracle port, and the
resulting qname table against the compiled reference's own -u pre-pass."""
import ctypes as C
import os

import numpy as np
import pytest

import conftest
import oracle_bindings as ob
import pomfret_b200 as pb


def _setup_host():
    host = pb.load_host()
    lib = host.lib
    lib.pomfret_host_contig_load.restype = C.c_void_p
    lib.pomfret_host_contig_load.argtypes = [C.c_void_p, C.c_char_p]
    lib.pomfret_host_load_variants.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    return host


def haptag_parity(gpu, data, chrom):
    host = _setup_host()
    hb = host.bam_open(data["bam"])
    w = host.lib.pomfret_host_contig_load(hb, chrom.encode())
    n = host.window_n(w)
    assert n > 100
    cap = 1 << 17
    vars_ = (pb.Variant * cap)()
    bases = np.zeros(cap * 4, np.uint8)
    nb = C.c_int()
    nk = host.lib.pomfret_host_load_variants(data["vcf"].encode(), chrom.encode(), vars_, cap, bases.ctypes.data, cap * 4, C.byref(nb))
    assert 0 < nk <= cap
    known = np.frombuffer(vars_, dtype=np.uint8, count=nk * C.sizeof(pb.Variant)).copy()
    descs = host.window_descs(w)
    sz = C.sizeof(pb.ReadDesc)
    starts = np.array([pb.ReadDesc.from_address(descs + i * sz).pos for i in range(n)], dtype=np.uint32)
    port = ob.port_lib()
    kf = np.zeros(n, np.uint32)
    port.port_haptag_cursors(starts.ctypes.data, n, known.ctypes.data, nk, kf.ctypes.data)
    ctx = gpu.init([0])
    b = gpu.batch_begin(ctx)
    b.add_reads(descs, n)
    b.submit()
    b.haptag(known, nk, bases[:nb.value], kf)
    tags, status = b.collect_haptags()
    votes = np.zeros(2 * n, np.int32)
    gpu.lib.pomfret_gpu_debug_get_votes.argtypes = [C.c_void_p, C.c_void_p]
    assert gpu.lib.pomfret_gpu_debug_get_votes(b.h, votes.ctypes.data) == 0
    assert not status.any()
    hist = {}
    for i in range(n):
        pv = (C.c_int * 2)()
        t = port.port_haptag_read(descs + i * sz, known.ctypes.data, nk, bases.ctypes.data, int(kf[i]), pv)
        assert t == tags[i], (i, t, tags[i], list(pv), votes[2 * i:2 * i + 2])
        assert list(pv) == list(votes[2 * i:2 * i + 2]), (i, list(pv), votes[2 * i:2 * i + 2])
        hist[t] = hist.get(t, 0) + 1
    assert hist.get(0, 0) > 10 and hist.get(1, 0) > 10
    if os.path.exists(ob.REF_SO):
        ref = ob.ref_lib()
        with ob.quiet_reference():
            st = ref.refh_pre_haplotag(data["vcf"].encode(), data["bam"].encode())
        names = host.window_qnames(w)
        seen = set()
        for i, qn in enumerate(names):
            if qn in seen:
                continue
            seen.add(qn)
            assert ref.refh_tag_lookup(st, qn.encode()) == tags[i], (qn, i)
        assert ref.refh_tags_count(st) == len(seen)
        ref.refh_tags_free(st)
    b.end()
    gpu.destroy(ctx)
    host.window_free(w)
    host.bam_close(hb)


@pytest.fixture(scope="module")
def untagged_small(built, tmp_path_factory):
    d = tmp_path_factory.mktemp("unt")
    return conftest.run_synth(str(d / "unt"), ["-c", "20", "-s", "41", "-C", "chrU:300000:0-150000", "--readlen", "5000",
                                               "--block", "60000", "--gap", "9000-12000", "--untagged", "--err", "0.03"])


@pytest.mark.emu
def test_haptag_emulated(untagged_small):
    import build_emu
    haptag_parity(pb.load_gpu(build_emu.build()), untagged_small, "chrU")


@pytest.mark.gpu
def test_haptag_gpu(untagged_small):
    haptag_parity(pb.load_gpu(), untagged_small, "chrU")


@pytest.mark.gpu
def test_haptag_gpu_long_reads(built, tmp_path):
    data = conftest.run_synth(str(tmp_path / "untl"), ["-c", "30", "-s", "43", "-C", "chr20:64444167:7000000-8000000", "--untagged"])
    haptag_parity(pb.load_gpu(), data, "chr20")
