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


# ---- votes for the phase of unphased variants (recover_variant_phase_in_one_interval, blockjoin.c:2475-2600) ----
def _read_variant_positions(d):
    """positions of the variants a record shows itself (parse_variants_for_one_read, blockjoin.c:1545-1691):
    insertions from the CIGAR, mismatches and deletions from MD — a plain re-derivation for the test"""
    import re
    out = []
    ref = d.pos
    cig = np.ctypeslib.as_array(C.cast(d.cigar, C.POINTER(C.c_uint32)), shape=(d.n_cigar,))
    for c in cig:
        op, ln = int(c) & 15, int(c) >> 4
        if op in (0, 2, 3, 7, 8):
            ref += ln
        elif op == 1:
            out.append(ref)
    md = C.string_at(d.md, d.md_len).decode()
    ref = d.pos
    for num, dele, mis in re.findall(r"(\d+)|(\^[A-Za-z]+)|([A-Za-z])", md):
        if num:
            ref += int(num)
        elif dele:
            out.append(ref)
            ref += len(dele) - 1
        else:
            out.append(ref)
            ref += 1
    return out


def variant_votes_parity(gpu, data, chrom):
    host = _setup_host()
    hb = host.bam_open(data["bam"])
    w = host.lib.pomfret_host_contig_load(hb, chrom.encode())
    n = host.window_n(w)
    cap = 1 << 17
    vars_ = (pb.Variant * cap)()
    bases = np.zeros(cap * 4, np.uint8)
    nb = C.c_int()
    nk = host.lib.pomfret_host_load_variants(data["vcf"].encode(), chrom.encode(), vars_, cap, bases.ctypes.data, cap * 4, C.byref(nb))
    poss = np.array(sorted(vars_[i].pos for i in range(nk)), dtype=np.uint32)
    poss = np.concatenate([poss[:50], poss[10:12], poss[50:]])  # a few duplicated positions
    poss.sort()
    rng = np.random.default_rng(5)
    read_hap = rng.choice(np.array([0, 1, 2, 255], dtype=np.uint8), size=n, p=[0.4, 0.4, 0.1, 0.1])
    descs = host.window_descs(w)
    sz = C.sizeof(pb.ReadDesc)
    want = np.zeros(2 * len(poss) + 1, np.int64)
    last = int(poss[-1])
    for i in range(n):
        if read_hap[i] == 255:
            continue
        for p in _read_variant_positions(pb.ReadDesc.from_address(descs + i * sz)):
            if p >= last:
                want[2 * len(poss)] += 1
            if read_hap[i] < 2:
                for k in range(int(np.searchsorted(poss, p, "left")), int(np.searchsorted(poss, p, "right"))):
                    want[2 * k + int(read_hap[i])] += 1
    ctx = gpu.init([0])
    b = gpu.batch_begin(ctx)
    b.add_reads(descs, n)
    b.submit()
    b.haptag(np.zeros(1, np.uint8), 0, np.zeros(0, np.uint8), np.zeros(max(n, 1), np.uint32))
    b.collect_haptags()
    votes = np.zeros(2 * len(poss) + 1, np.int32)
    gpu.lib.pomfret_gpu_variant_votes.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    assert gpu.lib.pomfret_gpu_variant_votes(b.h, poss.ctypes.data, len(poss), read_hap.ctypes.data, votes.ctypes.data) == 0
    assert np.array_equal(votes.astype(np.int64), want)
    assert want[:-1].sum() > 100
    b.end()
    gpu.destroy(ctx)
    host.window_free(w)
    host.bam_close(hb)


@pytest.mark.emu
def test_variant_votes_emulated(untagged_small):
    import build_emu
    variant_votes_parity(pb.load_gpu(build_emu.build()), untagged_small, "chrU")


@pytest.mark.gpu
def test_variant_votes_gpu(untagged_small):
    variant_votes_parity(pb.load_gpu(), untagged_small, "chrU")
