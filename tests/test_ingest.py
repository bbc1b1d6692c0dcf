"""Compressed ingest: BGZF inflate + BAM record slicing on the device.
  * inflate_kernel against zlib, byte for byte, on the blocks of real region queries at deflate levels 0 (stored
    blocks), 1, 6 and 9 (dynamic codes), and on hand-made members with fixed Huffman codes; corrupted blocks are
    reported (CRC-32 / ISIZE);
  * walk + slice kernels against the host loader: the records sam_itr_next returns for the same query, field by field;
  * the whole hot path fed by add_reads_device() against the oracle port, window by window.
The CPU run steps the kernels on the SIMT emulator; the `gpu` tests run the same checks on the CUDA library."""
import ctypes as C
import struct
import zlib

import numpy as np
import pytest

import conftest
import oracle_bindings as ob
import parity
import pomfret_b200 as pb
from pomfret_b200 import _ffi

READBACK = 50000


def region_of(s, e):
    beg = max(s - READBACK, 0) - 1
    return (max(beg, 0), e + READBACK)


def zlib_inflate_blocks(plan):
    """reference answer: every BGZF block of the plan through zlib, concatenated per stream"""
    comp = np.ctypeslib.as_array(C.cast(plan["comp"], C.POINTER(C.c_uint8)), shape=(plan["comp_bytes"],))
    blocks = C.cast(plan["blocks"], C.POINTER(_ffi.BgzfBlock))
    streams = C.cast(plan["streams"], C.POINTER(_ffi.BgzfStream))
    out = []
    for s in range(plan["n_streams"]):
        S = streams[s]
        data = b""
        for bi in range(S.first_block, S.first_block + S.n_blocks):
            B = blocks[bi]
            raw = bytes(comp[B.comp_off:B.comp_off + B.csize])
            xlen = struct.unpack_from("<H", raw, 10)[0]
            data += zlib.decompress(raw[12 + xlen:-8], -15)
        out.append(data[:S.out_bytes])
    return out


def check_ingest(gpu, data, cov, readlen, max_windows=2, full_pipeline=True):
    host = pb.load_host()
    hb = host.bam_open(data["bam"])
    cfg, ocfg = pb.make_config(cov, readlen=readlen), ob.make_config(cov, readlen=readlen)
    gaps = data["gaps"][:max_windows]
    chrom = gaps[0][0]
    gaps = [g for g in gaps if g[0] == chrom]
    regions = [region_of(s, e) for _, s, e, _ in gaps]
    plan = host.ingest_plan(hb, chrom, regions)
    assert plan["n_streams"] >= len(gaps) and plan["n_blocks"] >= plan["n_streams"]
    ctx = gpu.init()
    b = gpu.batch_begin(ctx)
    flt = _ffi.IngestFilter(cfg.min_mapq, cfg.readlen_threshold, 2, 1, 0.1)
    rc, recs, n = b.ingest_bgzf(plan["comp"], plan["comp_bytes"], plan["blocks"], plan["n_blocks"], plan["streams"], plan["n_streams"], flt)
    # (1) inflate, byte for byte
    for s, want in enumerate(zlib_inflate_blocks(plan)):
        got = bytes(b.inflated(s))
        assert got == want, "stream %d differs from zlib" % s
    # (1b) coverage estimator sum over the sliced records (estimate_read_coverage_dirtyfast's bin loop, blockjoin.c:1016-1021)
    for min_pos, bin_size, n_bins in ((0, 5000, 1 << 20), (regions[0][0] + 20000, 700, (regions[0][1] - 30000) // 700), (0, 1, 10)):
        want = 0
        for i in range(n):
            R = recs[i]
            if R.keep and R.pos >= min_pos:
                want += sum(1 for p in range(R.pos, R.end_pos, bin_size) if p // bin_size < n_bins)
        assert b.ingest_coverage(min_pos, bin_size, n_bins) == want, (min_pos, bin_size, n_bins)
    # (2) records of every query against the host loader (the shim's iterator + filters)
    wins = parity.load_windows(host, hb, gaps, cfg)
    dsz = C.sizeof(_ffi.ReadDesc)
    per_run = {}
    for i in range(n):
        R = recs[i]
        assert not R.bad
        per_run.setdefault(plan["stream_run"][R.stream], []).append(R)
    descs_all, layout = [], []
    for wi, ((w, nw, _, s, e), (beg0, end0)) in enumerate(zip(wins, regions)):
        mine = [R for R in per_run.get(wi, []) if R.keep and R.pos < end0 and R.end_pos > beg0]
        base = host.window_descs(w)
        names = host.window_qnames(w)
        assert len(mine) == nw, (wi, len(mine), nw)
        arr = (_ffi.ReadDesc * max(nw, 1))()
        for k, R in enumerate(mine):
            d = _ffi.ReadDesc.from_address(base + k * dsz)
            assert (R.pos, R.l_qseq, R.n_cigar, R.flag, R.mapq, R.tags_malformed, R.hp, R.mn, R.mm_len, R.ml_len) == \
                   (d.pos, d.l_qseq, d.n_cigar, d.flag, d.mapq, d.tags_malformed, d.hp, d.mn, d.mm_len, d.ml_len), (wi, k)
            assert R.qname.decode() == names[k]
            a = arr[k]
            a.pos, a.l_qseq, a.n_cigar, a.flag, a.mapq = R.pos, R.l_qseq, R.n_cigar, R.flag, R.mapq
            a.tags_malformed, a.hp, a.mn = R.tags_malformed, R.hp, R.mn
            a.cigar, a.seq = R.cigar, R.seq
            a.mm = R.mm if R.has_mm else None
            a.mm_len, a.ml, a.ml_len = R.mm_len, R.ml, R.ml_len
            a.md, a.md_len, a.reserved = None, 0, R.end_pos
        descs_all.append(arr)
    if full_pipeline:
        # (3) the hot path on the ingested records against the oracle
        for (w, nw, _, s, e), arr in zip(wins, descs_all):
            first = b.add_reads_device(arr, nw)
            b.add_window(s, e, first, nw)
            layout.append((first, nw))
        b.submit(); b.decode(cfg.lo, cfg.hi); b.pileup(cfg); b.join(cfg)
        res, tags, ids, rc = b.collect(check=False)
        assert rc == 0
        for wi, ((w, nw, chrom_, s, e), (first, _)) in enumerate(zip(wins, layout)):
            p = ob.port_window(host.window_descs(w), nw, s, e, ocfg)
            bad = parity.compare_window(b, wi, first, nw, res, tags, ids, p)
            assert not bad, (chrom_, s, e, bad[:8])
    for w, *_ in wins:
        host.window_free(w)
    b.end()
    gpu.destroy(ctx)
    host.ingest_free(plan)
    host.bam_close(hb)
    return n


def bgzf_member(payload, deflated):
    """a BGZF block around an already deflated payload"""
    bsize = 12 + 6 + len(deflated) + 8 - 1
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + deflated +
            struct.pack("<II", zlib.crc32(payload) & 0xffffffff, len(payload)))


def check_handmade_members(gpu):
    """fixed-Huffman members (zlib strategy Z_FIXED), a stored member, an empty member; then a flipped byte"""
    rng = np.random.default_rng(3)
    payloads = [bytes(rng.integers(0, 4, size=5000, dtype=np.uint8)), b"A" * 70 + bytes(range(256)) * 20, b"", b"x"]
    members = []
    for i, p in enumerate(payloads):
        co = zlib.compressobj(6, zlib.DEFLATED, -15, 9, zlib.Z_FIXED if i % 2 == 0 else zlib.Z_DEFAULT_STRATEGY)
        members.append(bgzf_member(p, co.compress(p) + co.flush()))
    stored = bytes(rng.integers(0, 256, size=300, dtype=np.uint8))
    co = zlib.compressobj(0, zlib.DEFLATED, -15)
    members.append(bgzf_member(stored, co.compress(stored) + co.flush()))
    payloads.append(stored)
    comp = np.frombuffer(b"".join(members), dtype=np.uint8).copy()
    blocks = (_ffi.BgzfBlock * len(members))()
    streams = (_ffi.BgzfStream * len(members))()
    co_, oo = 0, 0
    for i, (m, p) in enumerate(zip(members, payloads)):
        blocks[i] = _ffi.BgzfBlock(co_, len(m), len(p), oo)
        streams[i] = _ffi.BgzfStream(oo, len(p), len(p), 0, 0, i, 1, 0)  # ubeg = end: no records to walk
        co_ += len(m)
        oo += (len(p) + 15) & ~15
    ctx = gpu.init()
    b = gpu.batch_begin(ctx)
    rc, recs, n = b.ingest_bgzf(comp.ctypes.data, len(comp), blocks, len(members), streams, len(members))
    assert rc == 0 and n == 0
    for i, p in enumerate(payloads):
        assert bytes(b.inflated(i)) == p, i
    # one flipped payload byte: CRC-32 (or the code itself) must catch it
    bad = comp.copy()
    bad[40] ^= 0x10
    b.reset()
    rc, _, _ = b.ingest_bgzf(bad.ctypes.data, len(bad), blocks, len(members), streams, len(members), check=False)
    assert rc != 0
    b.end()
    gpu.destroy(ctx)


@pytest.fixture(scope="module")
def emu_gpu(built):
    import build_emu
    return pb.load_gpu(build_emu.build())


@pytest.fixture(scope="module")
def synth_tiny(built, tmp_path_factory):
    d = tmp_path_factory.mktemp("tiny")
    return conftest.run_synth(str(d / "t"), ["-c", "14", "-s", "13", "-C", "chrT:120000:0-100000", "--readlen", "2500", "--block", "18000",
                                              "--gap", "4000-5000"])


@pytest.mark.emu
def test_ingest_emulated(emu_gpu, synth_tiny):
    assert check_ingest(emu_gpu, synth_tiny, 14, 1200, max_windows=1) > 100


@pytest.mark.emu
def test_inflate_handmade_members_emulated(emu_gpu):
    check_handmade_members(emu_gpu)


@pytest.mark.emu
@pytest.mark.parametrize("level", [0, 6])
def test_ingest_deflate_levels_emulated(emu_gpu, built, tmp_path, level):
    data = conftest.run_synth(str(tmp_path / "lv"), ["-c", "6", "-s", "12", "-C", "chrT:120000:0-100000", "--readlen", "2500", "--block", "18000",
                                                     "--gap", "4000-5000", "-l", str(level), "--qual"])
    check_ingest(emu_gpu, data, 6, 1200, max_windows=1, full_pipeline=False)


@pytest.mark.gpu
def test_ingest_gpu_30x(built, synth30):
    assert check_ingest(pb.load_gpu(), synth30, 30, 15000, max_windows=4) > 300


@pytest.mark.gpu
def test_ingest_gpu_60x(built, synth60):
    check_ingest(pb.load_gpu(), synth60, 60, 15000, max_windows=5)


@pytest.mark.gpu
def test_ingest_gpu_long_cigar(built, synth_long_cigar):
    # CG-tag records: the slicing kernel hands out the real CIGAR
    check_ingest(pb.load_gpu(), synth_long_cigar, 30, 15000, max_windows=1)


@pytest.mark.gpu
def test_inflate_handmade_members_gpu(built):
    check_handmade_members(pb.load_gpu())


@pytest.mark.gpu
@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_ingest_deflate_levels_gpu(built, tmp_path, level):
    data = conftest.run_synth(str(tmp_path / "lv"), ["-c", "30", "-s", "12", "-C", "chr20:64444167:3000000-3600000", "--block", "200000",
                                                     "--gap", "20000-40000", "-l", str(level), "--qual"])
    check_ingest(pb.load_gpu(), data, 30, 15000, max_windows=2)


def _rewrite_case(tmp_path, data, args, gpu_lib, extra_env=None):
    """`methphase --write-bam` through the compressed ingest: the output BAM comes from retag_kernel + host block cutting +
    parallel block compression and the BAI from the block table (Worker::rewrite_bam_device); both must be the bytes the
    reference writes through bam_write1 / sam_index_build3 (blockjoin.c:3022-3103, 4714-4731)."""
    import filecmp
    import os
    import subprocess
    from test_host_frontend import MINE
    env = dict(os.environ)
    if gpu_lib:
        env["POMFRET_GPU_LIB"] = gpu_lib
    env.update(extra_env or {})
    out = {}
    for who, exe in (("ref", ob.REF_BIN), ("mine", MINE)):
        prefix = str(tmp_path / who)
        p = subprocess.run([exe, "methphase"] + args + ["--write-bam", "-o", prefix, "--vcf", data["vcf"], data["bam"]], env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert p.returncode == 0, (who, p.stderr[-2000:])
        out[who] = p.stderr
    assert "host writer" not in out["mine"]
    for suffix in (".mp.gtf", ".mp.vcf", ".mp.bam", ".mp.bam.bai"):
        assert filecmp.cmp(str(tmp_path / "ref") + suffix, str(tmp_path / "mine") + suffix, shallow=False), suffix


@pytest.mark.emu
@pytest.mark.skipif(not __import__("os").path.exists(ob.REF_BIN), reason="oracle/_ref/pomfret not built")
def test_write_bam_device_rewrite_emulated(built, synth_tiny, tmp_path):
    import build_emu
    # two chunks (the cut points come from the linear index), open block carried from one to the next
    _rewrite_case(tmp_path, synth_tiny, ["-c", "14", "-L", "1200", "-t", "3"], build_emu.build(), {"POMFRET_REWRITE_CHUNK_MB": "1"})


@pytest.mark.gpu
@pytest.mark.skipif(not __import__("os").path.exists(ob.REF_BIN), reason="oracle/_ref/pomfret not built")
@pytest.mark.parametrize("chunk_mb", ["1", "48"])
def test_write_bam_device_rewrite_gpu(built, synth30, tmp_path, chunk_mb):
    _rewrite_case(tmp_path, synth30, ["-c", "30", "-t", "6"], None, {"POMFRET_REWRITE_CHUNK_MB": chunk_mb})


@pytest.mark.gpu
@pytest.mark.skipif(not __import__("os").path.exists(ob.REF_BIN), reason="oracle/_ref/pomfret not built")
def test_write_bam_device_rewrite_untagged_gpu(built, tmp_path):
    # -u: the raw tags come from the read haplotagger, most records have no HP tag (appended), unphased ones get 255 (type S)
    data = conftest.run_synth(str(tmp_path / "unt"), ["-c", "30", "-s", "35", "-C", "chr20:64444167:5000000-6000000", "-F", "2", "--untagged"])
    _rewrite_case(tmp_path, data, ["-u", "-c", "30", "-t", "4"], None, {"POMFRET_REWRITE_CHUNK_MB": "4"})


# ---- retag_kernel alone: hand-made records with every kind of HP tag against a plain restatement of bam_aux_update_int ----
def _aux_update_int(rec, val):
    """bam_aux_update_int(aln, "HP", val) on the bytes of one BAM record (block_size included); htslib semantics as the
    reference uses them at blockjoin.c:3092"""
    sz, ty = (1, b"C") if val < 255 else (2, b"S")
    l_qname, n_cigar, l_seq = rec[12], struct.unpack_from("<H", rec, 16)[0], struct.unpack_from("<I", rec, 20)[0]
    p = 36 + l_qname + 4 * n_cigar + (l_seq + 1) // 2 + l_seq
    found = None
    while p + 3 <= len(rec):
        tag, t = rec[p:p + 2], chr(rec[p + 2])
        q = p + 3
        if t in "AcC":
            q += 1
        elif t in "sS":
            q += 2
        elif t in "iIf":
            q += 4
        elif t in "ZH":
            q = rec.index(b"\0", q) + 1
        elif t == "B":
            es = {"c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}[chr(rec[q])]
            q += 5 + es * struct.unpack_from("<I", rec, q + 1)[0]
        if tag == b"HP" and found is None:
            found = (p + 2, t)
        p = q
    if found is None:
        out = bytearray(rec) + b"HP" + ty + val.to_bytes(sz, "little")
    else:
        at, t = found
        old = {"c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4}.get(t)
        if old is None:
            return bytes(rec)                     # not an integer tag: the update fails, the record is written as it is
        if old >= sz:
            out = bytearray(rec)
            out[at:at + 1 + old] = {1: b"C", 2: b"S", 4: b"I"}[old] + val.to_bytes(old, "little")
        else:
            out = bytearray(rec[:at]) + ty + val.to_bytes(sz, "little") + rec[at + 1 + old:]
    struct.pack_into("<I", out, 0, len(out) - 4)
    return bytes(out)


def _handmade_record(rng, i, hp_tag):
    name = ("read%d" % i).encode() + b"x" * int(rng.integers(0, 60)) + b"\0"   # some names longer than the inline prefix
    l_seq = int(rng.integers(1, 400))
    cigar = struct.pack("<I", (l_seq << 4) | 0)
    seq = bytes(rng.integers(0, 256, size=(l_seq + 1) // 2, dtype=np.uint8))
    qual = bytes(rng.integers(0, 60, size=l_seq, dtype=np.uint8))
    aux = b"NMC" + bytes([i & 0xff]) + b"MDZ" + str(l_seq).encode() + b"\0"
    aux += hp_tag
    aux += b"deffff\x80\x3c" if i % 3 == 0 else b"XYBc" + struct.pack("<I", 3) + b"\x01\x02\x03"
    body = struct.pack("<iiBBHHHIiii", 0, 100 + 7 * i, len(name), 60, 4680, 1, 0, l_seq, -1, -1, 0) + name + cigar + seq + qual + aux
    return struct.pack("<I", len(body)) + body


def check_retag(gpu):
    rng = np.random.default_rng(11)
    tags = [b"", b"HPC\x01", b"HPc\x02", b"HPS\x01\x00", b"HPs\x02\x00", b"HPI\x01\0\0\0", b"HPi\x02\0\0\0", b"HPZab\0", b"HPA1"]
    recs = [_handmade_record(rng, i, tags[i % len(tags)]) for i in range(120)]
    payload = b"".join(recs)
    # BGZF blocks of at most 20000 bytes: records straddle them
    members, blocks_py = [], []
    for o in range(0, len(payload), 20000):
        piece = payload[o:o + 20000]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        members.append(bgzf_member(piece, co.compress(piece) + co.flush()))
    comp = np.frombuffer(b"".join(members), dtype=np.uint8).copy()
    blocks = (_ffi.BgzfBlock * len(members))()
    co_, oo = 0, 0
    for i, m in enumerate(members):
        n = min(20000, len(payload) - 20000 * i)
        blocks[i] = _ffi.BgzfBlock(co_, len(m), n, oo)
        co_ += len(m)
        oo += n
    streams = (_ffi.BgzfStream * 1)()
    streams[0] = _ffi.BgzfStream(0, len(payload), 0, -0x80000000, 0, 0, len(members), 0)   # POMFRET_GPU_ANY_TID
    ctx = gpu.init()
    b = gpu.batch_begin(ctx)
    flt = _ffi.IngestFilter(0, 0, 0, 0, 0.0, 1)
    rc, sl, n = b.ingest_bgzf(comp.ctypes.data, len(comp), blocks, len(members), streams, 1, flt)
    assert rc == 0 and n == len(recs)
    vals = np.array([[1, 2, 255, 0][int(v)] for v in rng.integers(0, 4, size=n)], dtype=np.uint8)
    want = [r if v == 0 else _aux_update_int(r, int(v)) for r, v in zip(recs, vals)]
    for i in range(n):
        assert sl[i].rec_bytes == len(recs[i]) and sl[i].tid == 0 and not sl[i].bad
        assert (sl[i].hp_type != 0) == (tags[i % len(tags)] != b"")
    dst = np.zeros(n, np.uint64)
    dst[1:] = np.cumsum([len(w) for w in want])[:-1]
    got = bytes(b.ingest_retag(dst, vals, sum(len(w) for w in want)))
    for i, w in enumerate(want):
        o = int(dst[i])
        assert got[o:o + len(w)] == w, (i, tags[i % len(tags)], int(vals[i]))
    b.end()
    gpu.destroy(ctx)


@pytest.mark.emu
def test_retag_kernel_emulated(emu_gpu):
    check_retag(emu_gpu)


@pytest.mark.gpu
def test_retag_kernel_gpu(built):
    check_retag(pb.load_gpu())
