"""Hand-built alignment records for the decode kernel: MM/ML grammar variants, both strands, clips, indels,
long reads (several SEQ scan steps and bitmap windows), dense and sparse lists, malformed tags.
The same pomfret_gpu_read_desc array goes to the CUDA path and to the oracle port."""
import ctypes as C

import numpy as np

import oracle_bindings as ob
import pomfret_b200 as pb
from pomfret_b200 import _ffi

NT16 = {"A": 1, "C": 2, "G": 4, "T": 8, "N": 15, "M": 3, "R": 5}
COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def pack_seq(s):
    codes = np.array([NT16[c] for c in s] + ([0] if len(s) & 1 else []), dtype=np.uint8)
    return ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8)


def make_cigar(rng, qlen, kind):
    """-> list of (op, len) whose query-consuming lengths sum to qlen"""
    ops = []
    left = qlen
    if kind in ("clip", "clipall") and left > 4:
        c = int(rng.integers(1, min(300, left - 2)))
        ops.append((4, c)); left -= c
    tail = 0
    if kind in ("clip", "tail") and left > 8:
        tail = int(rng.integers(1, min(200, left - 4))); left -= tail
    if kind == "fatal_head":
        ops.append((5, 7))
    err = {"clean": 0.0, "dense": 0.08}.get(kind, 0.012)
    while left > 0:
        m = left if err == 0 else min(left, int(rng.geometric(err)))
        ops.append((0, m)); left -= m
        if left <= 0:
            break
        r = rng.random()
        if r < 0.45:
            i = min(left, int(rng.geometric(0.6)))
            if left - i < 1:
                continue
            ops.append((1, i)); left -= i
        elif r < 0.9:
            ops.append((2, int(rng.geometric(0.6))))
        elif kind == "skip" and r < 0.95:
            ops.append((3, 50))
        elif kind == "fatal_mid" and r < 0.97:
            ops.append((7, 1) if left > 1 else (0, 1)); left -= 1
        else:
            ops.append((2, 1))
    if tail:
        ops.append((4, tail))
    return ops


def list_for(rng, seq, rev, base, mode, frac=1.0):
    """positions (scan order from the read's 5' end) of `base` on the original strand -> listed subset -> deltas"""
    want = COMP[base] if rev else base
    idx = [i for i, c in enumerate(seq) if c == want or base == "N"]
    if rev:
        idx = idx[::-1]
    n = len(seq)

    def is_cpg(p):
        if not (0 < p < n - 1):
            return True  # edge bases: dropped silently by the reference, never "implicit"
        return seq[p + 1] == "G" if not rev else seq[p - 1] == "C"
    if mode == "cpg":
        keep = [k for k, p in enumerate(idx) if base != "C" or is_cpg(p)]
        if frac < 1.0:
            keep = [k for k in keep if rng.random() < frac]
    elif mode == "all":
        keep = list(range(len(idx)))
    elif mode == "random":
        keep = [k for k in range(len(idx)) if rng.random() < frac]
    elif mode == "edges":
        keep = [k for k in (0, len(idx) - 1) if 0 <= k < len(idx)]
        keep = sorted(set(keep))
    else:
        keep = []
    deltas, prev = [], -1
    for k in keep:
        deltas.append(k - prev - 1); prev = k
    return deltas


class Records:
    def __init__(self):
        self.keep = []
        self.descs = []
        self.names = []

    def add(self, name, pos, flag, seq, cigar, mm, ml, hp=0, mn=-1, malformed=0):
        d = _ffi.ReadDesc()
        d.pos, d.l_qseq, d.n_cigar, d.flag, d.mapq = pos, len(seq), len(cigar), flag, 60
        d.tags_malformed, d.hp, d.mn = malformed, hp, mn
        cg = np.array([(l << 4) | op for op, l in cigar], dtype=np.uint32)
        sq = pack_seq(seq)
        self.keep += [cg, sq]
        d.cigar = cg.ctypes.data if len(cg) else None
        d.seq = sq.ctypes.data if len(sq) else None
        if mm is not None:
            mb = np.frombuffer(mm.encode(), dtype=np.uint8).copy() if mm else np.zeros(1, np.uint8)
            self.keep.append(mb)
            d.mm, d.mm_len = mb.ctypes.data, len(mm)
        else:
            d.mm, d.mm_len = None, 0
        if ml is not None:
            mlb = np.array(ml, dtype=np.uint8) if len(ml) else np.zeros(1, np.uint8)
            self.keep.append(mlb)
            d.ml, d.ml_len = mlb.ctypes.data, len(ml)
        else:
            d.ml, d.ml_len = None, -1
        d.md, d.md_len = None, 0
        self.descs.append(d)
        self.names.append(name)

    def array(self):
        arr = (_ffi.ReadDesc * len(self.descs))(*self.descs)
        self.keep.append(arr)
        return arr


def build_records(seed, n_random=60, max_len=9000, long_lens=(70000,)):
    rng = np.random.default_rng(seed)
    R = Records()

    def rand_seq(n, comp):
        p = {"uniform": [0.25, 0.25, 0.25, 0.25, 0, 0], "crich": [0.1, 0.45, 0.35, 0.1, 0, 0],
             "poor": [0.48, 0.02, 0.02, 0.48, 0, 0], "iupac": [0.24, 0.24, 0.24, 0.24, 0.02, 0.02]}[comp]
        return "".join(rng.choice(list("ACGTNM"), size=n, p=p))

    def one(name, n, rev, comp, cig_kind, style, mode, frac=1.0, ml_mode="ok", overshoot=0):
        seq = rand_seq(n, comp)
        cigar = make_cigar(rng, n, cig_kind)
        flag = 16 if rev else 0
        segs, ml = [], []

        def seg(base, codes, mark, deltas):
            segs.append("%s+%s%s%s;" % (base, codes, mark, "".join(",%d" % d for d in deltas)))
            nc = 1 if codes.isdigit() else len(codes)
            ml.extend(int(x) for x in rng.integers(0, 256, size=len(deltas) * nc))
        dl = list_for(rng, seq, rev, "C", mode, frac)
        if overshoot and dl:
            dl = dl + [overshoot]
        mark = rng.choice(["?", ".", ""])
        if style == "m":
            seg("C", "m", mark, dl)
        elif style == "hm":
            seg("C", "hm", mark, dl)
        elif style == "mh":
            seg("C", "mh", mark, dl)
        elif style == "h;m":
            seg("C", "h", mark, list_for(rng, seq, rev, "C", "random", 0.05))
            seg("C", "m", mark, dl)
        elif style == "a;m":
            seg("A", "a", mark, list_for(rng, seq, rev, "A", "random", 0.03))
            seg("C", "m", mark, dl)
        elif style == "m;a;N":
            seg("C", "m", mark, dl)
            seg("A", "a", mark, list_for(rng, seq, rev, "A", "random", 0.02))
            seg("N", "n", mark, list_for(rng, seq, rev, "N", "random", 0.01))
        elif style == "chebi":
            seg("C", "76792", mark, list_for(rng, seq, rev, "C", "random", 0.05))
            seg("C", "m", mark, dl)
        elif style == "m;m":
            seg("C", "m", mark, dl)
            seg("C", "m", mark, list_for(rng, seq, rev, "C", "cpg", 0.3))
        elif style == "onlyh":
            seg("C", "h", mark, dl)
        elif style == "minus":
            segs.append("G-m%s%s;" % (mark, "".join(",%d" % d for d in list_for(rng, seq, rev, "G", "random", 0.02))))
            ml.extend(int(x) for x in rng.integers(0, 256, size=segs[-1].count(",")))
            seg("C", "m", mark, dl)
        mm = "".join(segs)
        if ml_mode == "none":
            mlv = None
        elif ml_mode == "short":
            mlv = ml[:-1] if ml else [1]
        elif ml_mode == "long":
            mlv = ml + [7]
        else:
            mlv = ml
        R.add(name, int(rng.integers(1000, 5000000)), flag, seq, cigar, mm, mlv)

    # deterministic corner cases
    for rev in (0, 1):
        for n in (1, 2, 3, 31, 32, 33, 1023, 1024, 1025, 2047, 2048, 2049, 4097):
            one("len%d" % n, n, rev, "crich", "plain", "m", "all")
            one("len%d_cpg" % n, n, rev, "uniform", "clip", "m", "cpg")
        one("edges", 700, rev, "crich", "clean", "m", "edges")
        one("empty_list", 500, rev, "uniform", "plain", "m", "none")
        one("overshoot", 800, rev, "uniform", "plain", "m", "cpg", overshoot=5)
        one("overshoot_big", 800, rev, "poor", "plain", "m", "cpg", overshoot=100000)
        one("noml", 900, rev, "uniform", "plain", "m", "cpg", ml_mode="none")
        one("mlshort", 900, rev, "uniform", "plain", "m", "cpg", ml_mode="short")
        one("mllong", 900, rev, "uniform", "plain", "m", "cpg", ml_mode="long")
        one("clipall", 400, rev, "uniform", "clipall", "m", "cpg")
        one("skip", 3000, rev, "uniform", "skip", "m", "cpg")
        one("fatal_head", 600, rev, "uniform", "fatal_head", "m", "cpg")
        one("fatal_mid", 3000, rev, "uniform", "fatal_mid", "m", "cpg")
        one("dense_cigar", 6000, rev, "uniform", "dense", "m", "cpg")
        one("dense_cpg_list", 9000, rev, "crich", "plain", "m", "cpg")      # > 64 listed bases per SEQ tile
        one("big_deltas", 30000, rev, "crich", "plain", "m", "cpg", 0.002)  # 3- to 5-digit deltas
        for style in ("hm", "mh", "h;m", "a;m", "m;a;N", "chebi", "m;m", "onlyh", "minus"):
            one("style_" + style, 2500, rev, "uniform", "plain", style, "cpg")
            one("style_" + style + "_iupac", 1800, rev, "iupac", "clip", style, "cpg", 0.7)
        for n in long_lens:
            one("long%d" % n, n, rev, "uniform", "plain", "m", "cpg")
            one("long%d_crich_all" % n, n // 2, rev, "crich", "plain", "m", "all")
            one("long%d_sparse" % n, n, rev, "crich", "clip", "h;m", "cpg", 0.05)
    # malformed strings
    seq = rand_seq(300, "uniform")
    cg = [(0, 300)]
    for i, mm in enumerate(["C+m?,1,2", "C+m?,1,,2;", "C+m?,x;", "Q+m?,1;", "C*m,1;", "C+m?,1;;", ";", "C+m", "C+,1;",
                            "C+m?1;", "C+m?,99999999999;", "C+m?,4294967295,3;"]):
        R.add("bad%d" % i, 5000, 0, seq, cg, mm, [5] * mm.count(","))
        R.add("bad%d_rev" % i, 5000, 16, seq, cg, mm, [5] * mm.count(","))
    R.add("no_mm", 100, 0, seq, cg, None, None)
    R.add("empty_mm", 100, 0, seq, cg, "", [])
    R.add("flagged_malformed", 100, 0, seq, cg, "C+m?,1;", [9], malformed=1)
    R.add("no_cigar", 100, 0, seq, [], "C+m?,1;", [9])
    # random mix
    styles = ["m", "m", "m", "hm", "h;m", "a;m", "chebi", "m;a;N"]
    cigs = ["plain", "plain", "clip", "tail", "clean", "dense", "skip"]
    for i in range(n_random):
        n = int(rng.integers(40, max_len))
        one("rnd%d" % i, n, int(rng.integers(0, 2)), rng.choice(["uniform", "crich", "poor", "iupac"]),
            rng.choice(cigs), rng.choice(styles), rng.choice(["cpg", "cpg", "cpg", "random"]),
            float(rng.choice([1.0, 1.0, 0.5, 0.1])))
    return R


def pack_into_one_buffer(R, seed=0):
    """Move every field of every record into one byte buffer at random (mis)alignments and point the
    descriptors there: what a loader's record buffer looks like.  Returns (address, size) of the buffer."""
    rng = np.random.default_rng(seed)
    fields = []
    total = 64
    for d in R.descs:
        for name, size in (("cigar", d.n_cigar * 4), ("seq", (d.l_qseq + 1) // 2), ("mm", d.mm_len if d.mm else 0),
                           ("ml", d.ml_len if d.ml_len > 0 else 0), ("md", d.md_len if d.md else 0)):
            if getattr(d, name) and size:
                off = total + int(rng.integers(0, 16))
                if name == "cigar":
                    off = (off + 3) & ~3  # CIGAR words stay 4-byte aligned, as inside a BAM record
                fields.append((d, name, size, off))
                total = off + size
    buf = np.zeros(total + 64, dtype=np.uint8)
    base = buf.ctypes.data
    for d, name, size, off in fields:
        C.memmove(base + off, getattr(d, name), size)
        setattr(d, name, base + off)
    R.keep.append(buf)
    return base, len(buf)


def check_against_port(gpu, R, lo=100, hi=156, register=None):
    """Decode R on `gpu` (CUDA or emulated library) and with the oracle port; return the list of differences."""
    arr = R.array()
    n = len(R.descs)
    ctx = gpu.init()
    if register:
        gpu.host_register(ctx, register[0], register[1])
    b = gpu.batch_begin(ctx)
    b.add_reads(arr, n)
    b.submit()
    b.decode(lo, hi)
    if register:
        assert b.timing().launches == 3, "the records were expected to be gathered by the device"  # gather, decode, generic decode
    port = ob.port_lib()
    bad = []
    n_overflow = 0
    for i in range(n):
        out = ob.PortCalls()
        st, end = C.c_uint32(), C.c_uint32()
        port.port_decode_read(C.addressof(arr) + i * C.sizeof(_ffi.ReadDesc), lo, hi, C.byref(out), C.byref(st),
                              C.byref(end))
        gst, gnc, gend = b.read_info(i)
        ppos = np.ctypeslib.as_array(out.pos, shape=(out.n,)).copy() if out.n else np.zeros(0, np.uint32)
        pcat = np.ctypeslib.as_array(out.cat, shape=(out.n,)).copy() if out.n else np.zeros(0, np.uint8)
        if (gst & 15) != (st.value & 15):
            bad.append((R.names[i], i, "status", gst, st.value))
            continue
        if gend != end.value:
            bad.append((R.names[i], i, "end", gend, end.value))
        if gst & 64:
            n_overflow += 1  # out of call slots: pileup() would re-run it with more room (covered by the window tests)
            continue
        if st.value & 1:
            gpos, gcat = b.calls(i)
            if not (np.array_equal(gpos, ppos) and np.array_equal(gcat, pcat)):
                bad.append((R.names[i], i, "calls", len(gpos), len(ppos)))
    b.end()
    if register:
        gpu.host_unregister(ctx, register[0])
    gpu.destroy(ctx)
    if n_overflow > n // 10:
        bad.append(("too many call-slot overflows", n_overflow, n))
    return bad
