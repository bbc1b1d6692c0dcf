"""MD-tag fuzz for the read haplotagger (row a5/a6): hand-made records whose MD strings exercise the tokenizer of
parse_variants_for_one_read (reference blockjoin.c:1590-1672) beyond what an aligner writes — numbers that straddle
the 16-character lane pieces and the 512-character steps of the kernel, runs of more than 16 digits, leading zeros,
adjacent mismatch letters without a number between them, carets inside a deletion run, a deletion run that never
closes, known variants on equal positions — against the sequential oracle port, tag and both vote counts per record."""
import ctypes as C

import numpy as np
import pytest

import oracle_bindings as ob
import pomfret_b200 as pb

NT16 = {"A": 1, "C": 2, "G": 4, "T": 8}


def make_md(rng, style):
    """returns (md string, read bases consumed, reference bases consumed)"""
    toks = []
    n_read = n_ref = 0
    n_tok = int(rng.integers(1, 260 if style != "short" else 12))
    for _ in range(n_tok):
        r = rng.random()
        num = int(rng.integers(0, 120))
        if style == "long_numbers" and rng.random() < 0.15:
            s = "0" * int(rng.integers(1, 30)) + str(num)           # more digits than one lane holds
        elif style == "long_numbers" and rng.random() < 0.15:
            s = str((int(rng.integers(1, 1 << 40)) << 32) + num)    # natoi wraps (blockjoin.c:117-130): the value is num
        elif rng.random() < 0.1:
            s = "0" * int(rng.integers(1, 4)) + str(num)
        else:
            s = str(num)
        toks.append(s)
        n_read += num
        n_ref += num
        if r < 0.55:
            k = 1 if style in ("standard", "long_numbers") or rng.random() < 0.5 else int(rng.integers(2, 45))  # adjacent letters
            toks.append("".join(rng.choice(list("ACGTN"), size=k)))
            n_read += k
            n_ref += k
        else:
            k = int(rng.integers(1, 25 if style != "short" else 4))
            run = "^" + "".join(rng.choice(list("AC"), size=k))
            if style == "weird" and rng.random() < 0.3:
                run += "^" + "".join(rng.choice(list("ACGT"), size=int(rng.integers(1, 4))))  # a caret inside the run
            toks.append(run)
            n_ref += len(run) - 1
    if style == "weird" and rng.random() < 0.5:
        toks.append("^ACG")          # a run that never closes: no variant comes of it
    else:
        num = int(rng.integers(0, 90))
        toks.append(str(num))
        n_read += num
        n_ref += num
    return "".join(toks), n_read, n_ref


def build_records(rng, n, style):
    keep = []     # numpy arrays the descriptors point into
    descs = (pb.ReadDesc * n)()
    spans = []
    pos = 1000
    for i in range(n):
        md, n_read, n_ref = make_md(rng, style)
        # CIGAR: soft clip, matches cut by a few insertions (the MD walk steps over them), soft clip
        lead = int(rng.integers(0, 3)) * int(rng.integers(0, 40))
        ops = []
        if lead:
            ops.append((4, lead))
        left = n_read
        n_ins = int(rng.integers(0, 6 if style != "weird" else 80))
        total_ins = 0
        for _ in range(n_ins):
            if left < 2:
                break
            a = int(rng.integers(1, left))
            ops.append((0, a))
            left -= a
            x = int(rng.integers(1, 7))
            ops.append((1, x))
            total_ins += x
        ops.append((0, max(left, 0)) if left > 0 else (0, 0))
        ops = [o for o in ops if o[1] > 0 or o[0] != 0]
        l_qseq = lead + n_read + total_ins
        if l_qseq == 0:
            l_qseq = 1
            ops = [(4, 1)]
        cigar = np.array([(ln << 4) | op for op, ln in ops], dtype=np.uint32)
        bases = rng.choice([1, 2], size=l_qseq + (l_qseq & 1)).astype(np.uint8)   # A / C only: known alleles match often
        seq = ((bases[0::2] << 4) | bases[1::2]).astype(np.uint8)
        seq = np.concatenate([seq, np.zeros(16, np.uint8)])
        mdb = np.frombuffer(md.encode() + b"\0" * 16, dtype=np.uint8).copy()
        keep += [cigar, seq, mdb]
        d = descs[i]
        d.pos, d.l_qseq, d.n_cigar, d.flag, d.mapq = pos, l_qseq, len(cigar), 0, 60
        d.hp, d.mn, d.ml_len = 254, -1, -1
        d.cigar, d.seq, d.md, d.md_len = cigar.ctypes.data, seq.ctypes.data, mdb.ctypes.data, len(md)
        spans.append((pos, pos + max(n_ref, 1)))
        pos += int(rng.integers(0, 400))
    return descs, keep, spans


def md_variants(md, ref_pos):
    """(position, kind, length) of the variants the reference's MD loop emits (blockjoin.c:1604-1672), restated
    character by character: the known set is laid over these positions so that any slip of the kernel's tokenizer
    (a number's value, a run's start, a position) changes votes"""
    def cls(ch):
        return 0 if ch.isdigit() else 1 if ch == "^" else 2
    out = []
    prev_t, prev_i = cls(md[0]), 0
    if prev_t == 2:
        out.append((ref_pos, "X", 1))
        ref_pos += 1
        prev_t = -1
    i = 1
    while i < len(md):
        t = cls(md[i])
        if t != prev_t:
            if prev_t == 0:
                ref_pos += int(md[prev_i:i]) & 0xffffffff
            elif prev_t == 1:
                if t == 0:
                    out.append((ref_pos, "D", i - prev_i - 1))
                    ref_pos += i - prev_i - 1
                    prev_t, prev_i = t, i
                i += 1
                continue
            if t == 2:
                out.append((ref_pos, "X", 1))
                ref_pos += 1
                prev_t, prev_i = -1, i
            else:
                prev_t, prev_i = t, i
        i += 1
    return out


def build_known(rng, descs, n, spans):
    """phased known variants: on the positions of the records' own variants (same or another length, random bases:
    ALT votes, failed comparisons, the deletion-reaches-the-next-variant rule), on random positions (REF votes), and
    runs of two and three on one position"""
    cand = []
    for i in range(n):
        d = descs[i]
        md = C.string_at(d.md, d.md_len).decode()
        for p, kind, ln in md_variants(md, d.pos):
            if rng.random() < 0.5:
                cand.append((p, {"X": 1, "D": 3}[kind], ln if rng.random() < 0.8 else ln + 1))
        ref = d.pos
        cig = np.ctypeslib.as_array(C.cast(d.cigar, C.POINTER(C.c_uint32)), shape=(d.n_cigar,))
        for c in cig:
            op, ln = int(c) & 15, int(c) >> 4
            if op in (0, 2, 3, 7, 8):
                ref += ln
            elif op == 1 and rng.random() < 0.7:
                cand.append((ref, 2, ln if rng.random() < 0.8 else ln + 1))
    lo, hi = spans[0][0], max(e for _, e in spans)
    for p in rng.integers(lo, hi + 50, size=max(8, len(cand) // 3)):
        cand.append((int(p), 1, 1))
    for k in rng.integers(0, len(cand), size=len(cand) // 8):   # equal positions
        cand.append((cand[k][0], 1, 1))
        if rng.random() < 0.3:
            cand.append((cand[k][0], 1, 1))
    cand.sort(key=lambda c: c[0])
    nk = len(cand)
    vars_ = (pb.Variant * nk)()
    bases = []
    for i, (p, op, ln) in enumerate(cand):
        v = vars_[i]
        ln = min(ln, 40)
        v.pos, v.len, v.op, v.haptag, v.bases_off = int(p), ln, op, int(rng.integers(0, 2)), len(bases)
        bases += list(rng.integers(0, 2, size=ln))   # two letters only: comparisons succeed often enough
    return vars_, nk, np.array(bases + [0] * 8, dtype=np.uint8)


def fuzz(gpu, seed, style, n):
    rng = np.random.default_rng(seed)
    descs, keep, spans = build_records(rng, n, style)
    vars_, nk, bases = build_known(rng, descs, n, spans)
    known = np.frombuffer(vars_, dtype=np.uint8, count=nk * C.sizeof(pb.Variant)).copy()
    starts = np.array([descs[i].pos for i in range(n)], dtype=np.uint32)
    port = ob.port_lib()
    kf = np.zeros(n, np.uint32)
    port.port_haptag_cursors(starts.ctypes.data, n, known.ctypes.data, nk, kf.ctypes.data)
    ctx = gpu.init([0])
    b = gpu.batch_begin(ctx)
    b.add_reads(C.addressof(descs), n)
    b.submit()
    b.haptag(known, nk, bases, kf)
    tags, status = b.collect_haptags()
    votes = np.zeros(2 * n, np.int32)
    gpu.lib.pomfret_gpu_debug_get_votes.argtypes = [C.c_void_p, C.c_void_p]
    assert gpu.lib.pomfret_gpu_debug_get_votes(b.h, votes.ctypes.data) == 0
    sz = C.sizeof(pb.ReadDesc)
    n_votes = 0
    for i in range(n):
        pv = (C.c_int * 2)()
        t = port.port_haptag_read(C.addressof(descs) + i * sz, known.ctypes.data, nk, bases.ctypes.data, int(kf[i]), pv)
        md = C.string_at(descs[i].md, descs[i].md_len).decode()
        assert status[i] == 0, (i, status[i], md)
        assert (t, list(pv)) == (tags[i], list(votes[2 * i:2 * i + 2])), (seed, style, i, t, list(pv), tags[i], votes[2 * i:2 * i + 2], md[:200])
        n_votes += pv[0] + pv[1]
    assert n_votes > n // 2
    b.end()
    gpu.destroy(ctx)
    del keep


CASES = [(1, "standard", 60), (2, "long_numbers", 60), (3, "weird", 80), (4, "short", 120), (5, "weird", 80)]


@pytest.mark.emu
@pytest.mark.parametrize("seed,style,n", CASES[:4])
def test_haptag_md_fuzz_emulated(built, seed, style, n):
    import build_emu
    fuzz(pb.load_gpu(build_emu.build()), seed, style, n)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,style,n", [(s + 10 * k, st, 4 * n) for k in range(3) for s, st, n in CASES])
def test_haptag_md_fuzz_gpu(built, seed, style, n):
    fuzz(pb.load_gpu(), seed, style, n)
