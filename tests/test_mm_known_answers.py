"""Hand-derived known-answer vectors for the MM/ML layer (SAMtags "Base modifications": `C+m,d0,d1,..` = skip d0
unmodified cytosines of the ORIGINAL read strand, the next one carries the call, skip d1, ...; several codes on
one list interleave their ML bytes per position; several lists concatenate them; reversed alignments are counted
from the right end of SEQ on the complemented base) combined with pomfret's own rules on top (CpG filter and
categories blockjoin.c:846-882, CIGAR walk 605-792, lo=100 / hi=156).  The first vector uses the specification's
example read AGCTCTCCAGAGTCGNACGCCATYCGCGCGCCACCA with its `C+m,2,2,1,4,1` list (modified cytosines at read offsets
6, 17, 20, 31, 34).  Every expected value below was worked out by hand from those rules, not produced by any of
the implementations; the same vectors go through
  * the unmodified reference's fill_read_meth_record_from_bam_line on the hts shim (oracle/_ref),
  * the oracle port,
  * the decode kernel (emulated on the CPU, real on the GPU).
This ties the three restatements of htslib's base-modification iterator to an outside answer."""
import ctypes as C
import os

import numpy as np
import pytest

import decode_fuzz
import oracle_bindings as ob
import pomfret_b200 as pb
from pomfret_b200 import _ffi

SPEC = "AGCTCTCCAGAGTCGNACGCCATYCGCGCGCCACCA"
SPEC_RC = SPEC[::-1].translate(str.maketrans("ACGTNY", "TGCANR"))
TOY = "ACGTCGACCGTACGCGTTCA"  # C at 1,4,7,8,12,14,18; CpG cytosines 1,4,8,12,14; G at 2,5,9,13,15

# name, pos, flag, seq, cigar [(op,len)], MM, ML, MN, expected positions, expected categories (0 meth, 1 unmeth, 2 no call)
VECTORS = [
    # listed 6,17,20,31,34; only 17 is followed by G (ML 128 -> no call); the non-CpG entries switch the reference's
    # implicit mode on: every other CpG of SEQ (13, 24, 26, 28) becomes an unmethylated call
    ("spec_example_fwd", 100, 0, SPEC, [(0, 36)], "C+m,2,2,1,4,1;", [102, 128, 153, 179, 204], -1,
     [113, 117, 124, 126, 128], [1, 2, 1, 1, 1]),
    # the specification's three-list form: a ChEBI list and an any-base list behind the C+m list shift nothing
    ("spec_example_three_lists", 100, 0, SPEC, [(0, 36)], "C+m,2,2,1,4,1;C+76792,6,7;N+n,15,2;",
     [102, 128, 153, 179, 204, 161, 187, 212, 169], -1, [113, 117, 124, 126, 128], [1, 2, 1, 1, 1]),
    # the same molecule aligned to the other strand: SEQ is the reverse complement, the list is unchanged; the CpG of
    # read offsets 17/18 shows as C,G at SEQ 17,18 and is reported at the C (cgoffset -1)
    ("spec_example_rev", 100, 16, SPEC_RC, [(0, 36)], "C+m,2,2,1,4,1;", [102, 128, 153, 179, 204], -1,
     [106, 108, 110, 117, 121], [1, 1, 1, 2, 1]),
    # two codes on one list: ML holds (h, m) per listed base
    ("multicode_hm", 1000, 0, TOY, [(0, 20)], "C+hm?,0,0,1,0,0;", [10, 200, 20, 50, 30, 120, 40, 156, 50, 99], -1,
     [1001, 1004, 1008, 1012, 1014], [0, 1, 2, 0, 1]),
    # ... or (m, h); '.' instead of '?' changes nothing here
    ("multicode_mh", 1000, 0, TOY, [(0, 20)], "C+mh.,0,0,1,0,0;", [200, 10, 50, 20, 120, 30, 156, 40, 99, 50], -1,
     [1001, 1004, 1008, 1012, 1014], [0, 1, 2, 0, 1]),
    # two lists: the ML bytes of the first list (2 entries) come first
    ("two_lists", 1000, 0, TOY, [(0, 20)], "C+h?,1,3;C+m?,0,0,1,0,0;", [7, 8, 200, 50, 120, 156, 99], -1,
     [1001, 1004, 1008, 1012, 1014], [0, 1, 2, 0, 1]),
    # reversed alignment: cytosines of the read are the G of SEQ counted from the right: 15, (13 skipped), 9, 5
    ("reverse_toy", 1000, 16, TOY, [(0, 20)], "C+m?,0,1,0;", [210, 30, 130], -1, [1004, 1008, 1014], [2, 1, 0]),
    # MN disagrees with the SEQ length: the tags are stale, the record carries no modification
    ("mn_mismatch", 1000, 0, TOY, [(0, 20)], "C+m?,0,0,1,0,0;", [200, 50, 120, 156, 99], 19, [], []),
    # 3S 5M 2I 4M 3D 6M: offset 1 lies in the clip (dropped); 4 -> 501; 8 is the first inserted base but the
    # inclusive trigger of the preceding M handles it (blockjoin.c:663-665) -> 505; 12 -> 507; 14 is again taken by
    # the inclusive trigger of the 4M (before the deletion is applied) -> 509
    ("clip_ins_del", 500, 0, TOY, [(4, 3), (0, 5), (1, 2), (0, 4), (2, 3), (0, 6)], "C+m?,0,0,1,0,0;",
     [200, 50, 120, 156, 99], -1, [501, 505, 507, 509], [1, 2, 0, 1]),
    # the first base of SEQ can be listed but never called (0 < pos < len-1, blockjoin.c:846); it is not "implicit"
    ("first_base_listed", 10, 0, "CGACG", [(0, 5)], "C+m,0,0;", [250, 10], -1, [13], [1]),
]


def _records():
    decode_fuzz.NT16.update({"Y": 10, "R": 5})
    R = decode_fuzz.Records()
    for name, pos, flag, seq, cigar, mm, ml, mn, _, _ in VECTORS:
        R.add(name, pos, flag, seq, cigar, mm, ml, mn=mn, malformed=1 if (mn >= 0 and mn != len(seq)) else 0)
    return R


def test_known_answers_oracle_port(built):
    R = _records()
    arr = R.array()
    port = ob.port_lib()
    for i, v in enumerate(VECTORS):
        out = ob.PortCalls()
        st, end = C.c_uint32(), C.c_uint32()
        port.port_decode_read(C.addressof(arr) + i * C.sizeof(_ffi.ReadDesc), 100, 156, C.byref(out), C.byref(st), C.byref(end))
        pos = list(np.ctypeslib.as_array(out.pos, shape=(out.n,))) if out.n else []
        cat = list(np.ctypeslib.as_array(out.cat, shape=(out.n,))) if out.n else []
        assert (pos, cat) == (v[8], v[9]), v[0]
        assert bool(st.value & 1) == bool(v[8]), v[0]


@pytest.mark.skipif(not os.path.exists(ob.REF_SO), reason="oracle/_ref not built")
def test_known_answers_reference_on_shim(built):
    """the unmodified reference function on top of the shim's bam_parse_basemod / bam_mods_at_next_pos"""
    ref = ob.ref_lib()
    f = ref.refh_fill_read_meth
    f.restype = C.c_int
    f.argtypes = [C.c_uint32, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_int, C.c_int,
                  C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    decode_fuzz.NT16.update({"Y": 10, "R": 5})
    for name, pos, flag, seq, cigar, mm, ml, mn, epos, ecat in VECTORS:
        cg = np.array([(l << 4) | op for op, l in cigar], dtype=np.uint32)
        sq = decode_fuzz.pack_seq(seq)
        mlb = np.array(ml, dtype=np.uint8)
        opos, ocat = np.zeros(64, np.uint32), np.zeros(64, np.uint8)
        n, imp = C.c_int(), C.c_int()
        with ob.quiet_reference():
            stat = f(pos, flag, cg.ctypes.data, len(cg), sq.ctypes.data, len(seq), mm.encode(), mlb.ctypes.data, len(ml), mn,
                     100, 156, opos.ctypes.data, ocat.ctypes.data, 64, C.byref(n), C.byref(imp))
        got = (list(opos[:n.value]), list(ocat[:n.value])) if stat > 0 else ([], [])
        assert got == (epos, ecat), name


def _check_kernel(gpu):
    R = _records()
    assert not decode_fuzz.check_against_port(gpu, R)  # and, independently of the port, against the literal answers:
    arr = R.array()
    ctx = gpu.init()
    b = gpu.batch_begin(ctx)
    b.add_reads(arr, len(VECTORS))
    b.submit()
    b.decode(100, 156)
    for i, v in enumerate(VECTORS):
        st, nc, end = b.read_info(i)
        pos, cat = b.calls(i) if st & 1 else ([], [])
        assert (list(pos), list(cat)) == (v[8], v[9]), v[0]
    b.end()
    gpu.destroy(ctx)


@pytest.mark.emu
def test_known_answers_kernel_emulated(built):
    import build_emu
    _check_kernel(pb.load_gpu(build_emu.build()))


@pytest.mark.gpu
def test_known_answers_kernel_gpu(built):
    _check_kernel(pb.load_gpu())
