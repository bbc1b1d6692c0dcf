"""Pin the oracle: the plain-C port (oracle/port) against the compiled, unmodified reference
(oracle/_ref, built from /root/reference) on seeded synthetic windows, plus the SURVEY.md App. B
known-answer vectors that were produced from the reference code."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_bindings as ob
import parity
import pomfret_b200 as pb

needs_ref = pytest.mark.skipif(not os.path.exists(ob.REF_SO), reason="oracle/_ref not built (no /root/reference)")


def _check(data, cov, readlen=15000, **kw):
    host = pb.load_host()
    hb = host.bam_open(data["bam"])
    cfg = ob.make_config(cov, readlen=readlen, **kw)
    n_ok = 0
    for w, n, chrom, s, e in parity.load_windows(host, hb, data["gaps"], cfg):
        r = ob.ref_window(data["bam"], chrom, s, e, cfg)
        p = ob.port_window(host.window_descs(w), n, s, e, cfg)
        bad = parity.diff_dicts(r, p)
        assert not bad, (chrom, s, e, bad)
        host.window_free(w)
        n_ok += 1
    host.bam_close(hb)
    return n_ok


@needs_ref
def test_port_matches_reference_30x(synth30):
    assert _check(synth30, 30) >= 2


@needs_ref
def test_port_matches_reference_short_reads(synth_small):
    assert _check(synth_small, 36, readlen=2000) >= 1


@needs_ref
def test_port_matches_reference_implicit_calls(synth_implicit):
    # non-CpG C+m entries switch the reference into its implicit-canonical mode
    assert _check(synth_implicit, 34, readlen=1500) >= 1


@needs_ref
def test_port_matches_reference_sparse_lists(synth_sparse_implicit):
    assert _check(synth_sparse_implicit, 34, readlen=1500) >= 1


@needs_ref
def test_port_matches_reference_other_parameters(synth_small):
    assert _check(synth_small, 20, readlen=2000, k=2, k_span=800) >= 1
    assert _check(synth_small, 50, readlen=2000, k=4, lo=80, hi=180) >= 1


@needs_ref
def test_whole_function_equals_replayed_driver(synth_small):
    """refwin_run replays haplotag_region_given_bam call by call; check it against the function itself."""
    lib = ob.ref_lib()
    cfg = ob.make_config(36, readlen=2000)
    for chrom, s, e, _ in synth_small["gaps"]:
        r = ob.ref_window(synth_small["bam"], chrom, s, e, cfg)
        tags = np.zeros(8192, dtype=np.uint8)
        n = C.c_int()
        with ob.quiet_reference():
            dec = lib.refwin_run_whole(synth_small["bam"].encode(), chrom.encode(), s, e, cfg.k, cfg.k_span, cfg.lo,
                                       cfg.hi, cfg.cov_known, cfg.cov_for_selection, cfg.cov_for_runtime,
                                       cfg.readlen_threshold, cfg.min_mapq, cfg.n_candidates_per_iter, None,
                                       tags.ctypes.data, len(tags), C.byref(n))
        assert dec == r["decision"] and n.value == r["n_reads"]
        assert np.array_equal(tags[:n.value], r["tags_final"])


OPS = {'M': 0, 'I': 1, 'D': 2, 'N': 3, 'S': 4, 'H': 5}
B1 = [  # SURVEY.md App. B.1 (qs = 1000)
    ([(10, 'M'), (2, 'D'), (10, 'M')], 0, [3, 9, 10, 15], [0, 1, 0, 1], [(1003, 0), (1009, 1), (1010, 0), (1017, 1)]),
    ([(10, 'M'), (2, 'I'), (10, 'M')], 0, [3, 10, 11, 12, 15], [0, 1, 0, 1, 0], [(1003, 0), (1010, 1), (1013, 0)]),
    ([(5, 'S'), (10, 'M')], 0, [1, 3], [0, 1], [(998, 1)]),
    ([(5, 'S'), (10, 'M'), (3, 'S')], 1, [1, 5, 8, 14, 15, 16], [0, 1, 0, 1, 0, 1],
     [(999, 1), (1002, 0), (1008, 1), (1009, 0)]),
]


def _port_map(cig, strand, poss, cats):
    lib = ob.port_lib()
    c = np.array([(l << 4) | OPS[o] for l, o in cig], dtype=np.uint32)
    p = np.array(poss, dtype=np.uint32)
    q = np.array(cats, dtype=np.uint8)
    out = ob.PortCalls()
    rc = lib.port_map_mods_to_ref(c.ctypes.data, len(c), 1000, strand, p.ctypes.data, q.ctypes.data, len(p), None, 100,
                                  C.byref(out))
    res = [(int(out.pos[i]), int(out.cat[i])) for i in range(out.n)]
    return rc, res


@pytest.mark.parametrize("cig,strand,poss,cats,want", B1)
def test_appendix_b1_port(built, cig, strand, poss, cats, want):
    rc, got = _port_map(cig, strand, poss, cats)
    assert rc == 1 and got == want


def test_appendix_b1_fatal_cigar(built):
    rc, _ = _port_map([(5, 'H'), (10, 'M')], 0, [1], [0])
    assert rc == -6  # the reference exits: "fatal: unknown cigar operation (value=5)"


@needs_ref
@pytest.mark.parametrize("cig,strand,poss,cats,want", B1)
def test_appendix_b1_reference(built, cig, strand, poss, cats, want):
    lib = ob.ref_lib()
    c = np.array([(l << 4) | OPS[o] for l, o in cig], dtype=np.uint32)
    p = np.array(poss, dtype=np.uint32)
    q = np.array(cats, dtype=np.uint8)
    op = np.zeros(64, dtype=np.uint32)
    oc = np.zeros(64, dtype=np.uint8)
    n = C.c_int()
    with ob.quiet_reference():
        rc = lib.refh_get_mod_poss_on_ref(c.ctypes.data, len(c), 1000, strand, p.ctypes.data, q.ctypes.data, len(p), None,
                                          100, op.ctypes.data, oc.ctypes.data, 64, C.byref(n))
    assert rc == 1 and [(int(op[i]), int(oc[i])) for i in range(n.value)] == want


def test_fisher_against_scipy(built):
    """Only `p < 0.001` is consumed (blockjoin.c:3928); still keep the value itself tight."""
    from scipy.stats import fisher_exact
    lib = ob.port_lib()
    rng = np.random.default_rng(3)
    for _ in range(300):
        a, b, c, d = (int(x) for x in rng.integers(0, 40, size=4))
        if a + b == 0 or c + d == 0 or a + c == 0 or b + d == 0:
            continue
        want = fisher_exact([[a, b], [c, d]])[1]
        got = lib.port_fisher_two_sided(a, b, c, d)
        assert abs(got - want) <= 1e-9 + 1e-7 * want, (a, b, c, d, got, want)
        if os.path.exists(ob.REF_SO):
            shim = ob.ref_lib().refh_fisher_two_sided(a, b, c, d)
            assert (shim < 0.001) == (want < 0.001)
            assert abs(shim - want) <= 1e-9 + 1e-7 * want
