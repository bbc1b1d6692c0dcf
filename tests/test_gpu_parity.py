"""Parity tests proper: the CUDA path, called through the C ABI (libpomfret_gpu.so, nvcc sm_100a build),
against the oracle on the same seeded inputs.  Integer outputs, tags, decisions and the fp32 join scores
must be bit-exact."""
import os

import numpy as np
import pytest

import conftest
import oracle_bindings as ob
import parity
import pomfret_b200 as pb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(built):
    g = pb.load_gpu()  # raises if the CUDA library is missing: no fallback
    assert g.device_count() >= 1, "no CUDA device"
    return g


def _run(gpu, data, cov, readlen=15000, max_windows=None, check_ref=True, **kw):
    host = pb.load_host()
    hb = host.bam_open(data["bam"])
    cfg = pb.make_config(cov, readlen=readlen, **kw)
    ocfg = ob.make_config(cov, readlen=readlen, **kw)
    gaps = data["gaps"][:max_windows] if max_windows else data["gaps"]
    wins = parity.load_windows(host, hb, gaps, cfg)
    ctx = gpu.init([0])
    b, layout, res, tags, ids, rc = parity.run_gpu_batch(gpu, ctx, host, wins, cfg)
    assert rc == 0, gpu.strerror(rc)
    decisions = []
    for wi, ((w, n, chrom, s, e), (first, _)) in enumerate(zip(wins, layout)):
        p = ob.port_window(host.window_descs(w), n, s, e, ocfg)
        bad = parity.compare_window(b, wi, first, n, res, tags, ids, p)
        assert not bad, (chrom, s, e, bad[:10])
        if check_ref and os.path.exists(ob.REF_SO):
            r = ob.ref_window(data["bam"], chrom, s, e, ocfg)
            assert res[wi].decision == r["decision"]
            kept = [i for i in range(n) if ids[first + i] >= 0][:r["n_reads"]]
            assert np.array_equal(np.array([tags[first + i] for i in kept], dtype=np.uint8), r["tags_final"])
        decisions.append(res[wi].decision)
        host.window_free(w)
    b.end()
    gpu.destroy(ctx)
    host.bam_close(hb)
    return decisions


def test_gpu_matches_oracle_30x(gpu, synth30):
    d = _run(gpu, synth30, 30)
    assert len(d) >= 2


def test_gpu_matches_oracle_60x(gpu, synth60):
    # BASELINE.json config-3 depth (-c 60: cov_for_selection 7, runtime 14, 16 candidates, up to ~900 reads per window)
    d = _run(gpu, synth60, 60)
    assert sorted(set(d)) == [-1, 0, 1]


def test_gpu_matches_oracle_config1_quickstart(gpu, synth_config1):
    d = _run(gpu, synth_config1, 60)
    assert d == [1]  # the bundled example's own answer: the second phase set is flipped ("trans")


def test_gpu_matches_oracle_wgs5(gpu, synth_wgs5):
    d = _run(gpu, synth_wgs5, 30)
    assert len(d) == 7 and {c for c, _, _, _ in synth_wgs5["gaps"]} == {"chr1", "chr2", "chr20", "chrX"}


def test_gpu_matches_oracle_long_cigar(gpu, synth_long_cigar):
    # records with more than 65535 CIGAR operations (restored from the CG tag by the loader)
    host = pb.load_host()
    hb = host.bam_open(synth_long_cigar["bam"])
    import ctypes as C
    from pomfret_b200 import _ffi
    c, s, e, _ = synth_long_cigar["gaps"][0]
    w = host.window_load(hb, c, s, e, 15000, 10)
    d = C.cast(host.window_descs(w), C.POINTER(_ffi.ReadDesc))
    assert max(d[i].n_cigar for i in range(host.window_n(w))) > 65535
    host.window_free(w)
    host.bam_close(hb)
    _run(gpu, synth_long_cigar, 30)


def test_gpu_matches_oracle_short_reads(gpu, synth_small):
    _run(gpu, synth_small, 36, readlen=2000)


def test_gpu_matches_oracle_implicit(gpu, synth_implicit):
    _run(gpu, synth_implicit, 34, readlen=1500)


def test_gpu_matches_oracle_call_slot_overflow(gpu, synth_sparse_implicit):
    _run(gpu, synth_sparse_implicit, 34, readlen=1500)


@pytest.mark.parametrize("which", ["implicit", "sparse"])
def test_gpu_matches_oracle_implicit_streaming_path(gpu, synth_implicit, synth_sparse_implicit, which, monkeypatch):
    # implicit canonical calls through the streaming path (lean path off): what records over 65 535 bases take
    monkeypatch.setenv("POMFRET_GPU_DECODE_LEAN", "0")
    _run(gpu, synth_implicit if which == "implicit" else synth_sparse_implicit, 34, readlen=1500)


def test_gpu_implicit_long_reads(gpu, built, tmp_path):
    # 20 kb reads (some beyond the lean path's limits) with non-CpG entries: lean + streaming implicit walks on real sizes
    data = conftest.run_synth(str(tmp_path / "impl"), ["-c", "30", "-s", "47", "-C", "chr20:64444167:9000000-9400000", "--implicit", "0.05",
                                                        "--listed", "0.7"])
    _run(gpu, data, 30)


@pytest.mark.parametrize("kw", [dict(k=2, k_span=800), dict(k=4, lo=80, hi=180), dict(k=1), dict(k_span=300), dict(k=5, k_span=3000),
                                dict(k=6), dict(k=8, k_span=9000)])
def test_gpu_matches_oracle_parameters(gpu, synth_small, kw):
    _run(gpu, synth_small, 30, readlen=2000, check_ref=False, **kw)


def test_gpu_coverage_gate_and_empty(gpu, synth_small):
    """cov 200 => cov_for_selection 20: no site qualifies (window skipped, blockjoin.c:4266-4270);
    a window without records => abandoned by the left-coverage gate (blockjoin.c:1161-1163)."""
    _run(gpu, synth_small, 200, readlen=2000, check_ref=False)
    g = gpu
    ctx = g.init([0])
    b = g.batch_begin(ctx)
    b.add_window(1000, 2000, 0, 0)
    cfg = pb.make_config(30)
    b.submit(); b.decode(cfg.lo, cfg.hi); b.pileup(cfg); b.join(cfg)
    res, tags, ids, rc = b.collect()
    assert rc == 0 and res[0].decision == -1 and res[0].n_reads == 0
    b.end()
    g.destroy(ctx)


def test_gpu_batch_of_many_windows_is_order_independent(gpu, synth30):
    """Windows are independent units: a batch with the windows in reverse order gives the same answers."""
    host = pb.load_host()
    hb = host.bam_open(synth30["bam"])
    cfg = pb.make_config(30)
    wins = parity.load_windows(host, hb, synth30["gaps"], cfg)
    ctx = gpu.init([0])
    b1, l1, r1, t1, i1, rc1 = parity.run_gpu_batch(gpu, ctx, host, wins, cfg)
    b2, l2, r2, t2, i2, rc2 = parity.run_gpu_batch(gpu, ctx, host, wins[::-1], cfg)
    assert rc1 == 0 and rc2 == 0
    nw = len(wins)
    for wi in range(nw):
        a, z = r1[wi], r2[nw - 1 - wi]
        assert (a.decision, a.join_fwd, a.join_bwd, a.n_reads, a.n_sites_fwd, list(a.table_fwd), list(a.table_bwd)) == \
               (z.decision, z.join_fwd, z.join_bwd, z.n_reads, z.n_sites_fwd, list(z.table_fwd), list(z.table_bwd))
        f1, n1 = l1[wi]
        f2, n2 = l2[nw - 1 - wi]
        assert np.array_equal(t1[f1:f1 + n1], t2[f2:f2 + n2])
    b1.end(); b2.end()
    gpu.destroy(ctx)
    host.bam_close(hb)


def _run_tweaked(gpu, data, cov, readlen, tweak, max_windows=None):
    host = pb.load_host()
    hb = host.bam_open(data["bam"])
    cfg, ocfg = pb.make_config(cov, readlen=readlen), ob.make_config(cov, readlen=readlen)
    tweak(cfg); tweak(ocfg)
    gaps = data["gaps"][:max_windows] if max_windows else data["gaps"]
    wins = parity.load_windows(host, hb, gaps, cfg)
    ctx = gpu.init([0])
    b, layout, res, tags, ids, rc = parity.run_gpu_batch(gpu, ctx, host, wins, cfg)
    assert rc == 0, gpu.strerror(rc)
    for wi, ((w, n, chrom, s, e), (first, _)) in enumerate(zip(wins, layout)):
        p = ob.port_window(host.window_descs(w), n, s, e, ocfg)
        bad = parity.compare_window(b, wi, first, n, res, tags, ids, p, deep=False)
        assert not bad, (chrom, s, e, bad[:10])
        host.window_free(w)
    b.end()
    gpu.destroy(ctx)
    host.bam_close(hb)


@pytest.mark.parametrize("n_cand", [1, 3, 15, 40, 128, 300])
def test_gpu_join_candidate_counts(gpu, synth30, synth_small, n_cand):
    # fewer candidate slots than warps, more slots than warps, more slots than lanes
    def tweak(cfg):
        cfg.n_candidates_per_iter = n_cand
    _run_tweaked(gpu, synth_small, 36, 2000, tweak)
    _run_tweaked(gpu, synth30, 30, 15000, tweak, max_windows=2)


@pytest.mark.parametrize("mode", ["0", "half", "u16"])
def test_gpu_join_global_memory_paths(gpu, synth30, mode, monkeypatch):
    # windows too large for shared memory keep count tables / per-read state in global memory
    monkeypatch.setenv("POMFRET_GPU_JOIN_SMEM", mode)
    _run_tweaked(gpu, synth30, 30, 15000, lambda cfg: None)


def test_gpu_gather_from_registered_buffers(gpu, synth30):
    # registered loader buffers: no host copy, the gather kernel lays out the blob (unaligned sources, odd lengths)
    host = pb.load_host()
    hb = host.bam_open(synth30["bam"])
    cfg, ocfg = pb.make_config(30), ob.make_config(30)
    wins = parity.load_windows(host, hb, synth30["gaps"][:2], cfg)
    ctx = gpu.init([0])
    regs = []
    for w, n, chrom, s, e in wins:
        ptr, nbytes = host.window_arena(w)
        if nbytes:
            gpu.host_register(ctx, ptr, nbytes)
            regs.append(ptr)
    b, layout, res, tags, ids, rc = parity.run_gpu_batch(gpu, ctx, host, wins, cfg)
    assert rc == 0, gpu.strerror(rc)
    payload = sum(host.window_arena(w)[1] for w, _, _, _, _ in wins)
    assert b.timing().bytes_h2d < 1.1 * payload
    for wi, ((w, n, chrom, s, e), (first, _)) in enumerate(zip(wins, layout)):
        p = ob.port_window(host.window_descs(w), n, s, e, ocfg)
        assert not parity.compare_window(b, wi, first, n, res, tags, ids, p), (chrom, s, e)
    b.end()
    for ptr in regs:
        gpu.host_unregister(ctx, ptr)
    gpu.destroy(ctx)
    host.bam_close(hb)


def test_gpu_mixed_registered_and_unregistered_records(gpu, synth30):
    """One batch whose records come partly from registered buffers (gathered by the device) and partly from plain
    memory (copied by the host), in both orders, and again after a reset that follows a smaller batch: the pinned
    arena then holds fewer bytes than the laid-out blob (PinBuf::reserve copied `len` bytes)."""
    host = pb.load_host()
    hb = host.bam_open(synth30["bam"])
    cfg, ocfg = pb.make_config(30), ob.make_config(30)
    wins = parity.load_windows(host, hb, synth30["gaps"][:3], cfg)
    ctx = gpu.init([0])
    b = gpu.batch_begin(ctx)
    for registered in ([0], [1, 2], [0, 2]):
        regs = []
        for i in registered:
            ptr, nbytes = host.window_arena(wins[i][0])
            if nbytes:
                gpu.host_register(ctx, ptr, nbytes)
                regs.append(ptr)
        b.reset()
        layout = []
        for w, n, chrom, s, e in wins:
            first = b.add_reads(host.window_descs(w), n)
            b.add_window(s, e, first, n)
            layout.append((first, n))
        b.submit(); b.decode(cfg.lo, cfg.hi); b.pileup(cfg); b.join(cfg)
        res, tags, ids, rc = b.collect(check=False)
        assert rc == 0, gpu.strerror(rc)
        for wi, ((w, n, chrom, s, e), (first, _)) in enumerate(zip(wins, layout)):
            p = ob.port_window(host.window_descs(w), n, s, e, ocfg)
            assert not parity.compare_window(b, wi, first, n, res, tags, ids, p, deep=False), (registered, chrom, s, e)
        for ptr in regs:
            gpu.host_unregister(ctx, ptr)
    b.end()
    gpu.destroy(ctx)
    host.bam_close(hb)


def test_gpu_shared_records_decode_once(gpu, synth60, synth_sparse_implicit):
    """decode-once entry point: overlapping windows whose common records are staged and
    decoded once; every window still equals the oracle's answer for it; the call-slot overflow re-run keeps the sharing"""
    host = pb.load_host()
    for data, cov, readlen, shift in ((synth60, 60, 15000, 30000), (synth_sparse_implicit, 34, 1500, 2000)):
        hb = host.bam_open(data["bam"])
        cfg, ocfg = pb.make_config(cov, readlen=readlen), ob.make_config(cov, readlen=readlen)
        gaps = []
        for c, s, e, t in data["gaps"][:2]:
            gaps += [(c, s, e, t), (c, s + shift, e + shift, t), (c, s + 2 * shift, e + 2 * shift, t)]
        wins = parity.load_windows(host, hb, gaps, cfg)
        ctx = gpu.init([0])
        b, layout, res, tags, ids, rc, n_shared = parity.run_gpu_batch_shared(gpu, ctx, host, wins, cfg)
        assert rc == 0, gpu.strerror(rc)
        assert n_shared > sum(n for _, n, _, _, _ in wins) // 3
        b0, layout0, res0, tags0, ids0, rc0 = parity.run_gpu_batch(gpu, ctx, host, wins, cfg)
        assert b.timing().decode_bytes < 0.75 * b0.timing().decode_bytes  # shared records are decoded once
        for wi, ((w, n, chrom, s, e), (first, _)) in enumerate(zip(wins, layout)):
            p = ob.port_window(host.window_descs(w), n, s, e, ocfg)
            bad = parity.compare_window(b, wi, first, n, res, tags, ids, p)
            assert not bad, (chrom, s, e, bad[:10])
            host.window_free(w)
        b.end(); b0.end()
        gpu.destroy(ctx)
        host.bam_close(hb)


def test_gpu_long_reads_uncached_keys(gpu, built, tmp_path):
    # reads that span more than 256 methmer sites: their keys are scored straight from the pool
    import conftest
    data = conftest.run_synth(str(tmp_path / "long"), ["-c", "32", "-s", "41", "-C", "chrL:900000:0-600000", "--readlen", "70000",
                                                        "--block", "250000", "--gap", "20000-30000"])
    _run(gpu, data, 32, check_ref=False)
