"""Step the real kernel sources (pomfret_b200/csrc/gpu/*.cuh, unmodified) on the CPU SIMT emulator
and compare every stage with the oracle port.  This is developer/CI coverage of kernel *logic* for
boxes without a GPU; the product library is the nvcc build and is exercised by the `gpu` tests."""
import pytest

import oracle_bindings as ob
import parity
import pomfret_b200 as pb

pytestmark = pytest.mark.emu


@pytest.fixture(scope="module")
def emu_gpu(built):
    import build_emu
    return pb.load_gpu(build_emu.build())


def _run(emu_gpu, data, cov, readlen, max_windows=1, **kw):
    host = pb.load_host()
    hb = host.bam_open(data["bam"])
    cfg = pb.make_config(cov, readlen=readlen, **kw)
    ocfg = ob.make_config(cov, readlen=readlen, **kw)
    wins = parity.load_windows(host, hb, data["gaps"][:max_windows], cfg)
    ctx = emu_gpu.init()
    b, layout, res, tags, ids, rc = parity.run_gpu_batch(emu_gpu, ctx, host, wins, cfg)
    assert rc == 0
    for wi, ((w, n, chrom, s, e), (first, _)) in enumerate(zip(wins, layout)):
        p = ob.port_window(host.window_descs(w), n, s, e, ocfg)
        bad = parity.compare_window(b, wi, first, n, res, tags, ids, p)
        assert not bad, (chrom, s, e, bad[:10])
        host.window_free(w)
    b.end()
    emu_gpu.destroy(ctx)
    host.bam_close(hb)


def test_emulated_kernels_match_oracle(emu_gpu, synth_small):
    _run(emu_gpu, synth_small, 36, 2000)


def _overlapping_windows(data, shift):
    """the gaps of the data set plus shifted copies: neighbouring windows then share most of their records"""
    out = []
    for c, s, e, t in data["gaps"]:
        out += [(c, s, e, t), (c, s + shift, e + shift, t)]
    return out


def test_emulated_shared_records_decode_once(emu_gpu, synth_small):
    # decode-once entry point: overlapping windows, shared slots, same answers as the oracle
    host = pb.load_host()
    hb = host.bam_open(synth_small["bam"])
    cfg, ocfg = pb.make_config(36, readlen=2000), ob.make_config(36, readlen=2000)
    wins = parity.load_windows(host, hb, _overlapping_windows(synth_small, 3000)[:2], cfg)
    ctx = emu_gpu.init()
    b, layout, res, tags, ids, rc, n_shared = parity.run_gpu_batch_shared(emu_gpu, ctx, host, wins, cfg)
    assert rc == 0 and n_shared > 100
    for wi, ((w, n, chrom, s, e), (first, _)) in enumerate(zip(wins, layout)):
        p = ob.port_window(host.window_descs(w), n, s, e, ocfg)
        bad = parity.compare_window(b, wi, first, n, res, tags, ids, p)
        assert not bad, (chrom, s, e, bad[:10])
        host.window_free(w)
    b.end()
    emu_gpu.destroy(ctx)
    host.bam_close(hb)


def test_emulated_kernels_implicit_mode(emu_gpu, synth_implicit):
    _run(emu_gpu, synth_implicit, 34, 1500)


def test_emulated_kernels_k2(emu_gpu, synth_small):
    _run(emu_gpu, synth_small, 24, 2000, k=2, k_span=900)


def test_emulated_kernels_k5(emu_gpu, synth_small):
    # methmers of five symbols: 3^5 + 1 entries per site (the round-1 engine stopped at k = 4)
    _run(emu_gpu, synth_small, 24, 2000, k=5, k_span=2500)


def test_emulated_kernels_call_slot_overflow(emu_gpu, synth_sparse_implicit):
    _run(emu_gpu, synth_sparse_implicit, 34, 1500)


@pytest.mark.emu
def test_emulated_kernels_implicit_mode_streaming_path(emu_gpu, synth_sparse_implicit, monkeypatch):
    # the same with the lean path switched off: the streaming path squeezes the non-CpG holes out of its kept mods and
    # runs the implicit-call walk (what records over 65 535 bases take)
    monkeypatch.setenv("POMFRET_GPU_DECODE_LEAN", "0")
    _run(emu_gpu, synth_sparse_implicit, 34, 1500)


def _run_cfg(emu_gpu, data, cov, readlen, tweak):
    """like _run, with the engine / oracle configurations adjusted by tweak(cfg)"""
    host = pb.load_host()
    hb = host.bam_open(data["bam"])
    cfg, ocfg = pb.make_config(cov, readlen=readlen), ob.make_config(cov, readlen=readlen)
    tweak(cfg); tweak(ocfg)
    wins = parity.load_windows(host, hb, data["gaps"][:1], cfg)
    ctx = emu_gpu.init()
    b, layout, res, tags, ids, rc = parity.run_gpu_batch(emu_gpu, ctx, host, wins, cfg)
    assert rc == 0
    for wi, ((w, n, chrom, s, e), (first, _)) in enumerate(zip(wins, layout)):
        p = ob.port_window(host.window_descs(w), n, s, e, ocfg)
        bad = parity.compare_window(b, wi, first, n, res, tags, ids, p, deep=False)
        assert not bad, (chrom, s, e, bad[:10])
        host.window_free(w)
    b.end()
    emu_gpu.destroy(ctx)
    host.bam_close(hb)


@pytest.mark.parametrize("n_cand", [1, 40, 200])
def test_emulated_join_candidate_counts(emu_gpu, synth_small, n_cand):
    # fewer slots than warps, more slots than warps, more slots than lanes
    def tweak(cfg):
        cfg.n_candidates_per_iter = n_cand
    _run_cfg(emu_gpu, synth_small, 36, 2000, tweak)


@pytest.mark.parametrize("mode", ["0", "half"])
def test_emulated_join_global_memory_paths(emu_gpu, synth_small, mode, monkeypatch):
    # windows too large for shared memory keep tables / per-read state in global memory
    monkeypatch.setenv("POMFRET_GPU_JOIN_SMEM", mode)
    _run_cfg(emu_gpu, synth_small, 36, 2000, lambda cfg: None)


def test_emulated_rewind_reruns_with_other_parameters(emu_gpu, synth_small):
    # records stay resident: second pass with other thresholds and k must match the oracle for those
    host = pb.load_host()
    hb = host.bam_open(synth_small["bam"])
    cfg = pb.make_config(36, readlen=2000)
    wins = parity.load_windows(host, hb, synth_small["gaps"][:1], cfg)
    ctx = emu_gpu.init()
    b, layout, res, tags, ids, rc = parity.run_gpu_batch(emu_gpu, ctx, host, wins, cfg)
    assert rc == 0
    kw = dict(k=2, k_span=900, lo=80, hi=180)
    cfg2, ocfg2 = pb.make_config(24, readlen=2000, **kw), ob.make_config(24, readlen=2000, **kw)
    b.rewind()
    b.decode(cfg2.lo, cfg2.hi)
    b.pileup(cfg2)
    b.join(cfg2)
    res, tags, ids, rc = b.collect(check=False)
    assert rc == 0
    (w, n, chrom, s, e), (first, _) = wins[0], layout[0]
    p = ob.port_window(host.window_descs(w), n, s, e, ocfg2)
    assert not parity.compare_window(b, 0, first, n, res, tags, ids, p)
    b.end()
    emu_gpu.destroy(ctx)
    host.bam_close(hb)


def test_emulated_gather_from_registered_buffers(emu_gpu, synth_small):
    # the loader's record buffer is registered: add_reads() copies nothing, the gather kernel lays the blob out
    host = pb.load_host()
    hb = host.bam_open(synth_small["bam"])
    cfg, ocfg = pb.make_config(36, readlen=2000), ob.make_config(36, readlen=2000)
    wins = parity.load_windows(host, hb, synth_small["gaps"][:1], cfg)
    ctx = emu_gpu.init()
    ptr, nbytes = host.window_arena(wins[0][0])
    emu_gpu.host_register(ctx, ptr, nbytes)
    b, layout, res, tags, ids, rc = parity.run_gpu_batch(emu_gpu, ctx, host, wins, cfg)
    assert rc == 0
    t = b.timing()
    assert 7 <= t.launches <= 9  # one more than the copy path: the gather kernel
    (w, n, chrom, s, e), (first, _) = wins[0], layout[0]
    p = ob.port_window(host.window_descs(w), n, s, e, ocfg)
    assert not parity.compare_window(b, 0, first, n, res, tags, ids, p)
    assert t.bytes_h2d < 1.1 * nbytes  # gathered payload + descriptors, counted once
    b.end()
    emu_gpu.host_unregister(ctx, ptr)
    emu_gpu.destroy(ctx)
    host.bam_close(hb)


def test_emulated_long_reads_uncached_keys(emu_gpu, built, tmp_path):
    # reads that span more than 256 methmer sites: their keys are scored straight from the pool
    import conftest
    data = conftest.run_synth(str(tmp_path / "long"), ["-c", "32", "-s", "41", "-C", "chrL:900000:0-600000", "--readlen", "70000",
                                                        "--block", "250000", "--gap", "20000-30000"])
    host = pb.load_host()
    hb = host.bam_open(data["bam"])
    cfg, ocfg = pb.make_config(32), ob.make_config(32)
    wins = parity.load_windows(host, hb, data["gaps"][:1], cfg)
    ctx = emu_gpu.init()
    b, layout, res, tags, ids, rc = parity.run_gpu_batch(emu_gpu, ctx, host, wins, cfg)
    assert rc == 0
    (w, n, chrom, s, e), (first, _) = wins[0], layout[0]
    p = ob.port_window(host.window_descs(w), n, s, e, ocfg)
    assert max(len(b.mmrs(first + i, 0)[0]) for i in range(n)) > 256
    assert not parity.compare_window(b, 0, first, n, res, tags, ids, p, deep=False)
    b.end()
    emu_gpu.destroy(ctx)
    host.bam_close(hb)
