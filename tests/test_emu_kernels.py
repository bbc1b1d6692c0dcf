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


def test_emulated_kernels_implicit_mode(emu_gpu, synth_implicit):
    _run(emu_gpu, synth_implicit, 34, 1500)


def test_emulated_kernels_k2(emu_gpu, synth_small):
    _run(emu_gpu, synth_small, 24, 2000, k=2, k_span=900)


def test_emulated_kernels_call_slot_overflow(emu_gpu, synth_sparse_implicit):
    _run(emu_gpu, synth_sparse_implicit, 34, 1500)
