"""Decode kernel (MM/ML + CIGAR -> reference-coordinate calls) against the oracle port on hand-built
records: grammar variants, both strands, clips/indels/fatal ops, long reads, malformed tags."""
import pytest

import decode_fuzz
import pomfret_b200 as pb


@pytest.mark.emu
@pytest.mark.parametrize("seed", [1, 2])
def test_decode_fuzz_emulated(built, seed):
    import build_emu
    gpu = pb.load_gpu(build_emu.build())
    R = decode_fuzz.build_records(seed, n_random=40, max_len=6000, long_lens=(40000,))
    bad = decode_fuzz.check_against_port(gpu, R)
    assert not bad, bad[:10]


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_decode_fuzz_gpu(built, seed):
    gpu = pb.load_gpu()
    R = decode_fuzz.build_records(seed, n_random=400, max_len=30000, long_lens=(70000, 300000))
    bad = decode_fuzz.check_against_port(gpu, R)
    assert not bad, bad[:10]


@pytest.mark.emu
def test_decode_fuzz_gathered_emulated(built):
    # the same records in one registered buffer, every field at a random alignment: no host copy, gather_kernel
    import build_emu
    gpu = pb.load_gpu(build_emu.build())
    R = decode_fuzz.build_records(5, n_random=30, max_len=5000, long_lens=(20000,))
    reg = decode_fuzz.pack_into_one_buffer(R, seed=5)
    bad = decode_fuzz.check_against_port(gpu, R, register=reg)
    assert not bad, bad[:10]


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [6, 7])
def test_decode_fuzz_gathered_gpu(built, seed):
    gpu = pb.load_gpu()
    R = decode_fuzz.build_records(seed, n_random=300, max_len=30000, long_lens=(70000,))
    reg = decode_fuzz.pack_into_one_buffer(R, seed=seed)
    bad = decode_fuzz.check_against_port(gpu, R, register=reg)
    assert not bad, bad[:10]
