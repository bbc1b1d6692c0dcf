"""Host front end (rows a16/a17 and the loaders): interval bookkeeping against the compiled reference, and
whole-program runs of `pomfret methphase|report` whose output files must be byte-identical to the
reference binary's.  On a box without a GPU the program is pointed at the emulated build of the kernels
(tests/cuda_emu) through POMFRET_GPU_LIB; with a GPU (`-m gpu` variants in test_gpu_cli.py) it loads the
real libpomfret_gpu.so."""
import ctypes as C
import filecmp
import os
import subprocess

import numpy as np
import pytest

import conftest
import oracle_bindings as ob
import pomfret_b200 as pb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MINE = os.path.join(ROOT, "pomfret_b200", "bin", "pomfret")
needs_ref = pytest.mark.skipif(not os.path.exists(ob.REF_SO), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def two_contigs(built, tmp_path_factory):
    d = tmp_path_factory.mktemp("two")
    return conftest.run_synth(str(d / "two"), ["-c", "36", "-s", "21", "-C", "chrA:500000:0-330000", "-C", "chrB:400000:0-250000",
                                               "--readlen", "4000", "--block", "60000", "--gap", "9000-30000"])


def _host():
    host = pb.load_host()
    lib = host.lib
    vp = C.c_void_p
    lib.pomfret_host_intervals_load.restype = vp
    lib.pomfret_host_intervals_load.argtypes = [C.c_char_p, C.c_int]
    for f in ("nref", "finish", "free"):
        getattr(lib, "pomfret_host_intervals_" + f).argtypes = [vp]
    lib.pomfret_host_intervals_refname.restype = C.c_char_p
    lib.pomfret_host_intervals_refname.argtypes = [vp, C.c_int]
    lib.pomfret_host_intervals_n.argtypes = [vp, C.c_int]
    lib.pomfret_host_intervals_get.argtypes = [vp, C.c_int, vp, vp, vp]
    lib.pomfret_host_intervals_decide.argtypes = [vp, C.c_int, vp]
    lib.pomfret_host_intervals_nblocks.argtypes = [vp, C.c_int]
    lib.pomfret_host_intervals_blocks.argtypes = [vp, C.c_int, vp, vp]
    lib.pomfret_host_load_variants.argtypes = [C.c_char_p, C.c_char_p, vp, C.c_int, vp, C.c_int, C.POINTER(C.c_int)]
    return lib


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_interval_bookkeeping_matches_reference(two_contigs, seed):
    mine = _host()
    ref = ob.ref_lib()
    fn = two_contigs["vcf"].encode()
    with ob.quiet_reference():
        hr = ref.refh_load_intervals(fn, 1)
    hm = mine.pomfret_host_intervals_load(fn, 1)
    assert hm
    nref = ref.refh_intervals_nref(hr)
    assert nref == mine.pomfret_host_intervals_nref(hm) == 2
    rng = np.random.default_rng(seed)
    for i in range(nref):
        assert ref.refh_intervals_refname(hr, i) == mine.pomfret_host_intervals_refname(hm, i)
        n = ref.refh_intervals_n(hr, i)
        assert n == mine.pomfret_host_intervals_n(hm, i) and n > 0
        a = [np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(2, np.uint32)]
        b = [np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(2, np.uint32)]
        ref.refh_intervals_get(hr, i, *(x.ctypes.data for x in a))
        mine.pomfret_host_intervals_get(hm, i, *(x.ctypes.data for x in b))
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        dec = rng.integers(-1, 2, size=n).astype(np.int32)
        ref.refh_intervals_decide(hr, i, dec.ctypes.data)
        mine.pomfret_host_intervals_decide(hm, i, dec.ctypes.data)
    with ob.quiet_reference():
        ref.refh_intervals_finish(hr)
    mine.pomfret_host_intervals_finish(hm)
    for i in range(nref):
        nb = ref.refh_intervals_nblocks(hr, i)
        assert nb == mine.pomfret_host_intervals_nblocks(hm, i)
        a = [np.zeros(nb, np.uint32), np.zeros(nb, np.uint32)]
        b = [np.zeros(nb, np.uint32), np.zeros(nb, np.uint32)]
        ref.refh_intervals_blocks(hr, i, a[0].ctypes.data, a[1].ctypes.data)
        mine.pomfret_host_intervals_blocks(hm, i, b[0].ctypes.data, b[1].ctypes.data)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    ref.refh_intervals_free(hr)
    mine.pomfret_host_intervals_free(hm)


@needs_ref
def test_known_variants_match_reference(two_contigs):
    mine = _host()
    ref = ob.ref_lib()
    cap = 1 << 16
    for chrom in (b"chrA", b"chrB"):
        pos, ln, bo = (np.zeros(cap, np.uint32) for _ in range(3))
        op, hp = (np.zeros(cap, np.uint8) for _ in range(2))
        bases = np.zeros(cap * 4, np.uint8)
        with ob.quiet_reference():
            n = ref.refh_load_variants(two_contigs["vcf"].encode(), chrom, pos.ctypes.data, ln.ctypes.data, op.ctypes.data,
                                       hp.ctypes.data, bases.ctypes.data, bo.ctypes.data, cap, cap * 4)
        vars_ = (pb.Variant * cap)()
        mb = np.zeros(cap * 4, np.uint8)
        nb = C.c_int()
        m = mine.pomfret_host_load_variants(two_contigs["vcf"].encode(), chrom, vars_, cap, mb.ctypes.data, cap * 4, C.byref(nb))
        assert n == m and n > 50
        for i in range(n):
            v = vars_[i]
            assert (v.pos, v.len, v.op, v.haptag) == (pos[i], ln[i], op[i], hp[i])
            assert np.array_equal(mb[v.bases_off:v.bases_off + v.len], bases[bo[i]:bo[i] + ln[i]])


def run_both(tmp, data, args, gpu_lib, outputs, sub="methphase"):
    """Run the reference binary and this repo's front end with the same arguments; compare the files."""
    env = dict(os.environ)
    if gpu_lib:
        env["POMFRET_GPU_LIB"] = gpu_lib
        # (the emulator steps the inflate kernel far too slowly for whole files: the emulated runs use the host loader;
        #  test_golden.py runs one small case through the compressed ingest, the GPU tests all of them)
        env.setdefault("POMFRET_HOST_INFLATE", "1")
    res = {}
    for who, exe in (("ref", ob.REF_BIN), ("mine", MINE)):
        prefix = os.path.join(tmp, who)
        a = list(args)
        if who == "ref":  # options that only this implementation has (they never change results)
            for opt in ("--windows-per-batch", "--gpus"):
                while opt in a:
                    i = a.index(opt)
                    del a[i:i + 2]
        cmd = [exe, sub] + a + ["-o", prefix, "--vcf", data["vcf"], data["bam"]]
        p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert p.returncode == 0, (who, p.stderr[-2000:])
        res[who] = (prefix, p.stdout)
    for suffix in outputs:
        a, b = res["ref"][0] + suffix, res["mine"][0] + suffix
        assert os.path.exists(a) and os.path.exists(b), suffix
        assert filecmp.cmp(a, b, shallow=False), "%s differs" % suffix
    return res


@pytest.fixture(scope="module")
def emu_lib(built):
    import build_emu
    return build_emu.build()


@needs_ref
@pytest.mark.emu
def test_methphase_files_identical_two_contigs(two_contigs, emu_lib, tmp_path):
    # two contigs (abs_start quirk on the second), short blocks => merged gaps and dropped-interval rescue
    run_both(str(tmp_path), two_contigs, ["-t", "2", "-c", "36", "-L", "2000", "--write-bam", "--output-tsv"], emu_lib,
             [".mp.gtf", ".mp.tsv", ".mp.vcf", ".mp.bam", ".mp.bam.bai"])


@needs_ref
@pytest.mark.emu
def test_methphase_untagged_files_identical(built, emu_lib, tmp_path):
    data = conftest.run_synth(str(tmp_path / "unt"), ["-c", "36", "-s", "31", "-C", "chrU:300000:0-200000", "--readlen", "4000",
                                                      "--block", "70000", "--gap", "9000-12000", "--untagged"])
    run_both(str(tmp_path), data, ["-u", "-c", "36", "-L", "2000", "--write-bam"], emu_lib,
             [".mp.gtf", ".mp.vcf", ".mp.bam", ".mp.bam.bai"])


@needs_ref
@pytest.mark.emu
def test_report_identical(synth_small, emu_lib, tmp_path):
    res = run_both(str(tmp_path), synth_small, ["-c", "36", "-L", "2000", "--chunk-size", "9000", "--chunk-stride", "30000"],
                   emu_lib, [".report.tsv"], sub="report")
    assert res["ref"][1] == res["mine"][1]  # stdout: the accuracy lines
    assert sum(1 for _ in open(res["mine"][0] + ".report.tsv")) >= 2


def test_front_end_fails_loudly_without_engine(synth_small, tmp_path):
    env = dict(os.environ)
    env["POMFRET_GPU_LIB"] = "/nonexistent/libpomfret_gpu.so"
    env["LD_LIBRARY_PATH"] = ""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the default library would work")
    p = subprocess.run([MINE, "methphase", "-c", "36", "-L", "2000", "-o", str(tmp_path / "x"), "--vcf", synth_small["vcf"],
                        synth_small["bam"]], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode != 0
    assert "no CPU" in p.stderr or "no CUDA device" in p.stderr
    assert not os.path.exists(str(tmp_path / "x.mp.gtf"))
