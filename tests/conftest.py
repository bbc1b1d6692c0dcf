import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "cuda_emu"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "emu: steps the real kernel sources on the CPU SIMT emulator (tests/cuda_emu)")


@pytest.fixture(scope="session")
def built():
    """Host-side native pieces (front end library, generator, oracle).  The CUDA library is built by
    __graft_entry__.build(); GPU tests load it and fail loudly if it is missing."""
    from pomfret_b200 import build
    build.build_host()
    build.build_oracle()
    return True


def run_synth(prefix, args):
    exe = os.path.join(ROOT, "pomfret_b200", "bin", "pomfret-synth")
    subprocess.check_call([exe, "-o", prefix] + args, stderr=subprocess.DEVNULL)
    gaps = []
    for line in open(prefix + ".truth.tsv"):
        if line.startswith("#"):
            continue
        c, s, e, t = line.split()
        gaps.append((c, int(s), int(e), t))
    return dict(prefix=prefix, bam=prefix + ".bam", vcf=prefix + ".vcf.gz", gaps=gaps)


@pytest.fixture(scope="session")
def synth30(built, tmp_path_factory):
    """~1.3 Mb of chr20-like 30x data with a few phase-block gaps (cut down)."""
    d = tmp_path_factory.mktemp("synth30")
    return run_synth(str(d / "s30"), ["-c", "30", "-s", "7", "-C", "chr20:64444167:2000000-3300000", "-F", "2",
                                      "--block", "250000", "--gap", "20000-90000"])


@pytest.fixture(scope="session")
def synth_small(built, tmp_path_factory):
    """One or two windows of short-read data: quick enough for the emulator."""
    d = tmp_path_factory.mktemp("synthsmall")
    return run_synth(str(d / "small"), ["-c", "36", "-s", "11", "-C", "chrT:400000:0-260000", "--readlen", "4000",
                                        "--block", "90000", "--gap", "9000-12000"])


@pytest.fixture(scope="session")
def synth_implicit(built, tmp_path_factory):
    """Reads whose MM lists also carry non-CpG cytosines: exercises has_implicit (blockjoin.c:856, 666-700)."""
    d = tmp_path_factory.mktemp("synthimp")
    return run_synth(str(d / "imp"), ["-c", "34", "-s", "5", "-C", "chrI:300000:0-200000", "--readlen", "3000",
                                      "--block", "70000", "--gap", "8000-10000", "--implicit", "0.01"])


@pytest.fixture(scope="session")
def synth_sparse_implicit(built, tmp_path_factory):
    """MM lists that name only 40% of the CpGs plus a few non-CpG cytosines: the implicit-canonical fill
    produces far more calls than listed bases (call-slot overflow path of the engine)."""
    d = tmp_path_factory.mktemp("synthsparse")
    return run_synth(str(d / "sp"), ["-c", "34", "-s", "6", "-C", "chrS:300000:0-200000", "--readlen", "3000",
                                     "--block", "70000", "--gap", "8000-10000", "--implicit", "0.01", "--listed", "0.4"])


@pytest.fixture(scope="session")
def synth60(built, tmp_path_factory):
    """BASELINE.json config-3 depth: 60x, 20 kb reads, five windows whose reference decisions are -1, -1, 1, 0, 1
    (`-c 60` => cov_for_selection 7, cov_for_runtime 14, 16 candidates; ~400-900 reads per window)."""
    d = tmp_path_factory.mktemp("synth60")
    return run_synth(str(d / "s60"), ["-c", "60", "-s", "68", "-C", "chr20:64444167:4000000-6600000", "-F", "3", "--block", "220000",
                                      "--gap", "40000-330000", "--frac-meth", "0.815", "--frac-unmeth", "0.15"])


@pytest.fixture(scope="session")
def synth_wgs5(built, tmp_path_factory):
    """WGS-shaped 30x sample: five contigs of the hg38 list (chr7 has no phase-block gap at all)."""
    d = tmp_path_factory.mktemp("wgs5")
    return run_synth(str(d / "wgs5"), ["-c", "30", "-s", "90", "-C", "chr1:248956422:1000000-2200000", "-C", "chr2:242193529:5000000-6000000",
                                       "-C", "chr7:159345973:3000000-3300000", "-C", "chr20:64444167:2000000-2900000", "-C",
                                       "chrX:156040895:1000000-1800000", "--block", "300000", "--gap", "20000-120000"])


@pytest.fixture(scope="session")
def synth_config1(built, tmp_path_factory):
    """BASELINE.json config 1 (the bundled quick start): the example's own variants.vcf.gz, reads simulated on
    chr6:11.01-11.21 Mb of its 195-contig header carrying that call set (the example BAM itself is not distributed)."""
    d = tmp_path_factory.mktemp("config1")
    vcf = os.path.join(ROOT, "tests", "golden", "config1_variants.vcf.gz")
    data = run_synth(str(d / "q"), ["-c", "32", "-s", "64", "--vcf-in", vcf, "-C", "chr6:0:11010000-11210000"])
    data["vcf"] = vcf
    return data


@pytest.fixture(scope="session")
def synth_long_cigar(built, tmp_path_factory):
    """Reads of ~260 kb at 30 % error: more than 65535 CIGAR operations per record (CG:B,I tag in the BAM)."""
    d = tmp_path_factory.mktemp("longcigar")
    return run_synth(str(d / "u"), ["-c", "32", "-s", "3", "-C", "chrU:900000", "--readlen", "260000", "--err", "0.3", "--de-cap", "0.05",
                                    "--block", "300000", "--gap", "20000-30000"])
