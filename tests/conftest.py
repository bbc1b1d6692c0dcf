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
