"""`pomfret methphase|report` on the real CUDA library: output files byte-identical to the reference
binary (oracle/_ref/pomfret, prebuilt and shipped to the GPU box)."""
import os

import pytest

import conftest
import oracle_bindings as ob
from test_host_frontend import run_both

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(ob.REF_BIN), reason="oracle/_ref/pomfret not built")]


@pytest.fixture(scope="module")
def two_contigs(built, tmp_path_factory):
    d = tmp_path_factory.mktemp("two")
    return conftest.run_synth(str(d / "two"), ["-c", "36", "-s", "21", "-C", "chrA:500000:0-330000", "-C", "chrB:400000:0-250000",
                                               "--readlen", "4000", "--block", "60000", "--gap", "9000-30000"])


def test_methphase_30x_files_identical(synth30, tmp_path):
    run_both(str(tmp_path), synth30, ["-t", "4", "-c", "30", "--write-bam", "--output-tsv"], None,
             [".mp.gtf", ".mp.tsv", ".mp.vcf", ".mp.bam", ".mp.bam.bai"])


def test_methphase_two_contigs_dropped_intervals(two_contigs, tmp_path):
    run_both(str(tmp_path), two_contigs, ["-t", "3", "-c", "36", "-L", "2000", "--write-bam"], None,
             [".mp.gtf", ".mp.vcf", ".mp.bam", ".mp.bam.bai"])


def test_methphase_untagged(built, tmp_path):
    data = conftest.run_synth(str(tmp_path / "unt"), ["-c", "30", "-s", "33", "-C", "chr20:64444167:5000000-6200000", "-F", "4",
                                                      "--untagged"])
    run_both(str(tmp_path), data, ["-u", "-t", "2", "-c", "30", "--write-bam"], None, [".mp.gtf", ".mp.vcf", ".mp.bam", ".mp.bam.bai"])


def test_report(synth30, tmp_path):
    res = run_both(str(tmp_path), synth30, ["-c", "30", "--chunk-size", "50000", "--chunk-stride", "100000"], None,
                   [".report.tsv"], sub="report")
    assert res["ref"][1] == res["mine"][1]


def test_methphase_small_batches_and_many_workers(synth30, tmp_path):
    # chunking of windows into batches / workers must not change any output
    run_both(str(tmp_path), synth30, ["-t", "8", "-c", "30", "--windows-per-batch", "1", "--write-bam"], None,
             [".mp.gtf", ".mp.vcf", ".mp.bam"])
