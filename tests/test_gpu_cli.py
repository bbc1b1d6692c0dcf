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


def test_methphase_60x_files_identical(synth60, tmp_path):
    # BASELINE.json config-3 depth
    run_both(str(tmp_path), synth60, ["-t", "3", "-c", "60", "--write-bam", "--output-tsv"], None,
             [".mp.gtf", ".mp.tsv", ".mp.vcf", ".mp.bam", ".mp.bam.bai"])


def test_methphase_config1_quickstart(synth_config1, tmp_path):
    # BASELINE.json config 1: `methphase -c 60 --write-bam` on the bundled call set
    run_both(str(tmp_path), synth_config1, ["-c", "60", "--write-bam"], None, [".mp.gtf", ".mp.vcf", ".mp.bam", ".mp.bam.bai"])


def test_methphase_wgs5_one_thread_per_contig(synth_wgs5, tmp_path):
    run_both(str(tmp_path), synth_wgs5, ["-t", "5", "-c", "30", "--write-bam"], None, [".mp.gtf", ".mp.vcf", ".mp.bam", ".mp.bam.bai"])


def test_methphase_long_cigar(synth_long_cigar, tmp_path):
    run_both(str(tmp_path), synth_long_cigar, ["-c", "30", "--write-bam"], None, [".mp.gtf", ".mp.vcf", ".mp.bam", ".mp.bam.bai"])


def test_report_60x(synth60, tmp_path):
    # BASELINE.json config 5 shape: report --chunk-size 50000 --chunk-stride 100000 at -c 60
    res = run_both(str(tmp_path), synth60, ["-c", "60", "--chunk-size", "50000", "--chunk-stride", "100000"], None,
                   [".report.tsv"], sub="report")
    assert res["ref"][1] == res["mine"][1]


def _n_devices():
    import pomfret_b200 as pb
    return pb.load_gpu().device_count()


@pytest.mark.parametrize("gpus", ["2", "all"])
def test_methphase_multi_gpu_files_identical(synth_wgs5, tmp_path, gpus):
    """region sets sharded over several devices of one box: same files as the reference binary"""
    if _n_devices() < 2:
        pytest.skip("needs at least two CUDA devices")
    n = _n_devices() if gpus == "all" else int(gpus)
    run_both(str(tmp_path), synth_wgs5, ["-t", str(max(4, n)), "-c", "30", "--gpus", str(n), "--write-bam"], None,
             [".mp.gtf", ".mp.vcf", ".mp.bam", ".mp.bam.bai"])


def test_methphase_untagged_multi_gpu(built, tmp_path):
    if _n_devices() < 2:
        pytest.skip("needs at least two CUDA devices")
    data = conftest.run_synth(str(tmp_path / "unt"), ["-c", "30", "-s", "35", "-C", "chr1:248956422:1000000-1700000", "-C",
                                                      "chr2:242193529:5000000-5600000", "--untagged"])
    run_both(str(tmp_path), data, ["-u", "-t", "4", "-c", "30", "--gpus", "2"], None, [".mp.gtf", ".mp.vcf"])


def test_methphase_host_inflate_loader(synth30, tmp_path, monkeypatch):
    # the host loader (shim inflate into the record arena) stays available: same files
    monkeypatch.setenv("POMFRET_HOST_INFLATE", "1")
    run_both(str(tmp_path), synth30, ["-t", "4", "-c", "30", "--write-bam"], None, [".mp.gtf", ".mp.vcf", ".mp.bam", ".mp.bam.bai"])


def test_methphase_pinned_record_slabs(synth30, tmp_path, monkeypatch):
    # host loader with registered (pinned, mapped) record slabs: payloads gathered by the device over PCIe
    monkeypatch.setenv("POMFRET_HOST_INFLATE", "1")
    monkeypatch.setenv("POMFRET_PIN_RECORDS", "1")
    run_both(str(tmp_path), synth30, ["-t", "3", "-c", "30"], None, [".mp.gtf", ".mp.vcf"])


def test_methphase_hidden_options_k_and_n(synth30, tmp_path):
    # the reference's hidden -k / -n options (cli.c:267-283) beyond what round 1 compiled in: k = 5, 200 candidates
    run_both(str(tmp_path), synth30, ["-t", "3", "-c", "30", "-k", "5", "-n", "200"], None, [".mp.gtf", ".mp.vcf"])


def test_methphase_without_cov_uses_the_coverage_estimator(built, tmp_path):
    # no -c: estimate_read_coverage_dirtyfast (blockjoin.c:951-1040) sets cov_for_selection / n_candidates per contig
    # (:4381-4390); here the estimate comes from the compressed ingest + coverage_kernel
    import re
    import subprocess
    from test_host_frontend import MINE
    # two whole contigs at different depths (the second one behind two empty header targets), 20 kb reads: the estimator only
    # counts reads of 15 kb and more
    synth30 = conftest.run_synth(str(tmp_path / "cov"), ["-c", "34", "-s", "77", "-F", "2", "-C", "chrK:1400000:0-1400000",
                                                         "-C", "chrM:900000:0-900000", "--block", "200000"])
    est = {}
    for who, exe in (("ref", ob.REF_BIN), ("mine", MINE)):
        prefix = str(tmp_path / who)
        p = subprocess.run([exe, "methphase", "-t", "2", "-o", prefix, "--vcf", synth30["vcf"], synth30["bam"]],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert p.returncode == 0, (who, p.stderr[-2000:])
        est[who] = re.findall(r"\] (\S+) est\. coverage is (\d+)", p.stderr)
    assert est["ref"] == est["mine"] and any(int(c) > 0 for _, c in est["mine"]), est
    import filecmp
    for suffix in (".mp.gtf", ".mp.vcf"):
        assert filecmp.cmp(str(tmp_path / "ref") + suffix, str(tmp_path / "mine") + suffix, shallow=False), suffix
