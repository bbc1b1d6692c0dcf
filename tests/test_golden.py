"""Committed golden vectors (tests/golden/, produced from the unmodified reference by make_golden.py):
the oracle port, the front end on the emulated kernels (CPU) and the front end on the CUDA library (GPU)
must reproduce them.  Nothing here reads /root/reference or needs oracle/_ref at run time."""
import filecmp
import json
import os
import subprocess

import pytest

import conftest
import oracle_bindings as ob
import parity
import pomfret_b200 as pb

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))
MINE = os.path.join(os.path.dirname(HERE), "pomfret_b200", "bin", "pomfret")


def _inputs(tmp_path, name):
    case = MANIFEST[name]
    data = conftest.run_synth(str(tmp_path / "in"), [a.replace("{GOLD}", GOLD) for a in case["synth"]])
    if "vcf" in case:  # a committed input call set (config 1: the example's variants.vcf.gz)
        data["vcf"] = os.path.join(GOLD, case["vcf"])
    return data


def _run_front_end(tmp_path, name, gpu_lib, host_inflate=False):
    case = MANIFEST[name]
    data = _inputs(tmp_path, name)
    env = dict(os.environ)
    if gpu_lib:
        env["POMFRET_GPU_LIB"] = gpu_lib
    if host_inflate:
        env["POMFRET_HOST_INFLATE"] = "1"  # the host loader instead of the compressed ingest
    prefix = str(tmp_path / "mine")
    p = subprocess.run([MINE, case["sub"]] + case["args"] + ["-o", prefix, "--vcf", data["vcf"], data["bam"]], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    for suf in case["outputs"]:
        assert filecmp.cmp(prefix + suf, os.path.join(GOLD, name + suf), shallow=False), "%s%s differs from the golden" % (name, suf)


@pytest.mark.parametrize("name", [k for k, v in MANIFEST.items() if "windows" in v])
def test_oracle_port_reproduces_golden_windows(built, tmp_path, name):
    """per-window results of the reference's haplotag_region_given_bam, recorded in the manifest"""
    case = MANIFEST[name]
    data = _inputs(tmp_path, name)
    cov = int(case["args"][case["args"].index("-c") + 1])
    readlen = int(case["args"][case["args"].index("-L") + 1]) if "-L" in case["args"] else 15000
    host = pb.load_host()
    cfg, ocfg = pb.make_config(cov, readlen=readlen), ob.make_config(cov, readlen=readlen)
    hb = host.bam_open(data["bam"])
    assert [(c, s, e) for c, s, e, _ in data["gaps"]] == [(w["chrom"], w["start"], w["end"]) for w in case["windows"]]
    for (w, n, chrom, s, e), gold in zip(parity.load_windows(host, hb, data["gaps"], cfg), case["windows"]):
        p = ob.port_window(host.window_descs(w), n, s, e, ocfg)
        got = {"decision": int(p["decision"]), "join_fwd": int(p["join_fwd"]), "join_bwd": int(p["join_bwd"]),
               "n_reads": int(p["n_reads"]), "n_sites": int(len(p["sites_fwd"])), "n_calls": int(len(p["calls_pos"])),
               "calls_checksum": int(sum(int(x) for x in p["calls_pos"]) % (1 << 61)),
               "tags_final": "".join(str(int(t)) for t in p["tags_final"])}
        assert got == {k: gold[k] for k in got}, (chrom, s, e)
        host.window_free(w)
    host.bam_close(hb)


def test_config1_replica_reproduces_the_bundled_example_output():
    """The reference's committed quick-start output (example/output.mp.vcf) against what the compiled reference
    writes for the replica of its input: identical up to the one record that is stale against the reference's own
    current code; recorded by make_golden.py where /root/reference exists."""
    case = MANIFEST["config1_quickstart"]
    assert case["example_output_vcf_lines"] == 347 and case["example_output_vcf_diff_lines"] == [344]
    assert [w["decision"] for w in case["windows"]] == [1]


@pytest.mark.emu
@pytest.mark.parametrize("name", ["small_methphase", "untagged_methphase", "config1_quickstart"])  # the rest runs on the GPU (and in test_host_frontend.py)
def test_front_end_reproduces_golden_files_emulated(built, tmp_path, name):
    import build_emu
    # (host loader: the emulator steps the warp-cooperative inflate kernel far too slowly for whole files; the compressed
    #  ingest is covered on the emulator by tests/test_ingest.py on a tiny file, and by every GPU run of the front end)
    _run_front_end(tmp_path, name, build_emu.build(), host_inflate=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MANIFEST))
def test_front_end_reproduces_golden_files_gpu(built, tmp_path, name):
    _run_front_end(tmp_path, name, None)
