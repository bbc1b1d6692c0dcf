#!/usr/bin/env python
"""Generate the golden vectors of tests/golden/ from the UNMODIFIED reference (oracle/_ref, compiled from
/root/reference by oracle/Makefile).  Run in the build container, where the reference exists:

    python tests/golden/make_golden.py [case ...]

For each case: the synthetic BAM/VCF is produced by pomfret-synth from a fixed seed (the generator is
deterministic, so the GPU box regenerates identical inputs), the reference binary is run on it, and its
output text files are committed together with the per-window decisions of haplotag_region_given_bam
(reference blockjoin.c:4217-4335) obtained through oracle/ref_harness.c.  The reference ships no tests and
its example/ outputs are stale against its own code, so these are the pinned goldens."""
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    # name: (synth args, sub-command, program args, output suffixes, window cov/readlen for the decision dump)
    "small_methphase": (["-c", "36", "-s", "11", "-C", "chrT:400000:0-260000", "--readlen", "4000", "--block", "90000",
                         "--gap", "9000-12000"], "methphase", ["-c", "36", "-L", "2000", "--output-tsv"],
                        [".mp.gtf", ".mp.vcf", ".mp.tsv"], (36, 2000)),
    "two_contigs_methphase": (["-c", "36", "-s", "21", "-C", "chrA:500000:0-330000", "-C", "chrB:400000:0-250000", "--readlen",
                               "4000", "--block", "60000", "--gap", "9000-30000"], "methphase",
                              ["-t", "2", "-c", "36", "-L", "2000"], [".mp.gtf", ".mp.vcf"], (36, 2000)),
    "chr20_30x_methphase": (["-c", "30", "-s", "7", "-C", "chr20:64444167:2000000-3300000", "-F", "2", "--block", "250000",
                             "--gap", "20000-90000"], "methphase", ["-t", "4", "-c", "30", "--output-tsv"],
                            [".mp.gtf", ".mp.vcf", ".mp.tsv"], (30, 15000)),
    "untagged_methphase": (["-c", "36", "-s", "31", "-C", "chrU:300000:0-200000", "--readlen", "4000", "--block", "70000",
                            "--gap", "9000-12000", "--untagged"], "methphase", ["-u", "-c", "36", "-L", "2000"],
                           [".mp.gtf", ".mp.vcf"], None),
    # BASELINE.json config 3 depth: -c 60 => cov_for_selection 7, runtime 14, 16 candidates; decisions -1,-1,1,0,1
    "x60_methphase": (["-c", "60", "-s", "68", "-C", "chr20:64444167:4000000-6600000", "-F", "3", "--block", "220000", "--gap",
                       "40000-330000", "--frac-meth", "0.815", "--frac-unmeth", "0.15"], "methphase",
                      ["-t", "2", "-c", "60", "--output-tsv"], [".mp.gtf", ".mp.vcf", ".mp.tsv"], (60, 15000)),
    # BASELINE.json config 1 (bundled quick start): the example's own variants.vcf.gz (195-contig header, chr6 = tid 5,
    # one gap 11092382-11147866), reads simulated on chr6:11.01-11.21 Mb carrying that call set, `-c 60 --write-bam`
    "config1_quickstart": (["-c", "32", "-s", "64", "--vcf-in", "{GOLD}/config1_variants.vcf.gz", "-C", "chr6:0:11010000-11210000"],
                           "methphase", ["-c", "60"], [".mp.gtf", ".mp.vcf"], (60, 15000), "config1_variants.vcf.gz"),
    # WGS-shaped: five contigs of the hg38 list, chr7 without any gap, one worker thread per contig
    "wgs5_methphase": (["-c", "30", "-s", "90", "-C", "chr1:248956422:1000000-2200000", "-C", "chr2:242193529:5000000-6000000",
                        "-C", "chr7:159345973:3000000-3300000", "-C", "chr20:64444167:2000000-2900000", "-C",
                        "chrX:156040895:1000000-1800000", "--block", "300000", "--gap", "20000-120000"], "methphase",
                       ["-t", "4", "-c", "30"], [".mp.gtf", ".mp.vcf"], (30, 15000)),
    # alignments with more than 65535 CIGAR operations (CG:B,I tag, SAM spec 4.2.2)
    "long_cigar_methphase": (["-c", "32", "-s", "3", "-C", "chrU:900000", "--readlen", "260000", "--err", "0.3", "--de-cap", "0.05",
                              "--block", "300000", "--gap", "20000-30000"], "methphase", ["-c", "30"], [".mp.gtf", ".mp.vcf"],
                             (30, 15000)),
    "small_report": (["-c", "36", "-s", "11", "-C", "chrT:400000:0-260000", "--readlen", "4000", "--block", "90000",
                      "--gap", "9000-12000"], "report", ["-c", "36", "-L", "2000", "--chunk-size", "8000", "--chunk-stride", "20000"],
                     [".report.tsv"], None),
}


def main():
    import conftest
    import oracle_bindings as ob
    from pomfret_b200 import build
    build.build_host()
    build.build_oracle()
    assert os.path.exists(ob.REF_BIN), "oracle/_ref/pomfret missing: this script needs /root/reference"
    manifest = {}
    only = set(sys.argv[1:])
    if only:
        manifest = json.load(open(os.path.join(HERE, "manifest.json")))
    for name, case in CASES.items():
        if only and name not in only:
            continue
        synth, sub, args, outs, win = case[:5]
        vcf_fixture = case[5] if len(case) > 5 else None  # a committed input VCF instead of the simulated one
        tmp = tempfile.mkdtemp(prefix="golden_")
        data = conftest.run_synth(os.path.join(tmp, "in"), [a.replace("{GOLD}", HERE) for a in synth])
        if vcf_fixture:
            data["vcf"] = os.path.join(HERE, vcf_fixture)
        prefix = os.path.join(tmp, "ref")
        subprocess.run([ob.REF_BIN, sub] + args + ["-o", prefix, "--vcf", data["vcf"], data["bam"]], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        entry = {"synth": synth, "sub": sub, "args": args, "outputs": outs}
        if vcf_fixture:
            entry["vcf"] = vcf_fixture
        for suf in outs:
            shutil.copy(prefix + suf, os.path.join(HERE, name + suf))
        if name == "config1_quickstart":
            # the reference's own committed quick-start output: the replica must reproduce it up to the lines that
            # are stale against the reference's current code (last phased record, GTF spacing)
            ex = "/root/reference/example/output.mp.vcf"
            a, b = open(prefix + ".mp.vcf").read().split("\n"), open(ex).read().split("\n")
            assert len(a) == len(b)
            entry["example_output_vcf_lines"] = len(a)
            entry["example_output_vcf_diff_lines"] = [i + 1 for i, (x, y) in enumerate(zip(a, b)) if x != y]
        if win:
            cfg = ob.make_config(win[0], readlen=win[1])
            dec = []
            for chrom, s, e, _ in data["gaps"]:
                r = ob.ref_window(data["bam"], chrom, s, e, cfg)
                dec.append({"chrom": chrom, "start": s, "end": e, "decision": int(r["decision"]), "join_fwd": int(r["join_fwd"]),
                            "join_bwd": int(r["join_bwd"]), "n_reads": int(r["n_reads"]),
                            "n_sites": int(len(r["sites_fwd"])), "n_calls": int(len(r["calls_pos"])),
                            "calls_checksum": int(sum(int(x) for x in r["calls_pos"]) % (1 << 61)),
                            "tags_final": "".join(str(int(t)) for t in r["tags_final"])})
            entry["windows"] = dec
        manifest[name] = entry
        shutil.rmtree(tmp)
    json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
