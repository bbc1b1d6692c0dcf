#!/usr/bin/env python
"""Generate the golden vectors of tests/golden/ from the UNMODIFIED reference (oracle/_ref, compiled from
/root/reference by oracle/Makefile).  Run in the build container, where the reference exists:

    python tests/golden/make_golden.py

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
    for name, (synth, sub, args, outs, win) in CASES.items():
        tmp = tempfile.mkdtemp(prefix="golden_")
        data = conftest.run_synth(os.path.join(tmp, "in"), synth)
        prefix = os.path.join(tmp, "ref")
        subprocess.run([ob.REF_BIN, sub] + args + ["-o", prefix, "--vcf", data["vcf"], data["bam"]], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        entry = {"synth": synth, "sub": sub, "args": args, "outputs": outs}
        for suf in outs:
            shutil.copy(prefix + suf, os.path.join(HERE, name + suf))
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
