#!/usr/bin/env python
"""Executed warp instructions and stall samples per CUDA source line of one kernel of an .ncu-rep
(captured with --import-source on; read here, without a GPU, through ncu's own source page).
usage: ncu_by_line.py report.ncu-rep [top_n] [kernel name regex]"""
import csv
import os
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
    if len(sys.argv) > 3:
        cmd += ["-k", "regex:" + sys.argv[3], "-c", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr, lines = None, None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = os.path.basename(r[1])
        elif len(r) > 8 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].isdigit():
            d = dict(zip(hdr, r))
            try:
                inst = int(d["Instructions Executed"])
                smp = int(d["# Samples"])
            except ValueError:
                continue
            lines.append((inst, smp, cur_file, int(r[0]), r[1].strip()))
    tot_i = sum(x[0] for x in lines) or 1
    tot_s = sum(x[1] for x in lines) or 1
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    for inst, smp, f, ln, src in sorted(lines, reverse=True)[:top]:
        print("%6.2f%% inst %6.2f%% stall  %s:%d  %s" % (100.0 * inst / tot_i, 100.0 * smp / tot_s, f, ln, src[:100]))


if __name__ == "__main__":
    main()
