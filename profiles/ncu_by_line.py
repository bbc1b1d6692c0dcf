#!/usr/bin/env python
"""Join the per-instruction page of an .ncu-rep with the source lines of the cubin inside libpomfret_gpu.so
(nvdisasm -g), and print executed warp instructions and stall samples per source line.
usage: ncu_by_line.py report.ncu-rep kernel_regex [top_n]   (runs here, without a GPU)"""
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cubin_lines(kernel_rx):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "pomfret_b200", "lib", "libpomfret_gpu.so")], cwd=d,
                   stdout=subprocess.DEVNULL)
    out = []
    for cub in glob.glob(os.path.join(d, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
        cur_fn, line, inl = None, None, None
        for l in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
            if m:
                cur_fn = m.group(1)
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m and cur_fn and re.search(kernel_rx, cur_fn):
                out.append((int(m.group(1), 16), line, m.group(2)))
    return out


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[1]
    iS, iI, iA = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Address")
    inst = []
    seen = set()
    for r in rows[2:]:
        if len(r) <= iI or not r[iI].isdigit():
            continue
        if r[iA] in seen:
            break  # a second launch of the same kernel
        seen.add(r[iA])
        inst.append((int(r[iA], 16), int(r[iS] or 0), int(r[iI])))
    dis = cubin_lines(rx)
    if len(dis) != len(inst):
        print("warning: %d SASS instructions in the cubin vs %d in the report (stale build?)" % (len(dis), len(inst)))
    per = {}
    for (a, s, n), (off, line, text) in zip(inst, dis):
        e = per.setdefault(line, [0, 0])
        e[0] += n
        e[1] += s
    tot_i = sum(v[0] for v in per.values())
    tot_s = sum(v[1] for v in per.values())
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    src_cache = {}
    for line, (n, s) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if line:
            for base, _, names in os.walk(os.path.join(ROOT, "pomfret_b200", "csrc")):
                if line[0] in names:
                    p = os.path.join(base, line[0])
                    src_cache.setdefault(p, open(p).read().split("\n"))
                    text = src_cache[p][line[1] - 1].strip()[:90]
        print("%6.2f%% inst %6.2f%% stall  %s:%s  %s" % (100.0 * n / tot_i, 100.0 * s / max(tot_s, 1), line[0] if line else "?",
                                                        line[1] if line else "?", text))


if __name__ == "__main__":
    main()
