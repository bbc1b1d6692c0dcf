#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU) into the few counters the roofline discussion uses.
usage: summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_summary.txt"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__occupancy_limit_shared_mem", "occ_lim_smem"), ("launch__occupancy_limit_registers", "occ_lim_regs"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
        ("smsp__inst_executed.sum", "warp_insts"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "smem_atom_wavefronts"),
        ("smsp__inst_executed_op_shared_atom.sum", "smem_atom_insts")]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    rows = [r for r in rows if len(r) > 10]
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(name)
        for key, short in WANT:
            if key in hdr:
                i = hdr.index(key)
                print("    %-22s %s %s" % (short, r[i], units[i]))


if __name__ == "__main__":
    main(sys.argv[1])
