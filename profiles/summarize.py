#!/usr/bin/env python
"""Turns an ncu report (gpurun_out/*.ncu-rep) into the compact per-kernel text summary committed under profiles/.
usage: python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_<what>.txt"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
    print("# source: %s (ncu --set full --clock-control none; per-launch values, cold cache, serialised)" % rep)
    for r in rows[2:]:
        print("\n== %s   [launch id %s]" % (r[idx["Kernel Name"]][:110], r[idx["ID"]]))
        for w in WANT:
            if w in idx:
                print("  %-72s %18s %s" % (w, r[idx[w]], units[idx[w]]))
        vals = [(float(r[idx[n]].replace(",", "") or 0), n) for n in stall]
        tot = sum(v for v, _ in vals) or 1.0
        print("  stall samples: " + ", ".join("%s %.0f%%" % (n.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot)
                                              for v, n in sorted(vals, reverse=True)[:7]))


if __name__ == "__main__":
    main()
