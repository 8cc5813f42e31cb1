#!/usr/bin/env python
"""Per-launch summary of an ncu --set full report: the metrics the roofline discussion uses.  usage: python profiles/ncu_summary.py report.ncu-rep"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__lsuin_requests.avg.pct_of_peak_sustained_elapsed"]
ix = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
if not stall:
    stall = [h for h in hdr if "warp_issue_stalled" in h and h.endswith(".pct")]
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    print("==", r[ix["Kernel Name"]][:110], "  [id %s]" % r[ix["ID"]])
    for w in want:
        if w in ix:
            print("  %-78s %s" % (w, r[ix[w]]))
    st = []
    for s in stall:
        try:
            st.append((float(r[ix[s]].replace(",", "")), s.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        except ValueError:
            pass
    st.sort(reverse=True)
    print("  stalls (warps per issue):", ", ".join("%s %.2f" % (n, v) for v, n in st[:8]))
