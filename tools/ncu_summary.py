#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page + source page) into a short text report: key counters, stall
breakdown and per-opcode instruction counts.  Usage: ncu_summary.py file.ncu-rep [out.txt]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
kv = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__average_warp_latency_per_inst_issued.ratio"]
for k in keys:
    if k in kv:
        print("%-72s %20s %s" % (k, kv[k][0], kv[k][1]), file=out)
print("-- stall reasons (warps stalled per issue-active cycle)", file=out)
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        v = float(kv[h][0])
        if v > 0.005:
            print("   %-40s %8.3f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v), file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; ix = {x: i for i, x in enumerate(h)}
cnt = collections.Counter(); smp = collections.Counter(); tot = 0; tsm = 0
for r in rows[2:]:
    if len(r) < len(h):
        continue
    s = r[ix["Source"]].strip()
    op = (s.split()[1] if s.startswith("@") else s.split()[0]).rstrip(";")
    ie = int(r[ix["Instructions Executed"]]); ns = int(r[ix["# Samples"]])
    cnt[op] += ie; smp[op] += ns; tot += ie; tsm += ns
print("-- warp instructions executed by opcode (total %d, samples %d)" % (tot, tsm), file=out)
for op, n in cnt.most_common(24):
    print("   %-22s %14d %5.1f%%   samples %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * smp[op] / max(1, tsm)), file=out)
