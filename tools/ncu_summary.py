#!/usr/bin/env python
"""Compact summary of an ncu report: per kernel the headline metrics, and the SASS hot regions
(consecutive instructions grouped, with executed warp-instructions and average active lanes).
usage: tools/ncu_summary.py report.ncu-rep [kernel-substring] [group]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""; grp = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if pat and pat not in d["Kernel Name"]: continue
    print("=====", d["Kernel Name"][:70])
    for w in want:
        if w in d: print("   %-70s %s" % (w, d[w]))
    st = []
    for h in hdr:
        if "average_warps_issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h and d[h]:
            st.append((float(d[h]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    print("   stalls/issue:", ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
cur = None; kerns = []
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; kerns.append(cur); continue
    if r[0] == "Address": cur["hdr"] = r; continue
    if cur is not None: cur["rows"].append(r)
seen = set()
for k in kerns:
    if pat and pat not in k["name"]: continue
    if k["name"] in seen or "hdr" not in k: continue
    seen.add(k["name"])
    h = k["hdr"]; iS = h.index("Source"); iI = h.index("Instructions Executed"); iT = h.index("Thread Instructions Executed")
    iW = h.index("Warp Stall Sampling (All Samples)")
    tot = sum(int(r[iI]) for r in k["rows"]); smp = sum(int(r[iW]) for r in k["rows"])
    print("----- SASS regions of", k["name"][:60], "total warp-inst %.1fM, samples %d" % (tot / 1e6, smp))
    R = k["rows"]
    for a in range(0, len(R), grp):
        blk = R[a:a + grp]
        c = sum(int(r[iI]) for r in blk); th = sum(int(r[iT]) for r in blk); s = sum(int(r[iW]) for r in blk)
        if c < tot * 0.004 and s < smp * 0.004: continue
        ops = {}
        for r in blk:
            op = r[iS].split()[0] if not r[iS].strip().startswith("@") else r[iS].split()[1]
            op = op.split(".")[0]
            if int(r[iI]): ops[op] = ops.get(op, 0) + 1
        top = " ".join("%s%d" % (o, n) for o, n in sorted(ops.items(), key=lambda x: -x[1])[:7])
        print("  [%4d-%4d] inst %6.2fM (%4.1f%%) lanes %4.1f  stall-samples %4.1f%%  %s" % (a, a + len(blk) - 1, c / 1e6, 100.0 * c / tot, th / max(1, c), 100.0 * s / max(1, smp), top))
