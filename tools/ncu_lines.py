#!/usr/bin/env python
"""Per CUDA source line: executed warp-instructions, lanes, stall samples (from `ncu --page source --print-source cuda,sass --csv`).
usage: tools/ncu_lines.py report.ncu-rep kernel-regex [min-percent]"""
import csv, subprocess, sys, io
rep, pat = sys.argv[1], sys.argv[2]; thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fn = None; h = None; data = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1]; continue
    if r[0] == "Function Name":
        if fn is None: fn = r[1]
        elif r[1] != fn: cur_other = True
        curfn = r[1]; continue
    if r[0] == "Line No": h = r; continue
    if h is None or curfn != fn or not r[0].isdigit(): continue
    iI = h.index("Instructions Executed"); iT = h.index("Thread Instructions Executed"); iW = h.index("Warp Stall Sampling (All Samples)")
    key = (fpath.split("/")[-1], int(r[0]))
    num = lambda x: int(x) if x.isdigit() else 0
    c, t, w = num(r[iI]), num(r[iT]), num(r[iW])
    o = data.get(key, [r[1], 0, 0, 0]); o[1] += c; o[2] += t; o[3] += w; data[key] = o
tot = sum(v[1] for v in data.values()); smp = sum(v[3] for v in data.values())
print(fn[:100]); print("total warp-inst %.1fM, samples %d" % (tot / 1e6, smp))
for (f, ln), (src, c, t, w) in sorted(data.items()):
    if c > tot * thr / 100 or w > smp * thr / 100:
        print("%-12s %5d %7.2fM %5.1f%% lanes %4.1f stall %4.1f%% | %s" % (f, ln, c / 1e6, 100 * c / tot, t / max(c, 1), 100 * w / max(smp, 1), src.strip()[:100]))
