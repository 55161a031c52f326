#!/usr/bin/env python
"""Per-source-line instruction / stall summary of one kernel in an .ncu-rep (needs --import-source on, -lineinfo).
usage: python tools/ncu_src.py <report.ncu-rep> <kernel-regex> [top-N]"""
import csv, io, subprocess, sys
rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kname,
                      "-c", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines, tot_i, tot_s, cur_file = [], 0, 0, ""
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        ii, si = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit():
        continue
    try:
        ie, ss = int(r[ii]), int(r[si])
    except ValueError:
        continue
    lines.append((ie, ss, cur_file, r[0], r[1].strip()))
    tot_i += ie
    tot_s += ss
print(f"total warp-instr {tot_i}, stall samples {tot_s}")
for ie, ss, fn, ln, src in sorted(lines, key=lambda t: -t[0])[:top]:
    print(f"{100*ie/max(tot_i,1):5.1f}% instr {100*ss/max(tot_s,1):5.1f}% stall  {fn}:{ln:>4} {src[:100]}")
