#!/usr/bin/env python
"""Instruction accounting of one kernel by source regions / marker lines (ncu source page).
usage: python tools/ncu_regions.py <rep> <kernel-regex> <file> "name:a-b,name:a-b,..." "markerline,markerline,..." """
import csv, io, subprocess, sys
rep, kre, fname, regs = sys.argv[1:5]
marks = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 and sys.argv[5] else []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kre, "-c", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, cur, data = None, "", {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        ii, ti = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit():
        continue
    try:
        data[(cur, int(r[0]))] = (int(r[ii]), int(r[ti]), r[1].strip())
    except ValueError:
        pass
tot = sum(v[0] for v in data.values())
print("total warp-instr", tot)
for spec in regs.split(","):
    name, ab = spec.split(":")
    a, b = [int(x) for x in ab.split("-")]
    s = sum(v[0] for (f, l), v in data.items() if f == fname and a <= l <= b)
    t = sum(v[1] for (f, l), v in data.items() if f == fname and a <= l <= b)
    print(f"{name:28s} {100*s/tot:5.1f}%   lanes/instr {t/max(s,1):5.1f}")
files = sorted({f for f, _ in data})
for f in files:
    s = sum(v[0] for (ff, l), v in data.items() if ff == f)
    print(f"file {f:30s} {100*s/tot:5.1f}%")
for (f, l), v in sorted(data.items()):
    if f != fname and v[0] > tot * 0.004:
        print(f"  {f}:{l} {100*v[0]/tot:5.2f}%  {v[2][:100]}")
for l in marks:
    print("marker", l, data.get((fname, l)))
