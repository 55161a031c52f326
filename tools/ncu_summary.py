#!/usr/bin/env python
"""Compact per-launch summary CSV of an .ncu-rep (the file committed under profiles/) and, optionally, the measured
DRAM traffic per stage for bench.py's roofline.traffic (profiles/ncu_traffic.json).
usage: python tools/ncu_summary.py <rep> <out.csv> [<workload> <traffic.json>]"""
import csv, io, json, os, subprocess, sys
rep, out_csv = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
cols = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed"]
idx = [(c, hdr.index(c)) for c in cols if c in hdr]
def to_bytes(v, u):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
with open(out_csv, "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow([c + (f" [{units[i]}]" if units[i] else "") for c, i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for _, i in idx])
print(open(out_csv).read()[:3000])
if len(sys.argv) > 4:
    wl, tj = sys.argv[3], sys.argv[4]
    # (order matters: the first match wins -- preprocess_bwd before preprocess_kernel is not needed, names differ)
    stage_of = {"preprocess_kernel": "Preprocess", "tile_scan": "Preprocess", "duplicate_kernel": "Duplicate",
                "tile_sort": "Sort", "render_global_fwd": "Render", "render_global_bwd": "RenderBackward",
                "render_hier_kernel": "Render", "blend_replay_bwd": "RenderBackward", "render_full": "Render",
                "render_kbuffer": "Render", "preprocess_bwd": "PreprocessBackward"}
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    ii, ti = hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active"), hdr.index("gpu__time_duration.sum")
    acc, cnt, issue = {}, {}, {}
    for r in rows[2:]:
        for k, st in stage_of.items():
            if k in r[ki]:
                acc.setdefault(st, {})
                kk = r[ki][:60]
                a = acc[st].setdefault(kk, [0.0, 0])
                a[0] += to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi])
                a[1] += 1
                dur = float(r[ti].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[ti], 1.0)
                if dur >= issue.get(st, (0.0, 0.0))[0]:  # the longest kernel of the stage
                    issue[st] = (dur, float(r[ii]))
                break
    traffic = {st: sum(v[0] / v[1] for v in d.values()) for st, d in acc.items()}
    data = {}
    if os.path.exists(tj):
        data = json.load(open(tj))
    data[wl] = {k: round(v) for k, v in traffic.items()}
    data[wl + "_issue_active_pct"] = {k: round(v[1], 1) for k, v in issue.items()}
    json.dump(data, open(tj, "w"), indent=1, sort_keys=True)
    print(json.dumps(data[wl]))
