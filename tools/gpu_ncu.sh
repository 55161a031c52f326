#!/bin/bash
# ncu launch list (per-launch durations) + optional full capture of selected kernels for one workload.
# usage: bash tools/gpu_ncu.sh <tag> <workload> [kernel-regex for --set full] [launch-skip] [launch-count]
TAG=${1:-ncu}
WL=${2:-C2}
KRE=${3:-}
SKIP=${4:-0}
CNT=${5:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$WL.csv \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches_$WL.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("$OUT/launches_$WL.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(",",""))
    except: continue
    k=r[ki][:90]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{a[1]/a[0]/1e3:10.1f} us x{a[0]:3d} {100*a[1]/tot:5.1f}%  {k}")
PY
if [ -n "$KRE" ]; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c $CNT \
      -o $OUT/full_$WL python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$WL.log 2>&1
  ls -la $OUT
fi
