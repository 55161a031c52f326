#!/bin/bash
# ncu evidence for the final tree: per workload one profiled step (all kernels of the library)
#   launches_<w>.csv        gpu__time_duration per launch (shares of the step)
#   ncu_full_summary_<w>.csv  one line per kernel from a --set full capture (tools/ncu_summary.py), ncu_traffic.json
#   src_<w>_<kernel>.txt    hottest source lines of the dominant kernels (tools/ncu_src.py)
# The .ncu-rep files are summarised ON the box and deleted (gpurun_out is limited to 64 MiB).
# usage: bash tools/gpu_profile_r02.sh <tag> "<workloads>"
TAG=${1:-r02prof}; WLS=${2:-"C3b C4 C2"}
OUT=gpurun_out/$TAG; mkdir -p $OUT
cp profiles/ncu_traffic.json $OUT/ncu_traffic.json 2>/dev/null
for w in $WLS; do
  timeout ${T_LIST:-600} ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file $OUT/launches_$w.csv python tools/profile_step.py $w > $OUT/launches_$w.log 2>&1
  timeout ${T_FULL:-900} ncu --profile-from-start off --set full --clock-control none --import-source on -o /tmp/full_$w \
      python tools/profile_step.py $w > $OUT/full_$w.log 2>&1
  python tools/ncu_summary.py /tmp/full_$w.ncu-rep $OUT/ncu_full_summary_$w.csv $w $OUT/ncu_traffic.json > /dev/null 2>&1
  for k in render_hier_kernel blend_replay_bwd render_full_fast render_global_bwd; do
    python tools/ncu_src.py /tmp/full_$w.ncu-rep $k 45 > $OUT/src_${w}_$k.txt 2>/dev/null
    [ $(wc -l < $OUT/src_${w}_$k.txt) -lt 5 ] && rm -f $OUT/src_${w}_$k.txt
  done
  rm -f /tmp/full_$w.ncu-rep
  cut -d, -f1-3,9,15 $OUT/ncu_full_summary_$w.csv | cut -c1-200
done
cuobjdump -sass stopthepop-rasterization_b200/lib/libstp_rasterizer.so 2>/dev/null | grep -E "Function :|UBLKCP|SYNCS|UTMALDG" | grep -B1 -E "UBLKCP|SYNCS" | grep -v "^--" > $OUT/sass_tma_evidence.txt
ls -la $OUT
