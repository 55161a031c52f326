#!/bin/bash
# ncu evidence for the final tree: per workload one profiled step (all kernels of the library)
#   launches_<w>.csv   gpu__time_duration per launch (shares of the step)
#   full_<w>.ncu-rep   --set full of every kernel of the step (-> tools/ncu_summary.py -> profiles/*.csv, ncu_traffic.json)
# usage: bash tools/gpu_profile_r02.sh <tag> "<workloads>"
TAG=${1:-r02prof}; WLS=${2:-"C3b C4 C2"}
OUT=gpurun_out/$TAG; mkdir -p $OUT
for w in $WLS; do
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file $OUT/launches_$w.csv python tools/profile_step.py $w > $OUT/launches_$w.log 2>&1
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -o $OUT/full_$w \
      python tools/profile_step.py $w > $OUT/full_$w.log 2>&1
  tail -1 $OUT/full_$w.log
done
ls -la $OUT
