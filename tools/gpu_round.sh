#!/bin/bash
# One GPU visit: parity tests, smoke, both bench arms on C2 / C3a / C3b, ncu launch lists + full captures of the top kernels.
# usage (from the repo root on the GPU box):  bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
python -c "import os; print('cpu cores', os.cpu_count())" >> $OUT/gpu.txt
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1
tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
for w in C2 C3a C3b C5 C4; do
  S=100; [ $w != C2 ] && S=20; [ $w == C4 ] && S=3
  timeout 600 python bench.py --workload $w --impl reference --steps $S --warmup 5 > $OUT/bench_${w}_reference.json 2> $OUT/bench_${w}_reference.err
  NOCPU="--no-cpu-baseline"; [ $w == C2 ] && NOCPU=""
  timeout 600 python bench.py --workload $w --steps $S --warmup 5 $NOCPU > $OUT/bench_$w.json 2> $OUT/bench_$w.err
  python - <<PY
import json
for arm in ("_reference", ""):
    try:
        b=json.loads(open("$OUT/bench_$w%s.json" % arm).read().strip().splitlines()[-1])
        st = {k: round(v["ms"],3) for k,v in b.get("roofline",{}).get("stages",{}).items()}
        print("$w", arm or "ours", "ms/step", round(b["ms_per_step"],3), "Mpix/s", round(b["value"],1), "e2e", round(b["e2e"]["value"],1), st)
    except Exception as e:
        print("$w", arm, "failed", e)
PY
done
for w in C2 C3b C4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$w.csv \
      python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches_$w.log 2>&1
done
# full captures are large (gpurun_out is limited to 64 MiB per call): tools/gpu_ncu.sh <tag> <workload> <kernel-regex> <skip> <count>
ls -la $OUT
