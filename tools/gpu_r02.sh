#!/bin/bash
# Round-2 GPU visit: parity tests, selected benches, optional ncu capture of one kernel with source.
# usage: bash tools/gpu_r02.sh <tag> "<workloads to bench>" ["<workload>:<kernel regex (demangled)>" ...]
TAG=${1:-r02}
WLS=${2:-}
shift; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
tail -4 $OUT/pytest.log
for w in $WLS; do
  S=20; [ $w == C4 ] && S=5
  timeout 500 python bench.py --workload $w --steps $S --warmup 3 --no-cpu-baseline > $OUT/bench_$w.json 2> $OUT/bench_$w.err
  python - <<PY
import json
try:
    b=json.loads(open("$OUT/bench_$w.json").read().strip().splitlines()[-1])
    print("$w", "ms/step", round(b["ms_per_step"],3), "e2e ms", round(b["e2e"]["ms_per_step"],3), {k: round(v["ms"],3) for k,v in b["roofline"]["stages"].items()})
except Exception as e:
    print("$w bench failed", e); print(open("$OUT/bench_$w.err").read()[-1500:])
PY
done
for spec in "$@"; do
  w=${spec%%:*}; kre=${spec#*:}
  name=$(echo "$kre" | tr -c 'a-zA-Z0-9' '_' | cut -c1-40)
  timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$kre" -s 3 -c 1 \
      -o $OUT/ncu_${w}_$name python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_${w}_$name.log 2>&1
  tail -2 $OUT/ncu_${w}_$name.log
done
ls -la $OUT
