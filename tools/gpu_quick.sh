#!/bin/bash
# Quick GPU visit: parity tests (optionally filtered) + our bench arm on the listed workloads.
# usage: bash tools/gpu_quick.sh <tag> "<pytest -k expr or empty>" "<workloads, e.g. C2 C3a C3b>"
TAG=${1:-quick}
KEXPR=${2:-}
WLS=${3:-C2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -n "$KEXPR" ]; then
  timeout 1200 python -m pytest tests -m gpu -q -x -k "$KEXPR" > $OUT/pytest.log 2>&1
else
  timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1
fi
tail -15 $OUT/pytest.log
for w in $WLS; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_$w.json 2> $OUT/bench_$w.err
  python - <<PY
import json
try:
    b=json.loads(open("$OUT/bench_$w.json").read().strip().splitlines()[-1])
    print("$w", "ms/step", round(b["ms_per_step"],3), "e2e ms", round(b["e2e"]["ms_per_step"],3), {k: round(v["ms"],3) for k,v in b["roofline"]["stages"].items()})
except Exception as e:
    print("$w bench failed", e); print(open("$OUT/bench_$w.err").read()[-2000:])
PY
done
