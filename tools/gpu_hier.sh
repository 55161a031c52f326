#!/bin/bash
# HIER iteration loop: parity tests of the hier cases + C3a / C3b benches (ours only)
TAG=${1:-hier}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -k "hier or C2-3 or smoke" > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
for w in C3a C3b; do
  timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_$w.json 2> $OUT/bench_$w.err
  python - <<PY
import json
b=json.loads(open("$OUT/bench_$w.json").read().strip().splitlines()[-1])
print("$w", "ms/step", round(b["ms_per_step"],2), {k: round(v["ms"],2) for k,v in b["roofline"]["stages"].items()})
PY
done
