#!/bin/bash
# default bench, both arms, at N GPUs (as the driver launches it): bash tools/gpu_bench_n.sh <tag> <N> [steps] [extra bench args]
TAG=$1; N=$2; K=${3:-20}; shift; shift; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for impl in reference ours; do
  if [ "$N" == "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps $K --warmup 3 --impl $impl "$@" > $OUT/bench_n${N}_$impl.json 2> $OUT/bench_n${N}_$impl.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $N --steps $K --warmup 3 --impl $impl "$@" > $OUT/bench_n${N}_$impl.json 2> $OUT/bench_n${N}_$impl.err
  fi
  python - <<PY
import json
try:
    b=json.loads(open("$OUT/bench_n${N}_$impl.json").read().strip().splitlines()[-1])
    print("N=$N $impl", b["config"]["workload"][:40], "ms/step", round(b["ms_per_step"],3), "Mpix/s", round(b["value"],1), "e2e", round(b["e2e"]["value"],1), b.get("roofline",{}).get("stages") and {k: round(v["ms"],3) for k,v in b["roofline"]["stages"].items()})
    for w, r in b.get("extra", {}).items():
        print("   extra", w, "ms/step", round(r["ms_per_step"],3), "Mpix/s", round(r["value"],1), "e2e", round(r["e2e"]["value"],1), r.get("stages_ms"))
except Exception as e:
    print("N=$N $impl failed", e); print(open("$OUT/bench_n${N}_$impl.err").read()[-2500:])
PY
done
