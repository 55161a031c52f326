#!/bin/bash
# default bench, OUR arm only, at N GPUs (the reference arm of a tile-band run is the 1-GPU reference, measured separately)
TAG=$1; N=$2; K=${3:-20}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps $K --warmup 3 > $OUT/bench_n${N}_ours.json 2> $OUT/bench_n${N}_ours.err
python - <<PY
import json
try:
    b=json.loads(open("$OUT/bench_n${N}_ours.json").read().strip().splitlines()[-1])
    print("N=$N ours ms/step", round(b["ms_per_step"],3), "Mpix/s", round(b["value"],1), "e2e", round(b["e2e"]["value"],1), {k: round(v["ms"],3) for k,v in b["roofline"]["stages"].items()})
    for w, r in b.get("extra", {}).items():
        print("   extra", w, "ms/step", round(r["ms_per_step"],3), "Mpix/s", round(r["value"],1), "e2e", round(r["e2e"]["value"],1))
except Exception as e:
    print("failed", e); print(open("$OUT/bench_n${N}_ours.err").read()[-2000:])
PY
