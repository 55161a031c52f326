#!/bin/bash
# view-sharded bench (ours only, no extras) for several pipelining depths of the parameter-gradient exchange
N=${1:-4}; W=${2:-C5}; OUT=gpurun_out/${3:-r02views}; mkdir -p $OUT
for c in 1 2 4; do
  STP_VIEW_SYNC_CHUNKS=$c timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 3 --no-extras --workload $W > $OUT/${W}_n${N}_c$c.json 2> $OUT/${W}_n${N}_c$c.err
  python - <<PY
import json
try:
    b=json.loads(open("$OUT/${W}_n${N}_c$c.json").read().strip().splitlines()[-1])
    print("$W N=$N chunks=$c ms/step", round(b["ms_per_step"],3), "Mpix/s", round(b["value"],1), "e2e ms", round(b["e2e"]["ms_per_step"],3), {k: round(v["ms"],3) for k,v in b["roofline"]["stages"].items()})
except Exception as e:
    print("failed", e, open("$OUT/${W}_n${N}_c$c.err").read()[-1500:])
PY
done
