#!/bin/bash
# N-GPU tile-band bench (ours only, no extras) for several pipelining depths of the accumulator exchange
N=${1:-8}; OUT=gpurun_out/${2:-r02chunks}; mkdir -p $OUT
for c in 1 2 4; do
  STP_BAND_SYNC_CHUNKS=$c timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 40 --warmup 3 --no-extras > $OUT/n${N}_c$c.json 2> $OUT/n${N}_c$c.err
  python - <<PY
import json
try:
    b=json.loads(open("$OUT/n${N}_c$c.json").read().strip().splitlines()[-1])
    print("N=$N chunks=$c ms/step", round(b["ms_per_step"],3), "e2e ms", round(b["e2e"]["ms_per_step"],3), {k: round(v["ms"],3) for k,v in b["roofline"]["stages"].items()})
except Exception as e:
    print("failed", e, open("$OUT/n${N}_c$c.err").read()[-1500:])
PY
done
