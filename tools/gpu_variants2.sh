#!/bin/bash
# A/B library variants on arbitrary workloads: bash tools/gpu_variants2.sh "C2 C5"
for so in stopthepop-rasterization_b200/lib/var/libstp_*.so; do
  for w in $1; do
    STP_RASTERIZER_LIB=$PWD/$so timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > /tmp/b.json 2>/tmp/b.err
    python - <<PY
import json
try:
    b=json.loads(open("/tmp/b.json").read().strip().splitlines()[-1])
    st=b["roofline"]["stages"]; print("$so".split("libstp_")[1], "$w", round(b["ms_per_step"],3), {k: round(v["ms"],3) for k,v in st.items()})
except Exception as e:
    print("$so $w failed", e, open("/tmp/b.err").read()[-300:])
PY
  done
done
