#!/bin/bash
# A/B a set of library variants (stopthepop-rasterization_b200/lib/var/libstp_*.so) on C3a / C3b
for so in stopthepop-rasterization_b200/lib/var/libstp_*.so; do
  for w in C3a C3b; do
    STP_RASTERIZER_LIB=$PWD/$so timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > /tmp/b.json 2>/tmp/b.err
    python - <<PY
import json
try:
    b=json.loads(open("/tmp/b.json").read().strip().splitlines()[-1])
    st=b["roofline"]["stages"]; print("$so".split("libstp_")[1], "$w", "fwd", round(st["Render"]["ms"],2), "bwd", round(st["RenderBackward"]["ms"],2))
except Exception as e:
    print("$so $w failed", e, open("/tmp/b.err").read()[-300:])
PY
  done
done
