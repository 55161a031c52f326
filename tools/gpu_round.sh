#!/bin/bash
# One GPU visit: parity tests, smoke, both bench arms, ncu launch list + full capture of the top kernels.
# usage (from the repo root on the GPU box):  bash tools/gpu_round.sh <tag> [pytest -k expression]
TAG=${1:-r01}
KEXPR=${2:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
python -c "import os; print('cpu cores', os.cpu_count())" >> $OUT/gpu.txt
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -m gpu -q -k "$KEXPR" > $OUT/pytest_gpu.log 2>&1
else
  timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1
fi
tail -5 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; tail -c 600 $OUT/bench_reference.json
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; tail -c 3000 $OUT/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'render_|preprocess_' -s 12 -c 6 \
    -o $OUT/prof_kernels python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT
