#!/bin/bash
# compute-sanitizer over a cross-section of the GPU tests (all sort modes, forward + backward, slabs / TMA ring, async
# forward, debug visualisations, ragged sizes): memcheck, and racecheck for shared-memory hazards
OUT=gpurun_out/${1:-r02san}; mkdir -p $OUT
K='(fixture and (hier_default or hier_preset or hier_long or global_default or kbuffer16 or full_sort)) or ragged or async or debug_vis or head12 or (queue_matrix and head4_mid8)'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests -m gpu -q -x -k "$K" > $OUT/compute_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -3 $OUT/compute_sanitizer_memcheck.log
K2='(fixture and (hier_default or hier_preset or global_default or kbuffer16 or full_sort)) and forward'
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests -m gpu -q -x -k "$K2" > $OUT/compute_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -3 $OUT/compute_sanitizer_racecheck.log
