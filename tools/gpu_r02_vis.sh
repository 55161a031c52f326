#!/bin/bash
# one short GPU call: golden render_depth images of the reference build + the debug-visualisation parity tests
set -x
mkdir -p gpurun_out/golden
timeout 300 python tests/golden/make_golden_depth_vis.py > gpurun_out/golden/make_golden_depth_vis.log 2>&1
tail -3 gpurun_out/golden/make_golden_depth_vis.log
cp gpurun_out/golden/depth_vis.npz tests/golden/ 2>/dev/null
timeout 300 python -m pytest tests/test_oracle_golden.py -q -k depth_visualisation 2>&1 | tail -15
timeout 400 python -m pytest tests/test_gpu_parity.py -q -k debug_visualisation 2>&1 | tail -40
