#!/bin/bash
# round 2, GPU call 20: simplex -- direction masks and stored gradient geometry: parity (all simplex GPU tests), then timing on/off
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 1500 python -m pytest tests -m gpu -x -q -k "simplex or locate" > gpurun_out/r2_20_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2_20_tests.log
timeout 1200 python scripts/gpu_runs/r2_simplex_dir_masks.py > gpurun_out/r2_simplex_dir_masks.log 2>&1
cut -c1-420 gpurun_out/r2_simplex_dir_masks.log
