#!/bin/bash
# round 2, GPU call 11: border condition inside the contiguous-axis stage kernel -- parity tests, A/B timing at 1024^3, bench
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 1200 python -m pytest tests -m gpu -x -q -k "border or variants or engine_matches or binding or backend" > gpurun_out/r2_11_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2_11_tests.log
timeout 900 python scripts/gpu_runs/r2_variants.py --only default,separate_z_border_kernel,fp32_default,fp32_separate_z_border_kernel > gpurun_out/r2_11_variants.jsonl 2>&1
cat gpurun_out/r2_11_variants.jsonl | cut -c1-700
timeout 900 python bench.py > gpurun_out/r2_11_bench.json 2> gpurun_out/r2_11_bench.err
echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2_11_bench.json
