#!/bin/bash
# round 2, GPU call 12: marching kernels with the newest window plane left in the ring slot (no spills at 80 registers);
# bulk-copy marching kernels at 6 blocks/SM again.  Parity of every variant, then A/B timing at 1024^3.
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 1500 python -m pytest tests -m gpu -x -q -k "variants or border or engine_matches or random" > gpurun_out/r2_12_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2_12_tests.log
timeout 900 python scripts/gpu_runs/r2_variants.py --only default,tma_warp_pipes_separate_border,ldgsts_courant1,tma_courant1,fp32_default,tma_fp32 > gpurun_out/r2_12_variants.jsonl 2>&1
cut -c1-620 gpurun_out/r2_12_variants.jsonl
