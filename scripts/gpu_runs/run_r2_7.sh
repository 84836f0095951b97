#!/bin/bash
# round 2, GPU call 7: occupancy knobs of the simplex kernels (experiment build), measured bounds of fp32 / FMA, marching-segment knob
cd "$GRAFT_REPO_ROOT" || exit 1
python scripts/gpu_runs/r2_simplex_knobs.py > gpurun_out/r2_simplex_knobs.log 2>&1; cut -c1-420 gpurun_out/r2_simplex_knobs.log
python scripts/gpu_runs/r2_bounds.py > gpurun_out/r2_bounds.log 2>&1; cat gpurun_out/r2_bounds.log | cut -c1-300
for seg in 128 192; do GCMB_MARCH_SEG=$seg python scripts/gpu_runs/r2_variants.py --only default 2>&1 | cut -c1-330; done > gpurun_out/r2_variants4.log; cat gpurun_out/r2_variants4.log
python -m pytest tests -m gpu -x -q -k "many_materials or vtk or launcher" > gpurun_out/r2_gputest5.log 2>&1; tail -3 gpurun_out/r2_gputest5.log
