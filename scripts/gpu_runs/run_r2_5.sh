#!/bin/bash
# round 2, GPU call 5 (2 GPUs): decomposed runs against the fixtures, the 2-GPU bench line, and 1-GPU re-measurements
cd "$GRAFT_REPO_ROOT" || exit 1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 tests/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check_2gpu.log 2>&1; echo "multi_gpu_check rc=$?"; tail -4 gpurun_out/r2_multi_gpu_check_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench2 rc=$?"; tail -c 400 gpurun_out/r2_bench_2gpu.err; cut -c1-700 gpurun_out/r2_bench_2gpu.json
python -m pytest tests -m gpu -x -q -k "launcher or variants or golden or bitwise" > gpurun_out/r2_gputest4.log 2>&1; tail -3 gpurun_out/r2_gputest4.log
python scripts/gpu_runs/r2_variants.py --only default > gpurun_out/r2_variants3.log 2>&1; cut -c1-420 gpurun_out/r2_variants3.log
GCMB_BORDER_LAYERWISE=1 python scripts/gpu_runs/r2_variants.py --only default >> gpurun_out/r2_variants3.log 2>&1; tail -1 gpurun_out/r2_variants3.log | cut -c1-420
