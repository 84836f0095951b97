#!/bin/bash
# round 2, GPU calls 21, 36 and 37 (2 GPUs): final binary -- decomposed runs against the fixtures (incl. the contact across z), bench at 2 GPUs
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check_2gpu.log 2>&1; echo "multi_gpu_check rc=$?"; tail -4 gpurun_out/r2_multi_gpu_check_2gpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_37_bench_2gpu.json 2> gpurun_out/r2_37_bench_2gpu.err; echo "bench2 rc=$?"; cut -c1-300 gpurun_out/r2_37_bench_2gpu.json
