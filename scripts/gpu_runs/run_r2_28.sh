#!/bin/bash
# round 2, GPU call 28 (8 GPUs): final binary -- bench at 8 GPUs, decomposed runs against the fixtures on 8 GPUs
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_28_bench_8gpu.json 2> gpurun_out/r2_28_bench_8gpu.err; echo "bench8 rc=$?"; wc -l gpurun_out/r2_28_bench_8gpu.json; cut -c1-300 gpurun_out/r2_28_bench_8gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 tests/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check_8gpu.log 2>&1; echo "multi_gpu_check rc=$?"; tail -2 gpurun_out/r2_multi_gpu_check_8gpu.log
