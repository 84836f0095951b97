#!/bin/bash
# round 2, GPU call 6 (8 GPUs): decomposed runs against the fixtures and the 8-GPU bench line (headline + config 4)
cd "$GRAFT_REPO_ROOT" || exit 1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 tests/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check_8gpu.log 2>&1; echo "multi_gpu_check rc=$?"; grep -c "bitwise" gpurun_out/r2_multi_gpu_check_8gpu.log; tail -2 gpurun_out/r2_multi_gpu_check_8gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; echo "bench8 rc=$?"; tail -c 300 gpurun_out/r2_bench_8gpu.err; cut -c1-400 gpurun_out/r2_bench_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err; echo "bench4 rc=$?"; cut -c1-300 gpurun_out/r2_bench_4gpu.json
