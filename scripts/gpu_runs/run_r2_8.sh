#!/bin/bash
# round 2, GPU call 8 (2 GPUs): whole GPU suite, smoke, the 1-GPU bench line, decomposed runs (async seismogram all-reduce) and the 2-GPU bench line
cd "$GRAFT_REPO_ROOT" || exit 1
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest6.log 2>&1; tail -3 gpurun_out/r2_gputest6.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2_smoke.log | cut -c1-250
python bench.py > gpurun_out/r2_bench_1gpu_b.json 2> gpurun_out/r2_bench_1gpu_b.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/r2_bench_1gpu_b.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 tests/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check_2gpu_b.log 2>&1; echo "multi_gpu_check rc=$?"; tail -2 gpurun_out/r2_multi_gpu_check_2gpu_b.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu_b.json 2> gpurun_out/r2_bench_2gpu_b.err; echo "bench2 rc=$?"; cut -c1-330 gpurun_out/r2_bench_2gpu_b.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>&1; cut -c1-300 gpurun_out/r2_bench_reference_arm.json
