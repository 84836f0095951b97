#!/bin/bash
# round 2, GPU call 9 (8 GPUs): final binaries -- GPU suite, smoke, bench at 1 / 2 / 4 / 8 GPUs, decomposed runs against the fixtures
cd "$GRAFT_REPO_ROOT" || exit 1
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_final.log 2>&1; tail -3 gpurun_out/r2_gputest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/r2_bench_final_1gpu.json 2> gpurun_out/r2_bench_final_1gpu.err; echo "bench1 rc=$?"; cut -c1-260 gpurun_out/r2_bench_final_1gpu.json
for n in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_bench_final_${n}gpu.json 2> gpurun_out/r2_bench_final_${n}gpu.err; echo "bench$n rc=$?"; cut -c1-260 gpurun_out/r2_bench_final_${n}gpu.json
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 tests/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check_final_8gpu.log 2>&1; echo "multi_gpu_check rc=$?"; tail -1 gpurun_out/r2_multi_gpu_check_final_8gpu.log
