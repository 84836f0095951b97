set -x
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tests/multi_gpu_check.py > gpurun_out/multi_gpu_check_${N}gpu_c.log 2>&1; echo "multi check rc=$?"; grep "MULTI_GPU_OK\|bitwise" gpurun_out/multi_gpu_check_${N}gpu_c.log | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_c.log 2> gpurun_out/bench_${N}gpu_c.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_${N}gpu_c.log | cut -c1-330
