set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
free -g | head -2; nproc
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
GCMB_STAGE_IMPL=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_direct.log 2> gpurun_out/bench_direct.err; echo "bench direct rc=$?"
GCMB_MARCH_SEG=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_seg0.log 2> gpurun_out/bench_seg0.err; echo "bench seg0 rc=$?"
GCMB_MARCH_SEG=64 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_seg64.log 2> gpurun_out/bench_seg64.err; echo "bench seg64 rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --size 512 > gpurun_out/plain512.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --size 512 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --size 384 > gpurun_out/plain384.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_stage -s 9 -c 3 -o gpurun_out/prof_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --size 384 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
cat gpurun_out/bench.log
