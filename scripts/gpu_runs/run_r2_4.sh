#!/bin/bash
# round 2, GPU call 4: the GPU suite, the full bench line, launch list and DRAM traffic at 1024^3, full ncu capture at 512^3
cd "$GRAFT_REPO_ROOT" || exit 1
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest3.log 2>&1; tail -3 gpurun_out/r2_gputest3.log
python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2_bench_1gpu.err
B="python bench.py --steps 2 --warmup 3 --no-sections --no-cpu-baseline"
$B > gpurun_out/r2_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_traffic_1024.csv $B > gpurun_out/r2_ncu_l.log 2>&1
echo "ncu launches rc=$?"
$B --size 512 > gpurun_out/r2_ncu_plain512.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_stage -s 9 -c 3 -o gpurun_out/r2_prof_512 $B --size 512 > gpurun_out/r2_ncu_f.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -8
