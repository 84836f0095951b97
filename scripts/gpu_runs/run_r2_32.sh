#!/bin/bash
# round 2, GPU call 32: final binary (deep bulk-copy marching kernels, warps past the row end retire) -- full GPU suite, smoke,
# bench, then the ncu launch list with DRAM traffic at 1024^3 and a full capture of the stage kernels at 512^3
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_32_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2_32_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_32_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_32_smoke.log
timeout 900 python bench.py > gpurun_out/r2_32_bench.json 2> gpurun_out/r2_32_bench.err
echo "bench rc=$?"; cut -c1-260 gpurun_out/r2_32_bench.json
for n in 64 128; do timeout 300 python scripts/gpu_runs/r2_variants.py --size $n --steps 100 --only default | cut -c1-330; done
B="python bench.py --steps 2 --warmup 3 --no-sections --no-cpu-baseline"
$B > gpurun_out/r2_32_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_traffic_1024_final3.csv $B > gpurun_out/r2_32_ncu_l.log 2>&1
echo "ncu launches rc=$?"
$B --size 512 > gpurun_out/r2_32_plain512.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_stage" -s 9 -c 3 -o gpurun_out/r2_prof_512_final3 $B --size 512 > gpurun_out/r2_32_ncu_f.log 2>&1
echo "ncu full rc=$?"
