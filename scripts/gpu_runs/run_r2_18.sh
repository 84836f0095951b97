#!/bin/bash
# round 2, GPU call 18 (1 GPU): ncu launch list + DRAM traffic of the final binary's bench command at 1024^3 (the ghost fill is
# now inside the z-stage kernel), full capture of the stage kernels at 512^3
cd "$GRAFT_REPO_ROOT" || exit 1
B="python bench.py --steps 2 --warmup 3 --no-sections --no-cpu-baseline"
$B > gpurun_out/r2_18_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_traffic_1024_final2.csv $B > gpurun_out/r2_18_ncu_l.log 2>&1
echo "ncu launches rc=$?"
$B --size 512 > gpurun_out/r2_18_plain512.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_stage" -s 9 -c 3 -o gpurun_out/r2_prof_512_final2 $B --size 512 > gpurun_out/r2_18_ncu_f.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/*final2* | tail -4
