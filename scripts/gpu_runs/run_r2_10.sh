#!/bin/bash
# round 2, GPU call 10: launch list + DRAM traffic of the FINAL binary's bench command at 1024^3, full ncu capture of the three stage kernels at 512^3
cd "$GRAFT_REPO_ROOT" || exit 1
B="python bench.py --steps 2 --warmup 3 --no-sections --no-cpu-baseline"
$B > gpurun_out/r2_ncu_plain_final.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_traffic_1024_final.csv $B > gpurun_out/r2_ncu_l_final.log 2>&1
echo "ncu launches rc=$?"
$B --size 512 > gpurun_out/r2_ncu_plain512_final.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_stage|k_border" -s 9 -c 5 -o gpurun_out/r2_prof_512_final $B --size 512 > gpurun_out/r2_ncu_f_final.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/*final* | tail -8
