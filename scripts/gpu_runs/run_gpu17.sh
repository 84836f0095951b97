set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-simplex --size 512 > gpurun_out/plain512b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_border -s 4 -c 2 -o gpurun_out/prof_border python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-simplex --size 512 > gpurun_out/ncu_border.log 2>&1; echo "ncu rc=$?"
