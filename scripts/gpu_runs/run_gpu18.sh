set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke3.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke3.log
python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu_r18.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r18.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r18.log 2> gpurun_out/bench_r18.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r18.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r18.log 2> gpurun_out/bench_ref_r18.err; echo "bench ref rc=$?"; tail -2 gpurun_out/bench_ref_r18.log
timeout 600 python tests/simplex_perf.py 64 4 > gpurun_out/simplex_perf64b.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_s_" -s 60 -c 12 -o gpurun_out/prof_simplex_r3 python tests/simplex_perf.py 64 4 > gpurun_out/ncu_simplex3.log 2>&1; echo "ncu rc=$?"
