set -x
python -m pytest tests -m gpu -q -x --timeout 900 -k simplex > gpurun_out/pytest_simplex2.log 2>&1; echo "pytest simplex rc=$?"; tail -3 gpurun_out/pytest_simplex2.log
timeout 900 python tests/simplex_perf.py 96 3 > gpurun_out/simplex_perf96b.log 2>&1; echo rc=$?; cat gpurun_out/simplex_perf96b.log
timeout 600 python tests/simplex_perf.py 64 2 > gpurun_out/simplex_perf64.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_s_inner_nodes|k_s_border_nodes|k_s_gradient" -s 12 -c 6 -o gpurun_out/prof_simplex_r2 python tests/simplex_perf.py 64 2 > gpurun_out/ncu_simplex.log 2>&1; echo "ncu rc=$?"
