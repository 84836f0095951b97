set -x
timeout 600 python tests/simplex_perf.py 48 5 > gpurun_out/simplex_perf48.log 2>&1; echo rc=$?; cat gpurun_out/simplex_perf48.log
timeout 900 python tests/simplex_perf.py 96 3 > gpurun_out/simplex_perf96.log 2>&1; echo rc=$?; cat gpurun_out/simplex_perf96.log
