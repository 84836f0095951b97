set -x
python -m pytest tests -m gpu -q -x --timeout 900 -k simplex > gpurun_out/pytest_simplex3.log 2>&1; echo "pytest simplex rc=$?"; tail -3 gpurun_out/pytest_simplex3.log
timeout 900 python tests/simplex_perf.py 96 6 2>&1 | tee gpurun_out/simplex_perf96_c.log
