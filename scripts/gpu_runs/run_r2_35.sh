#!/bin/bash
# round 2, GPU call 35: the committed final binary -- full GPU suite, smoke, headline bench
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_35_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2_35_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_35_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_35_smoke.log
timeout 600 python bench.py --no-sections --no-cpu-baseline > gpurun_out/r2_35_bench.json 2> gpurun_out/r2_35_bench.err; echo "bench rc=$?"; wc -l gpurun_out/r2_35_bench.json; cut -c1-240 gpurun_out/r2_35_bench.json
