#!/bin/bash
# round 2, GPU call 17: border fill deferred inside the library (gcmb_cubic_border_apply -> next gcmb_cubic_stage): full GPU suite, bench, smoke
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_17_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2_17_tests.log
timeout 900 python bench.py > gpurun_out/r2_17_bench.json 2> gpurun_out/r2_17_bench.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/r2_17_bench.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_17_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_17_smoke.log
