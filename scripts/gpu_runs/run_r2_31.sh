#!/bin/bash
# round 2, GPU call 31: deep bulk-copy marching kernels as the default for the 9-component fp64 patterns -- small sizes against
# the cp.async kernels, full GPU suite, bench, smoke
cd "$GRAFT_REPO_ROOT" || exit 1
for n in 64 128 256 512; do
for impl in default 2; do
if [ $impl = default ]; then unset GCMB_STAGE_IMPL; else export GCMB_STAGE_IMPL=$impl; fi
timeout 300 python scripts/gpu_runs/r2_variants.py --size $n --steps 100 --only default | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('VARIANT '):
        d = json.loads(line[8:])
        print(d['n'], 'impl $impl', 'ms/step %.4f' % d['ms_per_step'], 'stages', ['%.4f' % x for x in d['stage_ms']])
"
done; done
unset GCMB_STAGE_IMPL
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_31_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2_31_tests.log
timeout 900 python bench.py > gpurun_out/r2_31_bench.json 2> gpurun_out/r2_31_bench.err
echo "bench rc=$?"; cut -c1-260 gpurun_out/r2_31_bench.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_31_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_31_smoke.log
