#!/bin/bash
# round 2, GPU call 22-23: small grids -- occupancy of the stage launches (segment length, rows per block)
cd "$GRAFT_REPO_ROOT" || exit 1
for n in 32 64 128 256 512; do
timeout 300 python scripts/gpu_runs/r2_variants.py --size $n --steps 200 --only default | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('VARIANT '):
        d = json.loads(line[8:])
        print(d['n'], 'ms/step %.4f' % d['ms_per_step'], 'stage kernels ms', ['%.4f' % x for x in d['stage_ms']], 'launches/step', d['launches_per_step'], 'node-updates/s %.3e' % d['node_updates_per_s'])
"
done
timeout 600 python scripts/gpu_runs/r2_variants.py --size 1024 --steps 5 --only default | cut -c1-420
timeout 900 python -m pytest tests -m gpu -x -q -k "variants or engine_matches or random or anchor" 2>&1 | tail -2
