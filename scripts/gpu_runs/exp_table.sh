#!/bin/bash
# usage: exp_table.sh variant,variant,...  -- one line per variant of scripts/gpu_runs/r2_variants.py
cd "$GRAFT_REPO_ROOT" || exit 1
python scripts/gpu_runs/r2_variants.py --only "$1" ${2:+--size $2} 2>&1 | tee -a gpurun_out/r2_exp_table_raw.jsonl | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('VARIANT '):
        d = json.loads(line[8:])
        print(d.get('variant'), d['error'][-300:] if 'error' in d else ('%.2f ms/step x %.2f y %.2f z %.2f checksum %.3f' % (d['ms_per_step'], *d['stage_ms'], d['checksum'])))
"
