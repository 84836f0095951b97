#!/bin/bash
# round 2, GPU call 27: segment length of the marching launches at 256^3 and 512^3 (tail of the last wave against the re-read planes)
cd "$GRAFT_REPO_ROOT" || exit 1
for n in 256 512; do for seg in auto 32 64 128 256; do
if [ $seg = auto ]; then unset GCMB_MARCH_SEG; else export GCMB_MARCH_SEG=$seg; fi
timeout 300 python scripts/gpu_runs/r2_variants.py --size $n --steps 40 --only default | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('VARIANT '):
        d = json.loads(line[8:])
        print(d['n'], 'seg $seg', 'ms/step %.4f' % d['ms_per_step'], 'stages', ['%.4f' % x for x in d['stage_ms']])
"
done; done
