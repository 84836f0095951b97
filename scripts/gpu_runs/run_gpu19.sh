set -x
for g in 32 64 128; do GCMB_L2_FETCH=$g python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-simplex > gpurun_out/bench_l2_$g.log 2>&1; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_l2_$g.log').read().strip().splitlines()[-1])
print('L2 fetch $g', '%.4e'%d['value'], round(d['ms_per_step'],2), {k[-6:]:round(v,2) for k,v in d['roofline']['per_stage_ms'].items()})
PY
done
