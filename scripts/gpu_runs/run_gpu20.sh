set -x
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-simplex > gpurun_out/bench_r20.log 2>&1
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r20.log').read().strip().splitlines()[-1])
print('%.4e'%d['value'], round(d['ms_per_step'],2), {k[-6:]:round(v,2) for k,v in d['roofline']['per_stage_ms'].items()})
PY
