python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu4.log
for v in 0 1 2; do GCMB_MARCH_VARIANT=$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench4_v$v.log 2> gpurun_out/bench4_v$v.err; echo "bench v$v rc=$?"; done
GCMB_MARCH_VARIANT=2 GCMB_MARCH_SEG=128 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench4_v2s128.log 2>&1
for f in bench4_v0 bench4_v1 bench4_v2 bench4_v2s128; do python - <<PY
import json
d=json.loads(open('gpurun_out/$f.log').read().strip().splitlines()[-1])
print('$f', '%.3e'%d['value'], round(d['ms_per_step'],2), {k[-8:]:round(v,2) for k,v in d['roofline']['per_stage_ms'].items()}, 'e2e %.3e'%d['e2e']['value'], d['clocks'])
PY
done
