python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu5.log
for v in 0 1 2 3; do GCMB_MARCH_VARIANT=$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench5_v$v.log 2> gpurun_out/bench5_v$v.err; done
GCMB_ZTILE_VARIANT=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench5_z1.log 2>&1
GCMB_MARCH_SEG=128 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench5_s128.log 2>&1
GCMB_MARCH_SEG=512 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench5_s512.log 2>&1
for f in bench5_v0 bench5_v1 bench5_v2 bench5_v3 bench5_z1 bench5_s128 bench5_s512; do python - <<PY
import json
d=json.loads(open('gpurun_out/$f.log').read().strip().splitlines()[-1])
print('$f', '%.3e'%d['value'], round(d['ms_per_step'],2), {k[-8:]:round(v,2) for k,v in d['roofline']['per_stage_ms'].items()}, 'e2e %.3e'%d['e2e']['value'], d['clocks'])
PY
done
