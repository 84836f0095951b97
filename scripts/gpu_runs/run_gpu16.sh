set -x
python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu_r16.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r16.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r16.log 2> gpurun_out/bench_r16.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r16.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r16.log').read().strip().splitlines()[-1])
print('%.4e'%d['value'], round(d['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), 'whole', round(d['roofline']['whole_step_frac'],3), 'e2e %.4e'%d['e2e']['value'], d['roofline']['per_stage_ms'])
print(d['simplex']['value'], d['simplex']['ms_per_step'])
PY
