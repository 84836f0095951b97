set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r24.log 2> gpurun_out/bench_r24.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r24.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r24.log').read().strip().splitlines()[-1])
print('%.4e'%d['value'], round(d['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), 'e2e %.4e'%d['e2e']['value'])
print(json.dumps(d['simplex'])[:1500])
PY
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r24.log').read().strip().splitlines()[-1])
print(json.dumps(d.get('e2e_host_state_every_step')))
PY
