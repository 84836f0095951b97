set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke2.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke2.log
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_r15.log 2> gpurun_out/bench_r15.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r15.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r15.log').read().strip().splitlines()[-1])
print('%.3e'%d['value'], round(d['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), 'e2e %.3e'%d['e2e']['value'])
print(json.dumps(d.get('simplex'), indent=1))
PY
