python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu2.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.log 2> gpurun_out/bench2.err; echo "bench rc=$?"
GCMB_MARCH_SEG=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench2_seg0.log 2>&1
GCMB_MARCH_SEG=128 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench2_seg128.log 2>&1
GCMB_ZTILE_ROWS=8 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench2_rows8.log 2>&1
GCMB_ZTILE_ROWS=128 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench2_rows128.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --size 384 > gpurun_out/plain384.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_stage -s 9 -c 3 -o gpurun_out/prof_r1b python bench.py --steps 2 --warmup 3 --no-cpu-baseline --size 384 > gpurun_out/ncu_full2.log 2>&1; echo "ncu full rc=$?"
for f in bench2 bench2_seg0 bench2_seg128 bench2_rows8 bench2_rows128; do python - <<PY
import json
d=json.loads(open('gpurun_out/$f.log').read().strip().splitlines()[-1])
print('$f', '%.3e'%d['value'], round(d['ms_per_step'],2), {k[-8:]:round(v,2) for k,v in d['roofline']['per_stage_ms'].items()}, 'e2e %.3e'%d['e2e']['value'], d['clocks'])
PY
done
