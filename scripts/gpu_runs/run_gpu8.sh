python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 tests/multi_gpu_check.py > gpurun_out/multi8.log 2>&1; echo "multi8 rc=$?"; grep -E "multi-gpu|MULTI|Error|error" gpurun_out/multi8.log | head -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench8_n8.log 2> gpurun_out/bench8_n8.err; echo "bench n8 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench8_n4.log 2> gpurun_out/bench8_n4.err; echo "bench n4 rc=$?"
for f in bench8_n4 bench8_n8; do python - <<PY
import json
d=json.loads(open('gpurun_out/$f.log').read().strip().splitlines()[-1])
print('$f', '%.3e'%d['value'], round(d['ms_per_step'],2), {k[-8:]:round(v,2) for k,v in d['roofline']['per_stage_ms'].items()}, 'e2e %.3e'%d['e2e']['value'], d['clocks'], d['gpu_launches'])
PY
done
