nvidia-smi -L
nvidia-smi topo -m | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 tests/multi_gpu_check.py > gpurun_out/multi2.log 2>&1; echo "multi rc=$?"; grep -E "multi-gpu|MULTI|Error|error" gpurun_out/multi2.log | head -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench6_n2.log 2> gpurun_out/bench6_n2.err; echo "bench n2 rc=$?"
python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench6_n1.log 2> gpurun_out/bench6_n1.err; echo "bench n1 rc=$?"
for f in bench6_n1 bench6_n2; do python - <<PY
import json
d=json.loads(open('gpurun_out/$f.log').read().strip().splitlines()[-1])
print('$f', '%.3e'%d['value'], round(d['ms_per_step'],2), {k[-8:]:round(v,2) for k,v in d['roofline']['per_stage_ms'].items()}, 'e2e %.3e'%d['e2e']['value'], d['clocks'], d['gpu_launches'])
PY
done
tail -5 gpurun_out/bench6_n2.err
