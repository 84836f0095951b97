#!/bin/bash
# round 2, GPU call 30: bulk-copy marching kernels with 6 ring slots per warp at 3 blocks/SM, every configuration of the bench (GCMB_STAGE_IMPL=3 forces them everywhere)
cd "$GRAFT_REPO_ROOT" || exit 1
GCMB_STAGE_IMPL=3 timeout 900 python bench.py --no-cpu-baseline --no-simplex --no-rotated --no-host-roundtrip > gpurun_out/r2_30_bench_impl3.json 2> gpurun_out/r2_30_bench_impl3.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_30_bench_impl3.json").read())
print("headline", d["value"], d["ms_per_step"], d["roofline"]["per_stage_ms"])
for k in ("config4", "config2", "fp32", "fma", "courant1", "courant09_same_size"):
    print(k, d[k].get("value"), d[k].get("ms_per_step"), d[k].get("per_class_ms"), d[k].get("error"))
print(json.dumps(d["small_grids"]))
PY
GCMB_STAGE_IMPL=3 timeout 300 python scripts/gpu_runs/r2_2d.py 1024 4096 8192 2>&1 | grep "^2D" | cut -c1-200
GCMB_STAGE_IMPL=3 timeout 900 python -m pytest tests -m gpu -x -q -k "variants or engine_matches or random or anchor or border" 2>&1 | tail -2
