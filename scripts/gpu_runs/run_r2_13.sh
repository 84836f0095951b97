#!/bin/bash
# round 2, GPU call 13: bulk-copy marching kernels after the uninitialised-window fix: blocks per SM x window layout
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 1500 python scripts/gpu_runs/r2_variants.py --only exp_ldgsts6_register_window,exp_tma6_slot_plane,exp_tma6_register_window,exp_tma7_slot_plane,exp_tma7_register_window,exp_ldgsts7_slot_plane,exp_tma5_slot_plane > gpurun_out/r2_13_variants.jsonl 2>&1
python - <<'PY'
import json
for line in open("gpurun_out/r2_13_variants.jsonl"):
    if not line.startswith("VARIANT "): print(line[:300]); continue
    d = json.loads(line[8:])
    if "error" in d: print(d["variant"], "ERROR", d["error"][-300:]); continue
    print("%-32s %.2f ms/step  x %.2f y %.2f z %.2f  whole %.3f  checksum %.6f" % (d["variant"], d["ms_per_step"], *d["stage_ms"], d["whole_step_frac"], d["checksum"]))
PY
