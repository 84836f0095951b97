"""Occupancy knobs of the simplex kernels in the uncached regime (a new random basis every step), one process per setting.
Needs a build with GCMB_NVCC_EXTRA=-DGCMB_SX_EXPERIMENTS."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import gcm_b200
    import bench
    out = bench.simplex_section(gcm_b200.library(), 0, 20, 3, False, basis="random")
    print("SX " + json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("GCMB_SX")}, "ms_per_step": out["ms_per_step"],
                              "value": out["value"], "per_class_ms": out["per_class_ms"]}))
    sys.exit(0)
for env in ({}, {"GCMB_SX_INNER_MINB": "3"}, {"GCMB_SX_INNER_MINB": "2"}, {"GCMB_SX_INNER_MINB": "5"}, {"GCMB_SX_GRAD_MINB": "4"}, {"GCMB_SX_GRAD_MINB": "5"},
            {"GCMB_SX_INNER_MINB": "3", "GCMB_SX_GRAD_MINB": "4"}):
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    lines = [x for x in r.stdout.splitlines() if x.startswith("SX ")]
    print(lines[0] if lines else "SX " + json.dumps({"env": env, "error": (r.stdout + r.stderr)[-400:]}), flush=True)
