"""Simplex path: cell location with the direction masks (simplex_fns.h "direction buckets") against the plain loop over the
incident cells (GCMB_SX_DIR_MASKS=0), and the gradient on its stored geometry with a thread per component against a thread
per vertex recomputing it (GCMB_SX_GRAD_GEOMETRY=0); both regimes; one fresh process per setting."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import gcm_b200
    import bench
    for basis in ("random", "identity"):
        out = bench.simplex_section(gcm_b200.library(), 0, 20, 3, False, basis=basis)
        print("SX " + json.dumps({"dir_masks": os.environ.get("GCMB_SX_DIR_MASKS", "1"), "grad_geometry": os.environ.get("GCMB_SX_GRAD_GEOMETRY", "1"), "basis": basis, "ms_per_step": out["ms_per_step"], "value": out["value"],
                                  "per_class_ms": out["per_class_ms"]}), flush=True)
    sys.exit(0)
for env in ({}, {"GCMB_SX_GRAD_GEOMETRY": "0"}, {"GCMB_SX_DIR_MASKS": "0"}, {"GCMB_SX_DIR_MASKS": "0", "GCMB_SX_GRAD_GEOMETRY": "0"}):
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
    lines = [x for x in r.stdout.splitlines() if x.startswith("SX ")]
    print("\n".join(lines) if lines else "SX " + json.dumps({"env": env, "error": (r.stdout + r.stderr)[-400:]}), flush=True)
