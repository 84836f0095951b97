"""Runs the unmodified reference (oracle/_ref/gcm_ref) on the task files shipped in gcm_b200/tasks/ (the reference
launcher's cubic demo tasks) and records step counts, end times and state checksums in launcher_tasks.json.
Run in the build container (needs /root/reference to build gcm_ref):  python tests/golden/make_launcher_golden.py"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import oracle_host as oh  # noqa: E402

out = {}
for name in ("cubic2d", "cubic3d", "acoustic"):
    text = open(os.path.join(ROOT, "gcm_b200", "tasks", name + ".task")).read()
    r = oh.run_reference(text, tempfile.mkdtemp())
    u = r[0]
    weights = np.arange(1, u.shape[1] + 1, dtype=np.float64)
    out[name] = {"steps": int(r["meta"]["steps"]), "time": r["meta"]["time"], "checksum": float((u * weights).sum()),
                 "abs_sum": float(np.abs(u).sum()), "nodes": int(u.shape[0])}
    print(name, out[name])
json.dump(out, open(os.path.join(HERE, "launcher_tasks.json"), "w"), indent=1)
