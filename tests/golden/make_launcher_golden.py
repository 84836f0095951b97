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

path = os.path.join(HERE, "launcher_tasks.json")
out = json.load(open(path)) if os.path.exists(path) else {}
# main.cpp:332-467 (cubic2d, cubic3d, acoustic) and ndi.hpp:162-317 (ndi_empty, ndi, titan)
for name in sys.argv[1:] or ("cubic2d", "cubic3d", "acoustic", "ndi_empty", "ndi", "titan"):
    text = open(os.path.join(ROOT, "gcm_b200", "tasks", name + ".task")).read()
    r = oh.run_reference(text, tempfile.mkdtemp())
    bodies = sorted(k for k in r if isinstance(k, int))
    first = bodies[0]
    u = r[first]
    weights = np.arange(1, u.shape[1] + 1, dtype=np.float64)
    out[name] = {"steps": int(r["meta"]["steps"]), "time": r["meta"]["time"], "body": first, "checksum": float((u * weights).sum()),
                 "abs_sum": float(np.abs(u).sum()), "nodes": int(u.shape[0]),
                 "bodies": {str(b): {"checksum": float((r[b] * weights).sum()), "abs_sum": float(np.abs(r[b]).sum())} for b in bodies}}
    if "detector" in r:
        out[name]["detector_last"] = [float(x) for x in r["detector"][-1]]
    print(name, out[name])
json.dump(out, open(path, "w"), indent=1)
