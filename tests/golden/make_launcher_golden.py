"""Runs the unmodified reference (oracle/_ref/gcm_ref) on the task files shipped in gcm_b200/tasks/ (the reference
launcher's cubic demo tasks; the simplex cube tasks go through oracle/_ref/gcm_ref_simplex) and records step counts, end times and state checksums in launcher_tasks.json.
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

SIMPLEX_TASKS = ("cubeAcs", "cubeEls")   # main.cpp:547-640; pinned by the reference's simplex engine on the product mesher's triangulation


def simplex_golden(name):
    """oracle/_ref/gcm_ref_simplex (the unmodified reference simplex engine, CGAL stand-in) on the triangulation the
    product's box mesher builds for the shipped task file"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    from gcm_b200 import capi
    from simplex_helpers import run_reference_simplex
    text = open(os.path.join(ROOT, "gcm_b200", "tasks", name + ".task")).read()
    eng = capi.SimplexHostEngine(helpers.emul_library(), text)
    tri = eng.triangulation()
    eng.close()
    M = 4 if "acoustic" in text.split("body 0")[1].split("\n")[0] else 9
    # (the reference driver of the tests writes no VTK files: the snapshot line is dropped for it)
    text_ref = "\n".join(ln for ln in text.splitlines() if not ln.startswith("vtk")) + "\n"
    ref, meta = run_reference_simplex(text_ref, tri, tempfile.mkdtemp(), range(1), M)
    u = ref[0][1]
    weights = np.arange(1, M + 1, dtype=np.float64)
    return {"steps": int(meta["steps"]), "time": float(meta["time"]), "body": 0, "checksum": float((u * weights).sum()),
            "abs_sum": float(np.abs(u).sum()), "nodes": int(u.shape[0]),
            "bodies": {"0": {"checksum": float((u * weights).sum()), "abs_sum": float(np.abs(u).sum())}}}


path = os.path.join(HERE, "launcher_tasks.json")
out = json.load(open(path)) if os.path.exists(path) else {}
# main.cpp:332-467 (cubic2d, cubic3d, acoustic) and ndi.hpp:162-317 (ndi_empty, ndi, titan)
for name in sys.argv[1:] or ("cubic2d", "cubic3d", "acoustic", "ndi_empty", "ndi", "titan") + SIMPLEX_TASKS:
    if name in SIMPLEX_TASKS:
        out[name] = simplex_golden(name)
        print(name, out[name])
        continue
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
