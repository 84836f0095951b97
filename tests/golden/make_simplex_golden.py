"""Golden vectors of the simplex path from the UNMODIFIED reference engine.

oracle/_ref/gcm_ref_simplex is the reference's simplex::Engine<3, CgalTriangulation> (engine/simplex/Engine.cpp,
grid/simplex/SimplexGrid.cpp, cgal/CgalTriangulation.cpp, LineWalker.hpp, the GCMs, correctors, interpolators, linal --
compiled where they lie) with CGAL replaced by the flat container of oracle/shim/CGAL/flat_triangulation_3.h and GSL's
LU by its restatement (oracle/shim/libgcm/util/math/GslUtils.hpp).  For each scenario this script
  1. builds a triangulation with the product's box mesher (host library on the stepping harness: mesh generation is
     plain host C++).  The first five fixtures were made before the host layer had its own restatement of the
     reference's clean-up of body ids, so their cell_grid_before_cleanup differs from cell_grid (the reference
     changed 2-4 cells); fixtures made later come out of the mesher already cleaned,
  2. runs the reference engine on it; the reference first cleans the triangulation's body ids (hanged cells,
     disconnected cell sets: CgalTriangulation.cpp:8-112) and the cleaned ids are kept,
  3. stores the triangulation with the cleaned ids, the task text and the reference's final PDE values per body.
Run in the build container:  python tests/golden/make_simplex_golden.py"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import helpers  # noqa: E402
import simplex_cases  # noqa: E402
from gcm_b200 import capi  # noqa: E402
from simplex_helpers import run_reference_simplex  # noqa: E402

SCENARIOS = {
    # name: (model, bodies, gcm_type, basis, steps)
    "elastic_cavity": (0, 1, 0, "identity", 4),
    "elastic_contact": (0, 2, 0, "identity", 4),
    "acoustic_contact_rotated": (1, 2, 0, "rotated", 4),
    "elastic_contact_pde_vectors": (0, 2, 1, "rotated", 4),
    "acoustic_pde_vectors": (1, 1, 1, "identity", 4),
    "elastic_contact_summ": (0, 2, 0, "rotated", 4, "splitting summ"),
    "elastic_contact_local_basis": (0, 2, 0, "rotated", 4, "border_calc_mode local"),
    "acoustic_cavity_local_basis_pde_vectors": (1, 1, 1, "identity", 4, "border_calc_mode local"),
    "elastic_maxwell": (0, 1, 0, "identity", 4, "MAXWELL"),
    "elastic_three_bodies": (0, 3, 0, "rotated", 4),
    # the reference launcher's cubeAcs / cubeEls tasks at reduced size (simplex_cases.cube_scenario)
    "cube_acs": (1, 1, 0, "text", 7),
    "cube_els": (0, 1, 0, "text", 7),
}


def main():
    lib = helpers.emul_library()
    only = sys.argv[1:]
    for name, cfg in SCENARIOS.items():
        if only and name not in only:
            continue
        model, bodies, gcm, basis, steps = cfg[:5]
        if basis == "text":
            text = simplex_cases.cube_scenario(model == 1, steps=steps)
        else:
            text = simplex_cases.golden_scenario(model, bodies, gcm, basis, steps) + "".join("\n" + extra for extra in cfg[5:] if extra != "MAXWELL")
        if "MAXWELL" in cfg[5:]:   # Maxwell viscosity ODE: the body line gets "ode maxwell", the material a relaxation time
            text = text.replace("body 0 elastic isotropic", "body 0 elastic isotropic ode maxwell")
            text = "\n".join(ln + " tau0 0.2" if ln.startswith("material body 0") else ln for ln in text.split("\n"))
        eng = capi.SimplexHostEngine(lib, text)
        tri = eng.triangulation()
        eng.close()
        M = 9 if model == 0 else 4
        ref, meta = run_reference_simplex(text, tri, tempfile.mkdtemp(), range(bodies), M)
        out = dict(task=np.array(text), xyz=tri["xyz"], cell_v=tri["cell_v"], cell_n=tri["cell_n"], cell_grid=meta["cell_grid"],
                   cell_grid_before_cleanup=tri["cell_grid"], time=meta["time"], tau=meta["tau"], steps=int(meta["steps"]), model=model,
                   bodies=bodies, gcm_type=gcm)
        for b in range(bodies):
            out["coords%d" % b], out["pde%d" % b] = ref[b]
            out["average_height%d" % b] = meta[("body", b)]["average_height"]
        np.savez_compressed(os.path.join(HERE, "simplex_%s.npz" % name), **out)
        print(name, "cells", len(tri["cell_v"]), "changed by the reference clean-up", int((meta["cell_grid"] != tri["cell_grid"]).sum()),
              "max |u|", max(np.abs(ref[b][1]).max() for b in range(bodies)))


def locate_golden():
    """SimplexGrid::findCellCrossedByTheRay of the reference over the protocol of TestLineWalkSearch3D.cpp:120-154
    (16 x 16 directions x 9 lengths) from every third vertex of both bodies of the elastic_contact fixture"""
    import subprocess
    from simplex_helpers import Mesh, protocol_queries, write_flat_dump
    g = simplex_cases.load_golden("elastic_contact")
    tri = dict(xyz=g["xyz"], cell_v=g["cell_v"], cell_n=g["cell_n"], cell_grid=g["cell_grid"])
    wd = tempfile.mkdtemp()
    os.makedirs(os.path.join(wd, "snapshots"))   # the reference's grids write cell-height histograms there
    write_flat_dump(tri, os.path.join(wd, "mesh.flat"))
    open(os.path.join(wd, "task.txt"), "w").write(str(g["task"]) + "\nsimplex_flat %s\n" % os.path.join(wd, "mesh.flat"))
    out = {}
    for body in (0, 1):
        m = Mesh.from_arrays(dict(tri, inc_off=None, inc_cell=None), body)
        vs = np.arange(0, m.n_local, 3)
        v, sh = protocol_queries(m, 16, 9, scale=0.4, vertices=vs)
        with open(os.path.join(wd, "queries.txt"), "w") as f:
            for a, s in zip(v, sh):
                f.write("%d %d %.17e %.17e %.17e\n" % (body, a, s[0], s[1], s[2]))
        exe = os.path.join(ROOT, "oracle", "_ref", "gcm_ref_simplex")
        subprocess.run([exe, os.path.join(wd, "task.txt"), os.path.join(wd, "out"), "--locate", os.path.join(wd, "queries.txt")], check=True, cwd=wd)
        rows = []
        for line in open(os.path.join(wd, "out.located")):
            rows.append([-9] * 5 if line.startswith("throw") else [int(x) for x in line.split()])
        out["located%d" % body] = np.array(rows, dtype=np.int16)
        out["vertices%d" % body] = vs.astype(np.int32)
        print("body", body, "queries", len(rows), "histogram of n", np.bincount(np.maximum(out["located%d" % body][:, 0], 0), minlength=5),
              "throws", int((out["located%d" % body][:, 0] == -9).sum()))
    np.savez_compressed(os.path.join(HERE, "simplex_locate_protocol.npz"), **out)


if __name__ == "__main__":
    main()
    if not sys.argv[1:]:
        locate_golden()
