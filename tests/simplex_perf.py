"""Times the simplex stage kernels on one GPU (not a test; run by hand on the GPU box):
    python tests/simplex_perf.py [cubes_per_side] [steps]"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), HERE, os.path.join(os.path.dirname(HERE), "oracle")]

import gcm_b200
from gcm_b200 import capi
from simplex_helpers import Mesh, SimplexBody


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    lib = gcm_b200.library()
    t0 = time.time()
    m = Mesh(lib.h, n, n, n, h=1.0 / n, jitter=0.3, seed=1)
    print("mesh: %d vertices, %d cells, built in %.1f s" % (m.nV, m.nC, time.time() - t0), flush=True)
    ctx = capi.Context(lib)
    for model, name in ((0, "elastic"), (1, "acoustic")):
        body = SimplexBody(lib, ctx, m, model)
        M = body.M
        U, U1, L = capi.host_matrices(lib, name, 3, ("isotropic", 2.0, 3.0, 1.2 if model == 0 else 0.0))
        body.set_material(U, U1, L, np.eye(3))
        _, st, _, cn = body.vertices()
        nodes = np.nonzero(st)[0].astype(np.int32)
        body.border_set(np.array([0], dtype=np.int32), nodes, cn[nodes], np.zeros(len(nodes), dtype=np.int32))
        X = m.local_xyz()
        pde = np.zeros((m.n_local, M))
        pde[:, M - 1] = np.exp(-(np.linalg.norm(X - 0.5, axis=1) / 0.2) ** 2)
        body.upload(pde)
        tau = 0.3 / n / np.abs(L).max()
        vals = np.zeros((1, 3 if model == 0 else 1))
        for warm in range(2):
            for s in range(3):
                body.stage(s, tau, vals)
        ctx.sync()
        ctx.profile_enable(True)
        ctx.timer_start()
        for step in range(steps):
            body.plain_border(vals)
            for s in range(3):
                body.stage(s, tau, vals)
        ms = ctx.timer_stop()
        pms, pn = ctx.profile_get()
        ctx.profile_enable(False)
        out = body.download()
        print("%s: %.3f ms/step, %.3e vertex-updates/s, errors %d, finite %s, max %.3g" %
              (name, ms / steps, m.n_local * steps / (ms * 1e-3), body.errors(), np.isfinite(out).all(), np.abs(out).max()))
        print("   per class ms/step:", {k: round(pms[k] / steps, 3) for k in range(8) if pn[k]}, flush=True)
        body.close()
    ctx.close()


if __name__ == "__main__":
    main()
