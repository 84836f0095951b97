"""TEST INFRASTRUCTURE ONLY — drives oracle/simplex_oracle.c (CPU restatement of the reference's simplex GCM, pinned
against the unmodified reference engine: see the header of simplex_oracle.h) over whole time steps of one body.  Used by bench.py's cpu_baseline
leg; nothing under gcm_b200/ may import this."""
import ctypes
import time

import numpy as np

import oracle_host as oh

ip = ctypes.POINTER(ctypes.c_int)
dp = ctypes.POINTER(ctypes.c_double)


class GcmoTri(ctypes.Structure):
    _fields_ = [("nV", ctypes.c_int), ("nC", ctypes.c_int), ("xyz", dp), ("cell_v", ip), ("cell_n", ip),
                ("cell_grid", ip), ("inc_off", ip), ("inc_cell", ip), ("grid_id", ctypes.c_int),
                ("local_of", ip), ("global_of", ip), ("n_local", ctypes.c_int)]


def _i(a):
    return a.ctypes.data_as(ip)


def _d(a):
    return a.ctypes.data_as(dp)


def body_view(tri, grid_id):
    """(GcmoTri, keep-alive arrays) of body grid_id in the triangulation dict
    {xyz, cell_v, cell_n, cell_grid, inc_off, inc_cell}"""
    used = np.zeros(len(tri["xyz"]), dtype=bool)
    used[tri["cell_v"][tri["cell_grid"] == grid_id].ravel()] = True
    global_of = np.nonzero(used)[0].astype(np.int32)
    local_of = np.full(len(used), -1, dtype=np.int32)
    local_of[global_of] = np.arange(len(global_of), dtype=np.int32)
    t = GcmoTri(len(tri["xyz"]), len(tri["cell_v"]), _d(tri["xyz"]), _i(tri["cell_v"]), _i(tri["cell_n"]), _i(tri["cell_grid"]),
                _i(tri["inc_off"]), _i(tri["inc_cell"]), grid_id, _i(local_of), _i(global_of), len(global_of))
    return t, (global_of, local_of)


def run_single_body(tri, grid_id, model, U, U1, L, basis, nodes, normals, conds, types, values, pde, tau, steps, time0=0.0):
    """`steps` time steps of simplex::Engine::nextTimeStep (engine/simplex/Engine.cpp:97-141) for one body;
    returns (state, seconds spent in the stages)"""
    lib = oh.lib()
    tp = ctypes.POINTER(GcmoTri)
    lib.gcmo_simplex_stage.argtypes = [tp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, dp, dp, dp, dp,
                                       ctypes.c_int, ip, dp, ip, ctypes.c_int, ip, dp, dp, dp]
    lib.gcmo_simplex_plain_border.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, dp, ip, ip, dp, dp]
    t, keep = body_view(tri, grid_id)
    M = pde.shape[1]
    cur = np.ascontiguousarray(pde, dtype=np.float64).copy()
    basis = np.ascontiguousarray(basis, dtype=np.float64)
    now = time0
    spent = 0.0
    for _ in range(steps):
        b = np.ascontiguousarray(values(now + tau), dtype=np.float64)
        t0 = time.perf_counter()
        lib.gcmo_simplex_plain_border(model, M, len(nodes), _i(nodes), _d(normals), _i(conds), _i(types), _d(b), _d(cur))
        for s in range(3):
            nxt = np.zeros_like(cur)
            errs = lib.gcmo_simplex_stage(ctypes.byref(t), model, M, s, tau, _d(U), _d(U1), _d(L), _d(basis), len(nodes), _i(nodes),
                                          _d(normals), _i(conds), len(types), _i(types), _d(b), _d(cur), _d(nxt))
            if errs:
                raise RuntimeError("the oracle hit %d node computations the reference would throw on" % errs)
            cur = nxt
        spent += time.perf_counter() - t0
        now += tau
    del keep
    return cur, spent
