"""Simplex-path checks shared by the CPU (stepping harness) and GPU test modules.
The oracle is oracle/simplex_oracle.c — parity UNPINNED by the reference (no CGAL here, no pinned values in its
tests); these checks therefore also assert the PROPERTIES the reference's own tests assert."""
import ctypes

import numpy as np

from gcm_b200 import capi
from simplex_helpers import Mesh, SimplexBody, _d, _i, directions, oracle, oracle_locate_all, protocol_queries


def make_mesh(lib, kind):
    if kind == "jitter_void":
        return Mesh(lib.h, 6, 5, 7, origin=(0.0, -0.5, 0.25), h=0.4, jitter=0.35, seed=3, void_box=(0.7, 0.2, 1.0, 1.7, 1.0, 2.2))
    if kind == "regular":
        return Mesh(lib.h, 4, 4, 4, h=1.0, jitter=0.0)
    if kind == "big":
        return Mesh(lib.h, 24, 24, 24, h=0.1, jitter=0.3, seed=5, void_box=(0.9, 0.9, 0.9, 1.5, 1.5, 1.5))
    raise ValueError(kind)


def check_vertex_info(lib, kind="jitter_void"):
    L = oracle()
    m = make_mesh(lib, kind)
    ctx = capi.Context(lib)
    body = SimplexBody(lib, ctx, m, 0)
    g, st, bn, cn = body.vertices()
    assert np.array_equal(g, m.global_of)
    t = m.oracle_view()
    for v in range(m.n_local):
        assert st[v] == L.gcmo_simplex_border_state(ctypes.byref(t), v)
        for which, arr in ((0, bn), (1, cn)):
            ref = np.zeros(3)
            L.gcmo_simplex_normal(ctypes.byref(t), v, which, _d(ref))
            assert np.array_equal(ref, arr[v]), (v, which, ref, arr[v])
    # every border normal points out of the body (towards the void or the outside of the box)
    assert (st == 0).sum() > 0 and (st == 1).sum() > 0
    body.close(); ctx.close()


def check_locate_protocol(lib, kind="jitter_void", n_dirs=16, lengths=9, vertices=None):
    """GPU/harness cell location == oracle, integer for integer; and the reference tests' containment property"""
    L = oracle()
    m = make_mesh(lib, kind)
    ctx = capi.Context(lib)
    body = SimplexBody(lib, ctx, m, 0)
    v, sh = protocol_queries(m, n_dirs, lengths, scale=0.4, vertices=vertices)
    got = body.locate(v, sh)
    assert body.errors() == 0
    ref, errs = oracle_locate_all(L, m, v, sh)
    assert errs == 0
    assert np.array_equal(got, ref), "cell location differs in %d of %d queries" % ((got != ref).any(axis=1).sum(), len(v))
    # TestLineWalkSearch3D.cpp:120-154: a returned cell contains the query (eps = 1e-9)
    X = m.local_xyz()
    full = got[:, 0] == 4
    q = X[v[full]] + sh[full]
    P = X[got[full, 1:5]]                      # [n, 4, 3]
    T = np.stack([P[:, 0] - P[:, 3], P[:, 1] - P[:, 3], P[:, 2] - P[:, 3]], axis=2)
    lam = np.linalg.solve(T, (q - P[:, 3])[:, :, None])[:, :, 0]
    lam = np.concatenate([lam, 1 - lam.sum(axis=1, keepdims=True)], axis=1)
    assert lam.min() > -1e-8
    # inner vertices always get an answer: a cell, or the border facet the ray leaves through
    _, st, _, _ = body.vertices()
    assert (got[st[v] == 0, 0] >= 1).all()
    hist = np.bincount(got[:, 0], minlength=5)
    assert hist[4] > 0 and hist[0] + hist[3] > 0
    body.close(); ctx.close()
    return hist


def check_gradient(lib, kind="jitter_void"):
    L = oracle()
    m = make_mesh(lib, kind)
    ctx = capi.Context(lib)
    rng = np.random.default_rng(1)
    for model, M in ((0, 9), (1, 4)):
        body = SimplexBody(lib, ctx, m, model)
        vals = rng.normal(size=(m.n_local, M))
        got = body.gradient(vals)
        ref = np.zeros((m.n_local, 3, M))
        t = m.oracle_view()
        assert L.gcmo_simplex_gradient(ctypes.byref(t), M, _d(vals), _d(ref)) == 0
        assert np.array_equal(got, ref)
        # a linear field has an exact gradient (the weighted least squares reproduces it)
        X = m.local_xyz()
        coef = rng.normal(size=(3, M))
        lin = X @ coef + rng.normal(size=M)
        g = body.gradient(lin)
        assert np.abs(g - coef[None]).max() < 1e-9
        body.close()
    ctx.close()


def _border_nodes(L, m, cond_areas):
    """Engine::addBorderNode (engine/simplex/Engine.cpp:288-309): the last condition whose area holds the vertex"""
    t = m.oracle_view()
    X = m.local_xyz()
    nodes, normals, conds = [], [], []
    for v in range(m.n_local):
        if L.gcmo_simplex_border_state(ctypes.byref(t), v) == 0:
            continue
        chosen = -1
        for c, contains in enumerate(cond_areas):
            if contains(X[v]):
                chosen = c
        if chosen < 0:
            continue
        n = np.zeros(3)
        assert L.gcmo_simplex_normal(ctypes.byref(t), v, 1, _d(n)) == 1
        nodes.append(v); normals.append(n); conds.append(chosen)
    return np.array(nodes, dtype=np.int32), np.array(normals).reshape(-1, 3), np.array(conds, dtype=np.int32)


def check_stage(lib, model, kind="jitter_void", steps=2, zero=False):
    """whole time steps (plain correction + 3 stages) == oracle, bit for bit; no node computation the reference
    would have thrown on"""
    import oracle_host as oh
    L = oracle()
    m = make_mesh(lib, kind)
    M = 9 if model == 0 else 4
    outer = 3 if model == 0 else 1
    mat = {"kind": "isotropic", "rho": 2.0, "lambda": 3.0, "mu": 1.2 if model == 0 else 0.0}
    U, U1, Lm = oh.matrices_for("elastic" if model == 0 else "acoustic", 3, mat)
    Up, U1p, Lp = capi.host_matrices(lib, "elastic" if model == 0 else "acoustic", 3, ("isotropic", 2.0, 3.0, mat["mu"]))
    assert np.array_equal(U, Up) and np.array_equal(U1, U1p) and np.array_equal(Lm, Lp)
    basis = np.eye(3)
    X = m.local_xyz()
    zmax = X[:, 2].max()
    # two conditions: free surface everywhere, a moving piston on the top face (later one wins there)
    areas = [lambda p: True, lambda p: p[2] > zmax - 1e-9]
    types = np.array([0, 1], dtype=np.int32)
    nodes, normals, conds = _border_nodes(L, m, areas)
    assert len(nodes) > 0 and (conds == 1).any() and (conds == 0).any()

    def values(time):
        if model == 0:
            return np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.0 if zero else 0.3 * np.sin(7 * time)]])
        return np.array([[0.0], [0.0 if zero else 0.3 * np.sin(7 * time)]])

    rng = np.random.default_rng(4)
    pde = np.zeros((m.n_local, M))
    if not zero:
        r = np.linalg.norm(X - X.mean(axis=0), axis=1)
        amp = np.exp(-(r / 0.6) ** 2)
        if model == 0:
            for i in (3, 6, 8):
                pde[:, i] = -amp
            pde[:, 0] = 0.1 * amp
        else:
            pde[:, 3] = amp
    tau = 0.35 * 0.4 / np.abs(Lm).max()
    ctx = capi.Context(lib)
    body = SimplexBody(lib, ctx, m, model)
    body.set_material(U, U1, Lm, basis)
    body.border_set(types, nodes, normals, conds)
    body.upload(pde)
    t = m.oracle_view()
    cur = pde.copy()
    time = 0.0
    for step in range(steps):
        b_next = values(time + tau)
        body.plain_border(b_next)
        L.gcmo_simplex_plain_border(model, M, len(nodes), _i(nodes), _d(normals), _i(conds), _i(types), _d(b_next), _d(cur))
        for s in range(3):
            body.stage(s, tau, b_next)
            nxt = np.zeros_like(cur)
            errs = L.gcmo_simplex_stage(ctypes.byref(t), model, M, s, tau, _d(U), _d(U1), _d(Lm), _d(basis),
                                        len(nodes), _i(nodes), _d(normals), _i(conds), 2, _i(types), _d(b_next), _d(cur), _d(nxt))
            assert errs == 0, (step, s, errs)
            cur = nxt
        time += tau
        got = body.download()
        assert np.array_equal(got, cur), (step, np.abs(got - cur).max())
    assert body.errors() == 0
    if zero:
        assert not cur.any()  # TestSimplexGcm.cpp:29-67: zero stays exactly zero, with and without fixed force
    else:
        assert np.isfinite(cur).all() and np.abs(cur).max() < 10 and np.abs(cur - pde).max() > 1e-3
    body.close(); ctx.close()
    return cur
