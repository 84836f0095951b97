"""Simplex-path checks shared by the CPU (stepping harness) and GPU test modules.
The oracle is oracle/simplex_oracle.c, pinned by the fixtures of the unmodified reference engine
(tests/golden/simplex_*.npz, see check_engine_against_reference / check_oracle_against_reference below); the checks
also assert the PROPERTIES the reference's own tests assert."""
import ctypes

import numpy as np

from gcm_b200 import capi
from simplex_helpers import Mesh, SimplexBody, SimplexContact, _d, _i, directions, oracle, oracle_locate_all, protocol_queries


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


def check_direction_masks(lib):
    """the direction masks of the incident-cell search (simplex_fns.h "direction buckets") never change the answer: rays along
    the axes, face and space diagonals of the regular mesh (they lie ON cell faces and edges: several cells pass the exact
    test and the first in the reference's order must win), rays on the bucket boundaries, random rays, lengths from far
    below the masks' validity limit to several cells -- with the masks against the plain loop, and against the oracle"""
    import os
    L = oracle()
    rng = np.random.default_rng(5)
    axes = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1], [0, 1, 1], [1, -1, 0], [1, 0, -1], [0, 1, -1],
                     [1, 1, 1], [1, 1, -1], [1, -1, 1], [-1, 1, 1], [1, 0.5, 0], [1, 0.5, 0.5], [0.5, 1, 0], [0, 0.5, 1], [1, 0.5, -0.5]], dtype=float)
    dirs = np.concatenate([axes, -axes, rng.normal(size=(40, 3))])
    lens = np.array([1e-12, 1e-7, 1e-3, 0.05, 0.3, 0.9, 2.5])
    for kind in ("regular", "jitter_void"):
        m = make_mesh(lib, kind)
        vs = rng.choice(m.n_local, size=min(m.n_local, 60), replace=False).astype(np.int32)
        v = np.repeat(vs, len(dirs) * len(lens)).astype(np.int32)
        sh = (dirs[None, :, None, :] * lens[None, None, :, None] * 0.4).reshape(1, -1, 3)
        sh = np.broadcast_to(sh, (len(vs), sh.shape[1], 3)).reshape(-1, 3).copy()
        got = {}
        for masks in ("1", "0"):
            os.environ["GCMB_SX_DIR_MASKS"] = masks
            try:
                ctx = capi.Context(lib)
                body = SimplexBody(lib, ctx, m, 0)
                got[masks] = (body.locate(v, sh), body.errors())
                body.close(); ctx.close()
            finally:
                os.environ.pop("GCMB_SX_DIR_MASKS", None)
        assert got["1"][1] == got["0"][1]
        assert np.array_equal(got["1"][0], got["0"][0]), "the masks changed %d of %d answers" % ((got["1"][0] != got["0"][0]).any(axis=1).sum(), len(v))
        ref, _ = oracle_locate_all(L, m, v, sh)
        assert np.array_equal(got["1"][0], ref)


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


def check_two_bodies(lib, model, steps=2, kind="layers", gcm_type=0):
    """two bodies glued (elastic, ADHESION) or sliding (acoustic, SLIDE) along a jittered interface, different
    materials, free/forced outer borders: whole time steps in the order of simplex::Engine::nextTimeStep
    (engine/simplex/Engine.cpp:97-141) == the oracle driven in the same order, bit for bit"""
    import oracle_host as oh
    L = oracle()
    base = Mesh(lib.h, 5, 4, 6, origin=(0.0, 0.0, 0.0), h=0.5, jitter=0.3, seed=11,
                void_box=(0.9, 0.6, 2.2, 1.6, 1.4, 2.8) if kind == "layers_void" else None)
    zcut = 1.5
    base.retag(lambda c: 0 if c[2] < zcut else 1)
    meshes = [base.view(0), base.view(1)]
    M = 9 if model == 0 else 4
    outer = 3 if model == 0 else 1
    name = "elastic" if model == 0 else "acoustic"
    mats = [("isotropic", 2.0, 3.0, 1.2 if model == 0 else 0.0), ("isotropic", 1.0, 2.0, 0.7 if model == 0 else 0.0)]
    mat3 = [capi.host_matrices(lib, name, 3, m) for m in mats]
    basis = np.eye(3)
    ctx = capi.Context(lib)
    bodies = [SimplexBody(lib, ctx, m, model) for m in meshes]
    for b, (U, U1, Lm) in zip(bodies, mat3):
        b.set_material(U, U1, Lm, basis)
        b.set_gcm_type(gcm_type)

    # Engine::addBorderOrContact (Engine.cpp:250-309)
    inc = base.incident_grids()
    views = [m.oracle_view() for m in meshes]
    zmax = base.xyz[:, 2].max()
    areas = [lambda p: True, lambda p: p[2] > zmax - 1e-9]
    types = np.array([0, 1], dtype=np.int32)
    border = [([], [], []) for _ in meshes]
    pair_a, pair_b, pair_n = [], [], []
    cn_lib = bodies[0].contact_normals(1)
    for g in range(base.nV):
        grids = set(inc[g])
        if len(grids) == 1:
            continue
        if -1 in grids:
            targets = sorted(grids - {-1})
        elif len(grids) == 2:
            la, lb = meshes[0].local_of[g], meshes[1].local_of[g]
            n = np.zeros(3)
            L.gcmo_simplex_contact_normal(ctypes.byref(views[0]), int(la), 1, _d(n))
            assert np.array_equal(n, cn_lib[la])          # device normals == oracle normals
            if n.any():
                pair_a.append(la); pair_b.append(lb); pair_n.append(n)
            continue
        else:
            targets = sorted(grids)
        for gid in targets:
            m = meshes[gid]
            lv = int(m.local_of[g])
            bn = np.zeros(3)
            multicontact = L.gcmo_simplex_normal(ctypes.byref(views[gid]), lv, 0, _d(bn)) == 0
            chosen = -1
            for c, contains in enumerate(areas):
                if contains(base.xyz[g]) and not multicontact:   # useForMulticontactNodes = false
                    chosen = c
            if chosen < 0:
                continue
            n = np.zeros(3)
            assert L.gcmo_simplex_normal(ctypes.byref(views[gid]), lv, 1, _d(n)) == 1
            border[gid][0].append(lv); border[gid][1].append(n); border[gid][2].append(chosen)
    pair_a = np.array(pair_a, dtype=np.int32); pair_b = np.array(pair_b, dtype=np.int32)
    pair_n = np.array(pair_n).reshape(-1, 3)
    assert len(pair_a) > 10
    border = [(np.array(n_, dtype=np.int32), np.array(r_).reshape(-1, 3), np.array(c_, dtype=np.int32)) for n_, r_, c_ in border]
    for b, (nodes, normals, conds) in zip(bodies, border):
        b.border_set(types, nodes, normals, conds)
    contact = SimplexContact(lib, bodies[0], bodies[1], pair_a, pair_b, pair_n)

    def values(time):
        amp = 0.3 * np.sin(9 * time)
        return np.array([[0.0] * outer, [0.0] * (outer - 1) + [amp]])

    state = []
    for m in meshes:
        X = m.local_xyz()
        r = np.linalg.norm(X - np.array([1.2, 1.0, 1.1]), axis=1)
        amp = np.exp(-(r / 0.7) ** 2)
        pde = np.zeros((m.n_local, M))
        if model == 0:
            for i in (3, 6, 8):
                pde[:, i] = -amp
            pde[:, 1] = 0.2 * amp
        else:
            pde[:, 3] = amp
        state.append(pde)
    for b, pde in zip(bodies, state):
        b.upload(pde)
    lam = max(np.abs(m3[2]).max() for m3 in mat3)
    tau = 0.3 * 0.5 / lam
    time = 0.0
    # the constructor's plain correction at t = 0 (Engine.cpp:45), then the steps
    def plain(at):
        b_at = values(at)
        contact.plain()
        L.gcmo_simplex_plain_contact(model, M, len(pair_a), _i(pair_a), _i(pair_b), _d(pair_n), _d(state[0]), _d(state[1]))
        for b, (nodes, normals, conds), pde in zip(bodies, border, state):
            b.plain_border(b_at)
            L.gcmo_simplex_plain_border(model, M, len(nodes), _i(nodes), _d(normals), _i(conds), _i(types), _d(b_at), _d(pde))
    plain(0.0)
    for step in range(steps):
        plain(time + tau)
        b_next = values(time + tau)
        for s in range(3):
            for b in bodies:
                b.before_stage(s, tau)
            for b in bodies:
                b.border_contact_stage()
            contact.correct()
            for b in bodies:
                b.border_correct(b_next)
            for b in bodies:
                b.inner_stage()
            for b in bodies:
                b.after_stage()
            nxt = [np.zeros_like(p) for p in state]
            hs = [L.gcmo_sx_begin(ctypes.byref(v), model, M, s, tau, _d(m3[0]), _d(m3[1]), _d(m3[2]), _d(basis), _d(cur), _d(nx), gcm_type)
                  for v, m3, cur, nx in zip(views, mat3, state, nxt)]
            for h in hs:
                L.gcmo_sx_nodes(h, 0)
            L.gcmo_sx_contact_correct(hs[0], hs[1], len(pair_a), _i(pair_a), _i(pair_b), _d(pair_n))
            for h, (nodes, normals, conds) in zip(hs, border):
                L.gcmo_sx_border_correct(h, len(nodes), _i(nodes), _d(normals), _i(conds), 2, _i(types), _d(b_next))
            for h in hs:
                L.gcmo_sx_nodes(h, 1)
            for h in hs:
                assert L.gcmo_sx_end(h) == 0, (step, s)
            state = nxt
        time += tau
        for b, pde in zip(bodies, state):
            got = b.download()
            assert np.array_equal(got, pde), (step, np.abs(got - pde).max())
    for b in bodies:
        assert b.errors() == 0
    # the contact condition holds at the interface: equal velocity (elastic) / equal pressure (acoustic)
    a, b = state[0][pair_a], state[1][pair_b]
    if model == 0:
        assert np.abs(a[:, :3] - b[:, :3]).max() < 1e-9
    else:
        assert np.abs(a[:, 3] - b[:, 3]).max() < 1e-9
    assert all(np.isfinite(p).all() and np.abs(p).max() < 10 for p in state)
    contact.close()
    for b in bodies:
        b.close()
    ctx.close()
    return state


# ---- the host engine (gcm_b200/host/simplex_engine.cpp) against the oracle driven in the reference's order ------
def engine_scenario(model, bodies=2, basis="identity", cavity=True, steps=3, gcm_type=0):
    """task text + the same border values as Python callables"""
    import math
    name = "elastic" if model == 0 else "acoustic"
    outer = 3 if model == 0 else 1
    lines = ["grid simplex", "dimensionality 3", "courant 0.7", "steps %d" % steps,
             "simplex_box 5 4 6 0 0 0 0.5 jitter 0.3 seed 7"]
    mats = [(2.0, 3.0, 1.2 if model == 0 else 0.0), (1.0, 2.0, 0.7 if model == 0 else 0.0)]
    for b in range(bodies):
        lines.append("body %d %s isotropic" % (b, name))
        lines.append("material body %d isotropic %r %r %r" % ((b,) + mats[b]))
    if bodies == 2:
        lines.append("region 0 box -10 -10 -10 10 10 1.5")
        lines.append("region 1 box -10 -10 1.5 10 10 10")
        lines.append("contact %s" % ("adhesion" if model == 0 else "slide"))
    if cavity:
        lines.append("cavity box 0.9 0.6 2.2 1.6 1.4 2.8")
    if basis == "identity":
        lines.append("basis 1 0 0 0 1 0 0 0 1")
    elif basis == "rotated":
        c, s_ = math.cos(0.4), math.sin(0.4)
        lines.append("basis %r %r 0 %r %r 0 0 0 1" % (c, -s_, s_, c))
    else:
        lines.append("basis random 5")
    zero = " ".join(["const 0"] * outer)
    lines.append("border_condition infinite fixed_force " + zero)
    lines.append("border_condition box -10 -10 2.999 10 10 10 fixed_velocity no_multicontact "
                 + " ".join(["const 0"] * (outer - 1) + ["sin 0.3 9"]))
    lines.append("initial quantity PRESSURE 1 sphere 0.7 1.2 1.0 1.1")
    lines.append("gcm_type " + ("pde_vectors" if gcm_type == 1 else "riemann_invariants"))
    values = lambda t: np.array([[0.0] * outer, [0.0] * (outer - 1) + [0.3 * math.sin(9 * t)]])
    return "\n".join(lines) + "\n", values


def check_engine(lib, model, bodies=2, basis="identity", cavity=True, steps=3, gcm_type=0, text=None, golden=None):
    """the host engine, step by step, against oracle/simplex_oracle.c driven in the reference's order; with `golden`
    (a fixture of the unmodified reference engine) the restatement's final state must equal the reference's too"""
    L = oracle()
    scenario_text, values = engine_scenario(model, bodies, basis, cavity, steps, gcm_type)
    text = scenario_text if text is None else text
    eng = capi.SimplexHostEngine(lib, text)
    M = 9 if model == 0 else 4
    tri = eng.triangulation()
    ids = list(range(bodies))
    meshes = [Mesh.from_arrays(tri, i) for i in ids]
    views = [m.oracle_view() for m in meshes]
    infos = [eng.simplex_body_info(i) for i in ids]
    types = np.array([0, 1], dtype=np.int32)
    border = []
    for i in ids:
        nodes, normals, conds = [], [], []
        for c in range(infos[i]["n_conditions"]):
            nd, nr = eng.border_nodes(i, c)
            nodes.append(nd); normals.append(nr); conds.append(np.full(len(nd), c, dtype=np.int32))
        border.append((np.concatenate(nodes).astype(np.int32), np.concatenate(normals).reshape(-1, 3), np.concatenate(conds).astype(np.int32)))
        assert len(border[-1][0]) > 0
    pairs = [(0, 1)] if bodies == 2 else []
    contacts = {p: eng.contact_nodes(*p) for p in pairs}
    for p, c in contacts.items():
        assert len(c[0]) > 10
    # the time step: Courant * average height / maximal eigenvalue, minimum over bodies (Engine.hpp:77-92)
    _, time, tau = eng.info()
    assert time == 0.0
    expect = None
    for i in ids:
        h = np.zeros(2)
        L.gcmo_simplex_heights(ctypes.byref(views[i]), _d(h))
        assert h[0] == infos[i]["average_height"] and h[1] == infos[i]["minimal_height"]
        U, U1, Lm = eng.simplex_matrices(i)
        assert np.abs(Lm).max() == infos[i]["maximal_eigenvalue"]
        t_i = 0.7 * h[0] / np.abs(Lm).max()
        expect = t_i if expect is None else min(expect, t_i)
    assert tau == expect
    state = [eng.simplex_pde(i) for i in ids]
    assert all(np.abs(s).max() > 0.5 for s in state[:1])
    for step in range(steps):
        eng.advance(1)
        mats = [eng.simplex_matrices(i) for i in ids]            # the basis may change every step
        basis_m = np.ascontiguousarray(eng.simplex_body_info(0)["basis"])
        if basis == "identity":
            assert np.array_equal(basis_m, np.eye(3))
        else:
            assert np.abs(basis_m.T @ basis_m - np.eye(3)).max() < 1e-12
        b_next = values(time + tau)
        for p, (fa, fb, fn) in contacts.items():
            L.gcmo_simplex_plain_contact(model, M, len(fa), _i(fa), _i(fb), _d(fn), _d(state[p[0]]), _d(state[p[1]]))
        for i in ids:
            nodes, normals, conds = border[i]
            L.gcmo_simplex_plain_border(model, M, len(nodes), _i(nodes), _d(normals), _i(conds), _i(types), _d(b_next), _d(state[i]))
        summ = "splitting summ" in text   # SplittingType::SUMM: every stage from the same layer, then the average
        layers = []
        for s in range(3):
            nxt = [np.zeros_like(p) for p in state]
            hs = [L.gcmo_sx_begin(ctypes.byref(views[i]), model, M, s, tau, _d(mats[i][0]), _d(mats[i][1]), _d(mats[i][2]),
                                  _d(basis_m), _d(state[i]), _d(nxt[i]), gcm_type) for i in ids]
            for h in hs:
                L.gcmo_sx_nodes(h, 0)
            for p, (fa, fb, fn) in contacts.items():
                L.gcmo_sx_contact_correct(hs[p[0]], hs[p[1]], len(fa), _i(fa), _i(fb), _d(fn))
            for i in ids:
                nodes, normals, conds = border[i]
                L.gcmo_sx_border_correct(hs[i], len(nodes), _i(nodes), _d(normals), _i(conds), 2, _i(types), _d(b_next))
            for h in hs:
                L.gcmo_sx_nodes(h, 1)
            for h in hs:
                assert L.gcmo_sx_end(h) == 0, (step, s)
            if summ:
                layers.append(nxt)
            else:
                state = nxt
        if summ:   # DefaultMesh::averageNewPdeLayersToCurrent: pde = 0; pde += new_s / 3
            state = [((0.0 + layers[0][i] / 3.0) + layers[1][i] / 3.0) + layers[2][i] / 3.0 for i in ids]
        time += tau
        assert eng.info()[1] == time
        for i in ids:
            got = eng.simplex_pde(i)
            assert np.array_equal(got, state[i]), (step, i, np.abs(got - state[i]).max())
    assert eng.errors() == 0
    assert all(np.isfinite(p).all() and np.abs(p).max() < 10 for p in state)
    if golden is not None:
        assert time == float(golden["time"]) and tau == float(golden["tau"])
        for i in ids:
            assert np.array_equal(state[i], golden["pde%d" % i]), "the restatement differs from the reference engine"
    eng.close()


def check_oracle_against_reference(lib, name, workdir):
    """oracle/simplex_oracle.c (driven in the reference's order from the host engine's set-up) reproduces the
    unmodified reference engine's final state bit for bit: this pins the restatement"""
    g = load_golden(name)
    basis = "identity" if "basis 1 0 0 0 1 0 0 0 1" in str(g["task"]) else "rotated"
    check_engine(lib, int(g["model"]), bodies=int(g["bodies"]), basis=basis, steps=int(g["steps"]), gcm_type=int(g["gcm_type"]),
                 text=golden_task_with_mesh(g, workdir), golden=g)


# ---- golden vectors from the unmodified reference engine (tests/golden/make_simplex_golden.py) -------------------
def golden_scenario(model, bodies, gcm, basis, steps):
    """engine_scenario on a 7x6x8 box with a cavity of about 100 cells that survives the reference's clean-up"""
    text, _ = engine_scenario(model, bodies=min(bodies, 2), steps=steps, gcm_type=gcm, cavity=True, basis=basis)
    if bodies == 3:
        # a third body along x: lines where three bodies meet (multicontact vertices without empty space around)
        name = "elastic" if model == 0 else "acoustic"
        text += "body 2 %s isotropic\nmaterial body 2 isotropic 1.5 2.5 %r\nregion 2 box 2.0 -10 -10 10 10 10\n" % (name, 0.9 if model == 0 else 0.0)
    lines = []
    for ln in text.split("\n"):
        if ln.startswith("simplex_box"):
            ln = "simplex_box 7 6 8 0 0 0 0.5 jitter 0.3 seed 7"
        if ln.startswith("cavity"):
            ln = "cavity box 1.1 0.9 1.6 2.3 2.1 2.9"
        if ln.startswith("region 0"):
            ln = "region 0 box -10 -10 -10 10 10 2.0"
        if ln.startswith("region 1"):
            ln = "region 1 box -10 -10 2.0 10 10 10"
        if ln.startswith("border_condition box"):
            ln = ln.replace("2.999", "3.999")
        if ln.startswith("initial quantity"):
            ln = "initial quantity PRESSURE 1 sphere 0.9 1.6 1.4 1.8"
        lines.append(ln)
    return "\n".join(lines)


def cube_scenario(acoustic, n=6, steps=7):
    """The reference launcher's cubeAcs / cubeEls tasks (src/launcher/main.cpp:547-640) at reduced size: the unit cube (meshed by
    the box mesher instead of CGAL's mesher on meshes/cube.off), Courant number 1, identity basis, zero fixed force on the
    whole border and a loading pulse on the face x = 0 that is switched off at t = 0.25"""
    name = "acoustic" if acoustic else "elastic"
    zero = "const 0" if acoustic else "const 0 const 0 const 0"
    pulse = "until 0.25 1" if acoustic else "const 0 const 0 until 0.25 -1"
    return "\n".join([
        "grid simplex", "dimensionality 3", "courant 1.0", "steps %d" % steps,
        "simplex_box %d %d %d 0 0 0 %r jitter 0.3 seed 3" % (n, n, n, 1.0 / n),
        "body 0 %s isotropic" % name,
        "material body 0 isotropic 4 %s" % ("4 0" if acoustic else "2 1"),
        "basis 1 0 0 0 1 0 0 0 1",
        "border_condition infinite fixed_force " + zero,
        "border_condition box -10 -10 -10 0.01 10 10 fixed_force " + pulse,
        "gcm_type riemann_invariants"]) + "\n"


GOLDEN_SIMPLEX = ["elastic_cavity", "elastic_contact", "acoustic_contact_rotated", "elastic_contact_pde_vectors", "acoustic_pde_vectors",
                  "elastic_contact_summ"]
# BorderCalcMode::LOCAL_BASIS and the Maxwell ODE: pinned on the engine level only (the C restatement keeps to
# GLOBAL_BASIS and has no ODE)
GOLDEN_SIMPLEX_LOCAL = ["elastic_contact_local_basis", "acoustic_cavity_local_basis_pde_vectors", "elastic_maxwell",
                        "elastic_three_bodies", "cube_acs", "cube_els"]


def load_golden(name):
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "simplex_%s.npz" % name))


def golden_task_with_mesh(g, workdir):
    """the fixture's task text reading the fixture's triangulation (with the reference-cleaned body ids) from an INM file"""
    import os
    from simplex_helpers import write_inm_all_cells
    tri = dict(xyz=g["xyz"], cell_v=g["cell_v"], cell_n=g["cell_n"], cell_grid=g["cell_grid"])
    inm = os.path.join(str(workdir), "mesh.inm")
    write_inm_all_cells(tri, g["cell_grid"], inm)
    keep = [ln for ln in str(g["task"]).split("\n") if not ln.startswith(("simplex_box", "region", "cavity"))]
    return "\n".join(keep) + "\nsimplex_mesh %s\n" % inm


def check_engine_against_reference(lib, name, workdir):
    """the product's simplex::Engine == the unmodified reference engine: time step, step count and every PDE value
    of every body, bit for bit"""
    g = load_golden(name)
    eng = capi.SimplexHostEngine(lib, golden_task_with_mesh(g, workdir))
    tri = eng.triangulation()
    assert np.array_equal(tri["cell_n"], g["cell_n"])        # the topology built from the file is the fixture's
    eng.run()
    steps, time, tau = eng.info()
    assert (steps, time, tau) == (int(g["steps"]), float(g["time"]), float(g["tau"]))
    if int(g["bodies"]) == 3:   # every pair of bodies is in contact, and there are vertices where all three meet
        assert all(len(eng.contact_nodes(a, b)[0]) > 0 for a, b in ((0, 1), (0, 2), (1, 2)))
    for b in range(int(g["bodies"])):
        assert eng.simplex_body_info(b)["average_height"] == float(g["average_height%d" % b])
        got = eng.simplex_pde(b)
        ref = g["pde%d" % b]
        assert got.shape == ref.shape and np.abs(ref).max() > 0.05
        assert np.array_equal(got, ref), (name, b, np.abs(got - ref).max())
    assert eng.errors() == 0
    eng.close()


def check_locate_against_reference(lib, with_oracle=False):
    """gcmb_simplex_locate == SimplexGrid::findCellCrossedByTheRay of the UNMODIFIED reference
    (grid/simplex/SimplexGrid.cpp:61-112) over the protocol of src/test/sequence/TestLineWalkSearch3D.cpp:120-154:
    16 x 16 directions x 9 lengths from every third vertex of both bodies of the elastic_contact fixture,
    428 544 queries, integer for integer"""
    import os
    g = load_golden("elastic_contact")
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "simplex_locate_protocol.npz"))
    tri = dict(xyz=np.ascontiguousarray(g["xyz"]), cell_v=np.ascontiguousarray(g["cell_v"]), cell_n=np.ascontiguousarray(g["cell_n"]),
               cell_grid=np.ascontiguousarray(g["cell_grid"]))
    # incident cells of every vertex in ascending cell id
    order = np.argsort(tri["cell_v"].ravel(), kind="stable")
    counts = np.bincount(tri["cell_v"].ravel(), minlength=len(tri["xyz"]))
    tri["inc_off"] = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    tri["inc_cell"] = (order // 4).astype(np.int32)
    ctx = capi.Context(lib)
    L = oracle() if with_oracle else None
    for body in (0, 1):
        m = Mesh.from_arrays(tri, body)
        vs = ref["vertices%d" % body]
        v, sh = protocol_queries(m, 16, 9, scale=0.4, vertices=vs)
        want = ref["located%d" % body].astype(np.int32)
        assert len(want) == len(v) and (want[:, 0] >= 0).all()
        sb = SimplexBody(lib, ctx, m, 0)
        got = sb.locate(v, sh)
        assert sb.errors() == 0
        assert np.array_equal(got, want), "cell location differs from the reference in %d of %d queries" % ((got != want).any(axis=1).sum(), len(v))
        sb.close()
        if with_oracle:
            pick = np.arange(0, len(v), 7)
            o, errs = oracle_locate_all(L, m, v[pick], sh[pick])
            assert errs == 0 and np.array_equal(o, want[pick])
    ctx.close()


def check_mesh_cleanup_against_reference(lib, name):
    """the box mesher + the host's restatement of the reference's clean-up of body ids == the ids the reference's
    CgalTriangulation constructor leaves (grid/simplex/cgal/CgalTriangulation.cpp:8-112), and the task run straight
    from the mesher then reproduces the reference engine's values"""
    g = load_golden(name)
    # the clean-up draws from libc's rand() like the reference's, which runs it in a fresh process: same state here
    ctypes.CDLL(None).srand(1)
    eng = capi.SimplexHostEngine(lib, str(g["task"]))
    tri = eng.triangulation()
    assert np.array_equal(tri["xyz"], g["xyz"]) and np.array_equal(tri["cell_v"], g["cell_v"])
    assert (g["cell_grid"] != g["cell_grid_before_cleanup"]).sum() > 0      # the clean-up had something to do
    assert np.array_equal(tri["cell_grid"], g["cell_grid"])
    eng.run()
    for b in range(int(g["bodies"])):
        assert np.array_equal(eng.simplex_pde(b), g["pde%d" % b])
    eng.close()
