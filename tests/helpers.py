"""Shared helpers of the parity tests: run a task through the gcm_b200 host engine and compare."""
import os

import numpy as np

from gcm_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


def emul_library():
    """The stepping harness (tests/emul): the product's CUDA sources stepped on the CPU.  Tests only."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("build_emul", os.path.join(ROOT, "tests", "emul", "build_emul.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return capi.Library(cuda_path=mod.build_emul(), host_path=mod.build_host_emul())


def run_engine(lib, task_text, **kw):
    eng = capi.HostEngine(lib, task_text, **kw)
    eng.run()
    return eng


def compare_with_golden(lib, name, task_text, exact=True):
    g = golden(name)
    eng = run_engine(lib, task_text)
    try:
        steps, time, tau = eng.info()
        assert steps == int(g["steps"]), (steps, int(g["steps"]))
        assert tau == float(g["tau"])
        assert time == float(g["time"])
        bid = 0
        worst = 0.0
        while "body%d" % bid in g.files:
            ref = g["body%d" % bid]
            got = eng.body_pde(bid)
            assert got.shape == ref.shape
            if exact:
                assert np.array_equal(ref, got), "body %d: max|diff| = %g (%s)" % (
                    bid, np.abs(ref - got).max(), [eng.kernel_name(bid, d) for d in range(eng.body_info(bid)[0])])
            scale = max(np.abs(ref).max(), 1e-300)
            worst = max(worst, np.abs(ref - got).max() / scale)
            U, U1, L = eng.body_matrices(bid)
            flat = np.concatenate([np.concatenate([U[t, s].ravel(), U1[t, s].ravel(), L[t, s]])
                                   for t in range(U.shape[0]) for s in range(U.shape[1])])
            assert np.array_equal(flat, g["mat%d" % bid]), "eigen-systems differ from the reference's"
            bid += 1
        assert worst <= 1e-12  # north_star tolerance for fp64
        if "detector" in g.files:
            t, v = eng.seismogram()
            det = g["detector"]
            assert len(t) == det.shape[0]
            assert np.allclose(t, det[:, 0], rtol=5e-6, atol=1e-30)
            assert np.allclose(v, det[:, 1], rtol=5e-6, atol=1e-30)
        return eng, worst
    except Exception:
        eng.close()
        raise


def random_stage_check(lib, cases, seed=7, real_bytes=8, n_materials=3):
    """C-ABI level: random state with ghosts, random material map, every direction, bitwise vs gcmo_stage
    (real_bytes=4: the fp32 kernels, within single-precision rounding of the fp64 oracle)."""
    import oracle_host as oh
    L = oh.lib()
    rng = np.random.default_rng(seed)
    ctx = capi.Context(lib, real_bytes=real_bytes)
    for (D, sizes, model, bs) in cases:
        mats = [{"kind": "isotropic", "rho": rng.uniform(1, 5), "lambda": rng.uniform(1, 5), "mu": rng.uniform(0.5, 3)}
                for _ in range(n_materials)]
        ms = [oh.matrices_for(model, D, m) for m in mats]
        U = np.ascontiguousarray(np.stack([m[0] for m in ms]))
        U1 = np.ascontiguousarray(np.stack([m[1] for m in ms]))
        Lm = np.ascontiguousarray(np.stack([m[2] for m in ms]))
        M = U.shape[-1]
        h = rng.uniform(0.5, 1.5, D)
        full = tuple(s + 2 * bs for s in sizes)
        state = rng.normal(size=full + (M,))
        table_full = rng.integers(0, n_materials, size=full).astype(np.uint8)
        real = tuple(slice(bs, bs + s) for s in sizes)
        body = capi.CubicBody(ctx, D, M, sizes, [0] * D, h, bs)
        body.set_materials(U, U1, Lm, np.ascontiguousarray(table_full[real]))
        tau = 0.4 * h.min() / np.abs(Lm).max()
        sz = np.array(sizes, dtype=np.int32)
        for s in range(D):
            body.upload(state, with_ghosts=True)
            body.stage(s, tau)
            got = body.download(with_ghosts=False)
            nxt = np.zeros_like(state)
            rc = L.gcmo_stage(D, M, oh._ip(sz), bs, oh._dp(h), s, tau, n_materials, oh._dp(U), oh._dp(U1), oh._dp(Lm),
                              oh._bp(table_full), oh._dp(state), oh._dp(nxt))
            assert rc == 0
            if real_bytes == 4:
                assert np.abs(nxt[real] - got).max() <= 2e-5 * np.abs(nxt[real]).max(), (D, sizes, model, bs, s, body.kernel_name(s))
                continue
            assert np.array_equal(nxt[real], got), (D, sizes, model, bs, s, body.kernel_name(s),
                                                    np.abs(nxt[real] - got).max())
        body.close()
    ctx.close()


def fused_border_check(lib, seed=11):
    """gcmb_cubic_stage_fill_next_border == gcmb_cubic_stage + gcmb_cubic_border_apply, ghost nodes included, bit for bit;
    returns the number of cases that took the fused path"""
    rng = np.random.default_rng(seed)
    ctx = capi.Context(lib)
    fused_cases = 0
    # (D, sizes, model, border size, conditions of the last direction: (quantity codes, which sides))
    cases = ((3, (6, 37, 64), "elastic", 2, [((5, 7, 8), 3)]), (3, (5, 9, 128), "elastic", 2, [((2,), 1), ((5, 7, 8), 2)]),
             (3, (4, 300, 96), "acoustic", 2, [((3,), 3)]), (2, (45, 160), "elastic", 2, [((3, 4), 3)]),
             (3, (5, 7, 64), "elastic", 3, [((5, 7, 8), 3), ((0, 8), 3)]), (3, (5, 7, 64), "elastic", 1, [((8,), 2)]),
             (3, (6, 37, 50), "elastic", 2, [((5, 7, 8), 3)]))
    for (D, sizes, model, bs, conds) in cases:
        mats = [("isotropic", rng.uniform(1, 5), rng.uniform(1, 5), rng.uniform(0.5, 3) if model == "elastic" else 0.0) for _ in range(3)]
        ms = [capi.host_matrices(lib, model, D, m) for m in mats]
        U, U1, Lm = (np.ascontiguousarray(np.stack([m[i] for m in ms])) for i in range(3))
        M = U.shape[-1]
        h = rng.uniform(0.5, 1.5, D)
        full = tuple(s + 2 * bs for s in sizes)
        state = rng.normal(size=full + (M,))
        table = rng.integers(0, 3, size=sizes).astype(np.uint8)
        tau = 0.4 * h.min() / np.abs(Lm).max()
        last, prev = D - 1, D - 2
        results = []
        for fused_call in (True, False):
            body = capi.CubicBody(ctx, D, M, sizes, [0] * D, h, bs)
            body.set_materials(U, U1, Lm, table)
            values = []
            for i, (codes, sides) in enumerate(conds):
                body.border_set_area(i, last, ("infinite",), list(codes), sides=sides)
                values += [0.25 * (i + 1) + 0.1 * c for c in codes]
            body.upload(state, with_ghosts=True)
            if fused_call:
                took = body.stage_fill_next_border(prev, tau, last, values)
                fused_cases += int(took)
                if not took:
                    body.border_apply(last, values)
            else:
                body.stage(prev, tau)
                body.border_apply(last, values)
            results.append(body.download(with_ghosts=True))
            body.close()
        real = tuple(slice(bs, bs + s) for s in sizes[:-1])  # real nodes of every axis but the last, whole rows
        assert np.array_equal(results[0][real], results[1][real]), (D, sizes, model, bs, np.abs(results[0][real] - results[1][real]).max())
    ctx.close()
    return fused_cases


def tile_border_check(lib, real_bytes=8, seed=12):
    """gcmb_cubic_stage_with_border == gcmb_cubic_border_apply + gcmb_cubic_stage on all real nodes, bit for bit, at row lengths
    around the tile and warp boundaries (ghost nodes in the halo of the tile before the last one, one-sided conditions,
    several conditions per face, rows shorter than a warp); returns the number of cases that took the in-tile path"""
    rng = np.random.default_rng(seed)
    ctx = capi.Context(lib, real_bytes=real_bytes)
    fused_cases = 0
    el = (5, 7, 8)
    cases = ((3, (3, 5, 33), "elastic", 2, [(el, 3)]), (3, (3, 4, 34), "elastic", 2, [(el, 3)]), (3, (3, 5, 65), "elastic", 3, [(el, 3)]),
             (3, (2, 3, 257), "elastic", 2, [(el, 3)]), (3, (2, 3, 258), "acoustic", 2, [((3,), 3)]), (3, (2, 3, 257), "elastic", 1, [((2,), 3)]),
             (3, (2, 3, 256), "elastic", 2, [(el, 2)]), (3, (3, 3, 255), "elastic", 3, [((0, 1, 2), 1), (el, 2)]),
             (2, (5, 513), "elastic", 2, [((3, 4), 1)]), (2, (5, 64), "elastic", 2, [((3, 4), 3), ((0, 1), 3)]),
             (1, (1000,), "acoustic", 1, [((1,), 3)]), (1, (31,), "acoustic", 2, [((0,), 3)]),
             (3, (4, 6, 3), "elastic", 2, [(el, 3)]), (3, (3, 3, 2), "elastic", 2, [(el, 3)]), (3, (2, 40, 96), "elastic", 2, [(el, 3)]))
    for (D, sizes, model, bs, conds) in cases:
        mats = [("isotropic", rng.uniform(1, 5), rng.uniform(1, 5), rng.uniform(0.5, 3) if model == "elastic" else 0.0) for _ in range(3)]
        ms = [capi.host_matrices(lib, model, D, m) for m in mats]
        U, U1, Lm = (np.ascontiguousarray(np.stack([m[i] for m in ms])) for i in range(3))
        M = U.shape[-1]
        h = rng.uniform(0.5, 1.5, D)
        full = tuple(s + 2 * bs for s in sizes)
        state = rng.normal(size=full + (M,))
        table = rng.integers(0, 3, size=sizes).astype(np.uint8)
        tau = 0.4 * h.min() / np.abs(Lm).max()
        last = D - 1
        results = []
        for one_call in (True, False):
            body = capi.CubicBody(ctx, D, M, sizes, [0] * D, h, bs)
            body.set_materials(U, U1, Lm, table)
            values = []
            for i, (codes, sides) in enumerate(conds):
                body.border_set_area(i, last, ("infinite",), list(codes), sides=sides)
                values += [0.25 * (i + 1) + 0.1 * c for c in codes]
            body.upload(state, with_ghosts=True)
            if one_call:
                took = body.stage_with_border(last, tau, values)
                assert took == (sizes[-1] > bs), (D, sizes, model, bs, body.kernel_name(last))
                fused_cases += int(took)
            else:
                body.border_apply(last, values)
                body.stage(last, tau)
            results.append(body.download())
            body.close()
        assert np.array_equal(results[0], results[1]), (D, sizes, model, bs, np.abs(results[0] - results[1]).max())
    ctx.close()
    return fused_cases


def read_vtk_appended(path):
    """arrays of a VTK XML file written with one raw appended block (gcm_b200/host/vtk_writer.cpp)"""
    import re
    raw = open(path, "rb").read()
    head, _, tail = raw.partition(b'<AppendedData encoding="raw">')
    data = tail[tail.index(b"_") + 1:]
    types = {"Float32": np.float32, "Int32": np.int32, "UInt8": np.uint8}
    out = {"header": head.decode()}
    for m in re.finditer(r'<DataArray type="(\w+)" Name="(\w+)" NumberOfComponents="(\d+)" format="appended" offset="(\d+)"/>', out["header"]):
        kind, name, comps, offset = m.group(1), m.group(2), int(m.group(3)), int(m.group(4))
        nbytes = int(np.frombuffer(data[offset:offset + 4], dtype=np.uint32)[0])
        arr = np.frombuffer(data[offset + 4:offset + 4 + nbytes], dtype=types[kind])
        out[name] = arr.reshape(-1, comps) if comps > 1 else arr
    return out
