"""TriangleInterpolator (reference util/math/interpolation/TriangleInterpolator.hpp:8-130): the C
restatement (oracle/simplex_oracle.c), the product's per-thread function on the stepping harness and on the GPU
(gcmb_triangle_interpolate) against
  * 20 000 queries answered by the reference's OWN class (oracle/_ref/gcm_ref_interp; tests/golden/triangle_interpolator.npz,
    made by tests/golden/make_triangle_golden.py): values and "the reference throws" flags, bit for bit;
  * the known answers of the reference's tests (src/test/sequence/TestInterpolator.cpp:129-189,267-274)."""
import ctypes
import os

import numpy as np
import pytest

from gcm_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "triangle_interpolator.npz"))
dp = ctypes.POINTER(ctypes.c_double)
ip = ctypes.POINTER(ctypes.c_int)


def oracle_interpolate(mode, points, values, grads, queries):
    import oracle_host as oh
    L = oh.lib()
    L.gcmo_triangle_interpolate.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, dp, ip]
    L.gcmo_triangle_interpolate.restype = None
    points, values, queries = (np.ascontiguousarray(a, dtype=np.float64) for a in (points, values, queries))
    g = None if grads is None else np.ascontiguousarray(grads, dtype=np.float64)
    out = np.zeros(len(queries))
    status = np.zeros(len(queries), dtype=np.int32)
    L.gcmo_triangle_interpolate(mode, len(queries), points.ctypes.data_as(dp), values.ctypes.data_as(dp),
                                None if g is None else g.ctypes.data_as(dp), queries.ctypes.data_as(dp), out.ctypes.data_as(dp), status.ctypes.data_as(ip))
    return out, status


def check_against_reference(fn):
    for mode in range(5):
        grads = GOLD["grads%d" % mode] if "grads%d" % mode in GOLD.files else None
        out, status = fn(mode, GOLD["points%d" % mode], GOLD["values%d" % mode], grads, GOLD["queries%d" % mode])
        assert np.array_equal(status, GOLD["status%d" % mode]), mode
        assert np.array_equal(out, GOLD["out%d" % mode]), (mode, np.abs(out - GOLD["out%d" % mode]).max())
        assert 0 < status.sum() < len(status)


def check_known_answers(fn):
    rng = np.random.default_rng(5)
    n = 1000
    tri = rng.uniform(-1e6, 1e6, size=(n, 3, 2))
    lam = rng.uniform(0, 1, size=(n, 3))
    lam /= lam.sum(axis=1, keepdims=True)
    q = np.einsum("nk,nkd->nd", lam, tri)
    # TEST(TriangleInterpolator, linear): f = 5x + 8y - 2 is reproduced; a point outside throws
    f = lambda x: 5 * x[..., 0] + 8 * x[..., 1] - 2
    out, status = fn(0, tri, f(tri), None, q)
    assert not status.any() and np.all(np.abs(out - f(q)) <= 1e-9 * np.abs(f(q)) + 1e-6)
    outside = -2 * tri[:, 0] + tri[:, 1] + 2 * tri[:, 2]
    assert fn(0, tri, f(tri), None, outside)[1].all()
    # TEST(TriangleInterpolator, quadratic): exact for a quadratic polynomial with its gradients
    f2 = lambda x: 8 * x[..., 0] ** 2 + 10 * x[..., 0] * x[..., 1] - 15 * x[..., 1] ** 2 + 5 * x[..., 0] + 8 * x[..., 1] - 2
    g2 = lambda x: np.stack([16 * x[..., 0] + 10 * x[..., 1] + 5, -30 * x[..., 1] + 10 * x[..., 0] + 8], axis=-1)
    out, status = fn(1, tri, f2(tri), g2(tri), q)
    assert not status.any() and np.all(np.abs(out - f2(q)) <= 1e-9 * np.abs(f2(q)) + 1e-3)
    assert fn(1, tri, f2(tri), g2(tri), 3 * tri[:, 0] - tri[:, 1] - tri[:, 2])[1].all()
    # TEST(TriangleInterpolator, interpolateInOwner) == 1 exactly
    out, status = fn(4, [[[0, 0], [0, 1], [1, 0], [1, 1]]], [[1, 1, 1, 1e100]], None, [[0.2, 0.2]])
    assert status[0] == 0 and out[0] == 1
    # TEST(TriangleInterpolator, quadraticMinMax) == 1 exactly: f = x^2 + y^2, limited to the vertex values
    out, status = fn(2, [[[0, 1], [1, 0], [-1, -1]]], [[1, 1, 2]], [[[0, 2], [2, 0], [-2, -2]]], [[0, 0]])
    assert status[0] == 0 and out[0] == 1


def test_oracle_restatement_matches_the_reference_class():
    check_against_reference(oracle_interpolate)
    check_known_answers(oracle_interpolate)


def test_product_function_on_the_stepping_harness():
    from helpers import emul_library
    ctx = capi.Context(emul_library())
    fn = lambda *a: capi.triangle_interpolate(ctx, *a)
    check_against_reference(fn)
    check_known_answers(fn)
    ctx.close()


@pytest.mark.gpu
def test_gpu_triangle_interpolator_matches_the_reference_class():
    import gcm_b200
    ctx = capi.Context(gcm_b200.library())
    fn = lambda *a: capi.triangle_interpolate(ctx, *a)
    check_against_reference(fn)
    check_known_answers(fn)
    ctx.close()
