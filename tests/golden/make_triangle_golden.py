"""tests/golden/triangle_interpolator.npz: seeded random queries answered by the reference's OWN TriangleInterpolator<real>
(oracle/_ref/gcm_ref_interp = src/libgcm/util/math/interpolation/TriangleInterpolator.hpp compiled unmodified).
Run where /root/reference exists:  make -C oracle ref_interp && python tests/golden/make_triangle_golden.py"""
import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
EXE = os.path.join(ROOT, "oracle", "_ref", "gcm_ref_interp")


def queries(mode, n, rng):
    np_ = 4 if mode == 4 else 3
    scale = 10.0 ** rng.integers(-3, 7, size=(n, 1, 1))
    points = rng.uniform(-1, 1, size=(n, np_, 2)) * scale
    values = rng.normal(size=(n, np_)) * 10.0 ** rng.integers(-2, 3, size=(n, 1))
    grads = rng.normal(size=(n, 3, 2)) / scale if mode in (1, 2, 3) else None
    lam = rng.uniform(0, 1, size=(n, 3))
    lam /= lam.sum(axis=1, keepdims=True)
    # every fourth query lies outside (the reference throws), every fifth on an edge, every seventh in a vertex
    lam[::4] = rng.uniform(-1.5, 1.5, size=lam[::4].shape)
    lam[::5, 2] = 0.0
    lam[::5, :2] /= np.maximum(lam[::5, :2].sum(axis=1, keepdims=True), 1e-300)
    lam[::7] = np.array([1.0, 0.0, 0.0])
    q = np.einsum("nk,nkd->nd", lam, points[:, :3])
    if mode == 4:
        q[1::2] = 0.5 * (points[1::2, 1] + points[1::2, 3]) * 0.7 + 0.3 * points[1::2, 2]   # often only in a later triangle
    # degenerate triangles now and then
    points[3::97, 2] = points[3::97, 0] + 2.0 * (points[3::97, 1] - points[3::97, 0])
    return points, values, grads, q


def run_reference(mode, points, values, grads, q):
    n = len(q)
    with tempfile.TemporaryDirectory() as tmp:
        fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("ii", mode, n))
            f.write(np.ascontiguousarray(points).tobytes())
            f.write(np.ascontiguousarray(values).tobytes())
            if grads is not None:
                f.write(np.ascontiguousarray(grads).tobytes())
            f.write(np.ascontiguousarray(q).tobytes())
        subprocess.run([EXE, fin, fout], check=True)
        raw = open(fout, "rb").read()
    out = np.frombuffer(raw[:8 * n], dtype=np.float64).copy()
    status = np.frombuffer(raw[8 * n:8 * n + 4 * n], dtype=np.int32).copy()
    return out, status


def main():
    rng = np.random.default_rng(2024)
    arrays = {}
    for mode in range(5):
        points, values, grads, q = queries(mode, 4000, rng)
        out, status = run_reference(mode, points, values, grads, q)
        arrays.update({"points%d" % mode: points, "values%d" % mode: values, "queries%d" % mode: q, "out%d" % mode: out, "status%d" % mode: status})
        if grads is not None:
            arrays["grads%d" % mode] = grads
        print("mode", mode, "queries", len(q), "thrown", int(status.sum()))
    np.savez_compressed(os.path.join(HERE, "triangle_interpolator.npz"), **arrays)


if __name__ == "__main__":
    main()
