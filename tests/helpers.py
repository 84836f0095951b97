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
            assert np.allclose(t, det[:, 0], rtol=2e-6, atol=1e-30)
            assert np.allclose(v, det[:, 1], rtol=2e-6, atol=1e-30)
        return eng, worst
    except Exception:
        eng.close()
        raise
