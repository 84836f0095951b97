"""The reference's own engine tests (src/test/sequence/TestEngine.cpp) restated on gcm_b200's host engine.
Used twice: on the stepping harness (CPU, tests/test_host_logic_emul.py) and on the CUDA library (gpu)."""
import math

import numpy as np

from gcm_b200 import capi


def run_statement(lib):
    """TestEngine.cpp:91-136: an S-wave packet translates 19 cells in 9 steps at Courant 4.5, border size 5."""
    text = """
dimensionality 2
courant 4.5
border_size 5
h %r %r
steps 9
body 0 elastic isotropic sizes 20 40 start 0 0
material default isotropic 4 2 0.5
initial wave S1_FORWARD 1 Vx 1 box -1 0.1125 -1 8 0.6375 1
""" % (7.0 / 19, 3.0 / 39)
    eng = capi.HostEngine(lib, text)
    expected = eng.body_pde(0).reshape(20, 40, 5)[10, 3].copy()
    assert np.any(expected != 0)
    eng.run()
    assert eng.info()[0] == 9
    actual = eng.body_pde(0).reshape(20, 40, 5)[10, 22]
    # linal::approximatelyEqual with EQUALITY_TOLERANCE = 1e-9 (util/infrastructure/Types.hpp:10)
    assert np.allclose(expected, actual, rtol=0, atol=1e-9 * max(1.0, np.abs(expected).max())), (expected, actual)
    names = [eng.kernel_name(0, d) for d in range(2)]
    eng.close()
    return names


def two_layers(lib, vary):
    """TestEngine.cpp:139-296: reflection of a P-wave from an interface for five contrasts of density (vary='rho')
    or stiffness (vary='E'); Courant 1.5, border size 3, run until t = 0.24."""
    results = []
    for i in range(5):
        rho0, lambda0, mu0 = 1.0, 2.0, 0.8
        k = 0.25 * 2 ** i
        rho, lam, mu = (k * rho0, lambda0, mu0) if vary == "rho" else (rho0, k * lambda0, k * mu0)
        text = """
dimensionality 2
courant 1.5
border_size 3
h %r %r
required_time 0.24
body 0 elastic isotropic sizes 50 100 start 0 0
material default isotropic %r %r %r
material area box -10 %r -10 10 10 10 isotropic %r %r %r
initial wave P_FORWARD 1 Vy -2 box -1 0.015 -1 4 0.455 1
""" % (2.0 / 49, 1.0 / 99, rho0, lambda0, mu0, 0.5 - 1e-5, rho, lam, mu)
        eng = capi.HostEngine(lib, text)
        init = eng.body_pde(0).reshape(50, 100, 5)[25, 25].copy()
        assert np.any(init != 0)
        eng.run()
        reflect = eng.body_pde(0).reshape(50, 100, 5)[25, 25].copy()
        steps = eng.info()[0]
        eng.close()
        E0 = mu0 * (3 * lambda0 + 2 * mu0) / (lambda0 + mu0)
        Z0 = math.sqrt(E0 * rho0)
        E = mu * (3 * lam + 2 * mu) / (lam + mu)
        Z = math.sqrt(E * rho)
        # sigma(1,1) is component 4, velocity(1) component 1 of the 2-D elastic PDE vector
        results.append((steps, reflect[4] / init[4], (Z - Z0) / (Z + Z0), reflect[1] / init[1], (Z0 - Z) / (Z + Z0),
                        init, reflect))
    return results
