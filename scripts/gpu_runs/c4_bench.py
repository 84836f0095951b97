"""BASELINE config 4 at full size on one GPU: four stacked bodies n x n/4 x n along y, the carbon-fibre
composite of launcher/ndi.hpp:120-131 alternating with titanium written as an orthotropic material (ndi.hpp:136-159),
automatic ADHESION contacts, fixed normal velocity on a disc of the top face.  Node-updates/s of Engine::run's loop, timed on
the device over K steps after 3 warm-up steps, state resident in HBM.
Usage: python scripts/gpu_runs/c4_bench.py [edge] [steps]"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

COMPOSITE = "1580 10.30e9 6.96e9 6.96e9 23.25e9 6.96e9 10.30e9 5.01e9 1.67e9 5.01e9"
_E, _NU = 120e9, 0.31
_LAMBDA = _E * _NU / (1 + _NU) / (1 - 2 * _NU)
_MU = _E / (2 + 2 * _NU)
_C11 = _LAMBDA + 2 * _MU
TITANIUM = "4500 %r %r %r %r %r %r %r %r %r" % (_C11, _LAMBDA, _LAMBDA, _C11, _LAMBDA, _C11, _MU, _MU, _MU)


NB = int(os.environ.get("C4_BODIES", "4"))        # experiment knob: number of stacked bodies (4 = BASELINE config 4)
ISO = bool(os.environ.get("C4_ISOTROPIC"))        # experiment knob: the same stack with isotropic materials
ONLY = os.environ.get("C4_ONLY", "")              # experiment knob: "titanium" / "composite" in every body


def task(n, steps=10 ** 6):
    h = repr(1.0 / (n - 1))
    q = n // NB
    lines = ["dimensionality 3", "courant 0.9", "border_size 2", "h %s %s %s" % (h, h, h), "steps %d" % steps]
    for b in range(NB):
        lines.append("body %d elastic %s sizes %d %d %d start 0 %d 0" % (b, "isotropic" if ISO else "orthotropic", n, q, n, b * q))
    for b in range(NB):
        if ISO:
            lines.append("material body %d isotropic %s" % (b, "1580 6.96e9 1.67e9" if b % 2 == 0 else "4500 %r %r" % (_LAMBDA, _MU)))
        else:
            m = {"titanium": TITANIUM, "composite": COMPOSITE}.get(ONLY, COMPOSITE if b % 2 == 0 else TITANIUM)
            lines.append("material body %d orthotropic %s" % (b, m))
    lines.append("initial quantity PRESSURE 1 sphere 0.2 0.5 0.5 0.5")
    lines.append("border %d 1 sphere 0.3 0.5 1.0 0.5 Vy sin 1.0 5.0" % (NB - 1))
    return "\n".join(lines) + "\n"


def main():
    import torch
    import gcm_b200
    from gcm_b200 import capi
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    W = 3
    free, _ = torch.cuda.mem_get_info()
    while n > 128 and NB * (n + 4) * (n // NB + 4) * (n + 4) * (9 * 8 * 2 + 2) > 0.92 * free:
        n -= 128
    lib = gcm_b200.library()
    os.chdir(tempfile.mkdtemp(prefix="gcmb_c4_"))
    eng = capi.HostEngine(lib, task(n), device=0)
    ctxh = eng.context_handle()
    eng.advance(W)
    kernels = [[eng.kernel_name(b, d) for d in range(3)] for b in range(min(2, NB))]
    lib.check(lib.c.gcmb_sync(ctxh))
    torch.cuda.synchronize()
    launches0 = lib.c.gcmb_launch_count(ctxh)
    lib.check(lib.c.gcmb_timer_start(ctxh))
    eng.advance(K)
    ms = capi.ctypes.c_float(0)
    lib.check(lib.c.gcmb_timer_stop(ctxh, capi.ctypes.byref(ms)))
    launches = lib.c.gcmb_launch_count(ctxh) - launches0
    nodes = NB * n * (n // NB) * n
    per_s = nodes * K / (ms.value * 1e-3)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    print(json.dumps({"workload": "BASELINE config 4: %d glued bodies %dx%dx%d, %s, fp64, bs 2" % (NB, n, n // NB, n, "isotropic stand-ins" if ISO else "composite / titanium-as-orthotropic"),
                      "kernels_body0_body1": kernels, "steps": K, "warmup": W, "ms_per_step": ms.value / K, "gpu_launches": launches,
                      "node_updates_per_s": per_s, "roofline": {"bound": "hbm", "achieved": per_s * 432 / 1e9, "peak": peak,
                                                                "unit": "GB/s", "frac": per_s * 432 / 1e9 / peak}}))
    eng.close()


if __name__ == "__main__":
    main()
