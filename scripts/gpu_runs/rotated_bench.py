"""Rotated orthotropic plies (BASELINE config 4 with the plies turned +-45 degrees about the stacking
axis): node-updates/s of the dense-eigen-system stage kernels, timed on the device over K steps after W warm-up steps.
Usage: python scripts/gpu_runs/rotated_bench.py [edge] [steps]        (GCMB_DENSE_LITERAL=1 times the literal kernel)"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def task(n, steps=10 ** 6):
    h = repr(1.0 / (n - 1))
    half = n // 2
    ply = "1580 10.30e9 6.96e9 6.96e9 23.25e9 6.96e9 10.30e9 5.01e9 1.67e9 5.01e9"
    return f"""
dimensionality 3
courant 0.9
border_size 2
h {h} {h} {h}
steps {steps}
body 0 elastic orthotropic sizes {n} {half} {n} start 0 0 0
body 1 elastic orthotropic sizes {n} {half} {n} start 0 {half} 0
material body 0 orthotropic {ply} angles 0.2 0.7853981633974483 0
material body 1 orthotropic {ply} angles 0.2 -0.7853981633974483 0
initial quantity PRESSURE 1 sphere 0.35 0.5 0.5 0.5
border 1 1 sphere 0.3 0.5 1.0 0.5 Vy sin 1.0 5.0
"""


def main():
    import torch
    import gcm_b200
    from gcm_b200 import capi
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    W = 3
    lib = gcm_b200.library()
    os.chdir(tempfile.mkdtemp(prefix="gcmb_rot_"))
    eng = capi.HostEngine(lib, task(n), device=0)
    ctxh = eng.context_handle()
    eng.advance(W)
    kernels = [eng.kernel_name(0, d) for d in range(3)]
    lib.check(lib.c.gcmb_sync(ctxh))
    torch.cuda.synchronize()
    lib.check(lib.c.gcmb_timer_start(ctxh))
    eng.advance(K)
    ms = capi.ctypes.c_float(0)
    lib.check(lib.c.gcmb_timer_stop(ctxh, capi.ctypes.byref(ms)))
    nodes = n * (n // 2) * n * 2
    per_s = nodes * K / (ms.value * 1e-3)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    print(json.dumps({"workload": "two glued bodies %dx%dx%d of rotated orthotropic plies, fp64, bs 2" % (n, n // 2, n),
                      "kernels": kernels, "literal": bool(os.environ.get("GCMB_DENSE_LITERAL")), "steps": K, "warmup": W,
                      "ms_per_step": ms.value / K, "node_updates_per_s": per_s,
                      "algorithmic_GBps": per_s * 432 / 1e9, "peaks": peak}))
    eng.close()


if __name__ == "__main__":
    main()
