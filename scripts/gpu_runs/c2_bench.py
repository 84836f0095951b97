"""BASELINE config 2: 3-D acoustic n^3 (default 512; 1024 also fits), point source, PRESSURE -> 0 on all six
faces, border size 2, Courant 0.9, fp64.  Node-updates/s of Engine::run's loop, device-timed over K steps after 3 warm-up steps,
state resident in HBM.  Algorithmic bytes: 3 stages x 2 x M x 8 B = 192 B per node-update (M = 4).
Usage: python scripts/gpu_runs/c2_bench.py [edge] [steps]"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import gcm_b200
    from gcm_b200 import capi
    from scenarios import acoustic3d_free
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    W = 3
    text = acoustic3d_free(n, 10 ** 6).replace("sphere 0.2 0.5 0.5 0.5", "sphere 0.05 0.5 0.5 0.5")
    lib = gcm_b200.library()
    os.chdir(tempfile.mkdtemp(prefix="gcmb_c2_"))
    eng = capi.HostEngine(lib, text, device=0)
    ctxh = eng.context_handle()
    eng.advance(W)
    kernels = [eng.kernel_name(0, d) for d in range(3)]
    lib.check(lib.c.gcmb_sync(ctxh))
    torch.cuda.synchronize()
    launches0 = lib.c.gcmb_launch_count(ctxh)
    lib.check(lib.c.gcmb_timer_start(ctxh))
    eng.advance(K)
    ms = capi.ctypes.c_float(0)
    lib.check(lib.c.gcmb_timer_stop(ctxh, capi.ctypes.byref(ms)))
    launches = lib.c.gcmb_launch_count(ctxh) - launches0
    per_s = n ** 3 * K / (ms.value * 1e-3)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    print(json.dumps({"workload": "BASELINE config 2: 3-D acoustic %d^3, PRESSURE -> 0 on six faces, fp64, bs 2" % n, "kernels": kernels,
                      "steps": K, "warmup": W, "ms_per_step": ms.value / K, "gpu_launches": launches, "node_updates_per_s": per_s,
                      "roofline": {"bound": "hbm", "achieved": per_s * 192 / 1e9, "peak": peak, "unit": "GB/s", "frac": per_s * 192 / 1e9 / peak}}))
    eng.close()


if __name__ == "__main__":
    main()
