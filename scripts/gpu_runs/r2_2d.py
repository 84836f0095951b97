"""2-D grids (the reference launcher's seismic / acoustic demos are 2-D): node-updates/s of the engine loop, device-timed.
  python scripts/gpu_runs/r2_2d.py [sizes...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import gcm_b200
from gcm_b200 import capi
import bench
from scenarios import elastic2d_pwave

lib = gcm_b200.library()
os.chdir("/tmp")
peak = bench.measured_hbm_peak()[0]
for n in [int(x) for x in sys.argv[1:]] or [512, 1024, 2048, 4096, 8192]:
    steps = max(20, min(400, (1 << 24) // n))
    eng = capi.HostEngine(lib, elastic2d_pwave(n, 10 ** 6))
    ctxh = eng.context_handle()
    eng.advance(5)
    lib.check(lib.c.gcmb_sync(ctxh))
    lib.check(lib.c.gcmb_profile_enable(ctxh, 1))
    lib.check(lib.c.gcmb_timer_start(ctxh))
    eng.advance(steps)
    ms = capi.ctypes.c_float()
    lib.check(lib.c.gcmb_timer_stop(ctxh, capi.ctypes.byref(ms)))
    prof_ms = np.zeros(8)
    prof_n = np.zeros(8, dtype=np.int64)
    lib.check(lib.c.gcmb_profile_get(ctxh, 8, capi.dp(prof_ms), prof_n.ctypes.data_as(capi.c_ll_p)))
    per = ms.value / steps
    print("2D " + json.dumps({"n": n, "ms_per_step": per, "node_updates_per_s": n * n / (per * 1e-3),
                              "of_hbm_roofline": n * n * 160 / (per * 1e-3) / 1e9 / peak,
                              "class_ms": [float(x) / steps for x in prof_ms[:4]], "kernels": [eng.kernel_name(0, d) for d in range(2)]}), flush=True)
    eng.close()
