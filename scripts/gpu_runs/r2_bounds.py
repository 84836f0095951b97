"""Measured deviation bounds of the non-bit-exact modes against the fp64 bit-exact run (== the reference, bitwise):
fp32 (gcmb_create(..., 4)) and fp64 with FMA contraction (gcmb_set_fma), max|g - r| / max|r| after 1, 10, 100 steps."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import gcm_b200
from gcm_b200 import capi
from scenarios import acoustic3d_free, elastic3d_layers, ortho3d_contact, elastic3d_ortho_rotated

lib = gcm_b200.library()
os.chdir("/tmp")
out = {}
for name, text in (("elastic3d_layers 64^3", elastic3d_layers(n=64, steps=1000)), ("acoustic3d_free 64^3", acoustic3d_free(n=64, steps=1000)),
                   ("ortho3d_contact 48^3", ortho3d_contact(n=48, steps=1000)), ("elastic3d_ortho_rotated 32^3", elastic3d_ortho_rotated(n=32, steps=1000)),
                   ("elastic3d_layers 64^3 Courant 1", elastic3d_layers(n=64, steps=1000, courant=1.0))):
    ref = capi.HostEngine(lib, text)
    f32 = capi.HostEngine(lib, text, real_bytes=4)
    fma = capi.HostEngine(lib, text, fma=True)
    done, rows = 0, {}
    for n in (1, 10, 100):
        for e in (ref, f32, fma):
            e.advance(n - done)
        done = n
        r = ref.body_pde(0)
        scale = np.abs(r).max()
        rows[n] = {"fp32": float(np.abs(f32.body_pde(0) - r).max() / scale), "fma": float(np.abs(fma.body_pde(0) - r).max() / scale)}
    out[name] = rows
    for e in (ref, f32, fma):
        e.close()
    print("BOUNDS", name, json.dumps(rows), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_bounds.json"), "w"), indent=1)
