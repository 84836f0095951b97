"""Which of the bulk-copy (TMA) stage kernels differ from the oracle?  One random state per case, every direction, printed per
(case, direction).  Run with GCMB_STAGE_IMPL=3 and the GCMB_TMA_* modes to test."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import numpy as np
import gcm_b200
from gcm_b200 import capi
import oracle_host as oh

lib = gcm_b200.library()
L = oh.lib()
rng = np.random.default_rng(7)
ctx = capi.Context(lib)
cases = ((3, (19, 13, 37), "elastic", 2, 0.4), (3, (7, 9, 300), "acoustic", 2, 0.4), (3, (5, 40, 700), "elastic", 2, 0.4),
         (3, (300, 3, 33), "elastic", 2, 0.4), (3, (40, 300, 33), "elastic", 2, 0.4), (3, (40, 300, 33), "elastic", 2, 1.0),
         (2, (23, 131), "elastic", 2, 0.4), (3, (6, 11, 130), "elastic", 3, 0.4), (3, (19, 13, 37), "elastic", 1, 0.4))
for (D, sizes, model, bs, courant) in cases:
    mats = [{"kind": "isotropic", "rho": rng.uniform(1, 5), "lambda": rng.uniform(1, 5), "mu": rng.uniform(0.5, 3)} for _ in range(3)]
    ms = [oh.matrices_for(model, D, m) for m in mats]
    U, U1, Lm = (np.ascontiguousarray(np.stack([m[i] for m in ms])) for i in range(3))
    M = U.shape[-1]
    h = rng.uniform(0.5, 1.5, D)
    full = tuple(s + 2 * bs for s in sizes)
    state = rng.normal(size=full + (M,))
    table_full = rng.integers(0, 3, size=full).astype(np.uint8)
    real = tuple(slice(bs, bs + s) for s in sizes)
    body = capi.CubicBody(ctx, D, M, sizes, [0] * D, h, bs)
    body.set_materials(U, U1, Lm, np.ascontiguousarray(table_full[real]))
    tau = courant * h.min() / np.abs(Lm).max()
    sz = np.array(sizes, dtype=np.int32)
    for s in range(D):
        body.upload(state, with_ghosts=True)
        body.stage(s, tau)
        got = body.download(with_ghosts=False)
        nxt = np.zeros_like(state)
        rc = L.gcmo_stage(D, M, oh._ip(sz), bs, oh._dp(h), s, tau, 3, oh._dp(U), oh._dp(U1), oh._dp(Lm), oh._bp(table_full), oh._dp(state), oh._dp(nxt))
        ref = nxt[real]
        bad = np.argwhere(np.any(ref != got, axis=-1))
        print("case", D, sizes, model, "bs", bs, "courant", courant, "dir", s, body.kernel_name(s), "OK" if len(bad) == 0 else
              "WRONG nodes %d of %d, first %s last %s, comps %s" % (len(bad), ref[..., 0].size, bad[0], bad[-1], np.unique(np.argwhere(ref != got)[:, -1])), flush=True)
    body.close()
ctx.close()
