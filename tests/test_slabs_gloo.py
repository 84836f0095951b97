"""The N > 1 path on the build machine: two processes (torch.distributed, gloo), each owning one x-slab of a
3-D body on the stepping harness, exchanging ghost planes through the host-staged halo API of the C ABI.
Contract: the slabs put together equal the undivided body bit for bit (the same statement the reference makes
for glued bodies in test/sequence/TestEngine.cpp:27-87), and hence equal the reference fixture."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path[:0] = [%(root)r, os.path.join(%(root)r, "tests"), os.path.join(%(root)r, "oracle")]
import numpy as np
import torch
import torch.distributed as dist
from gcm_b200 import capi
from helpers import emul_library, golden
import oracle_host as oh
from scenarios import SCENARIOS

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lib = emul_library()
name = "elastic3d_layers"
task = oh.parse_task(SCENARIOS[name])
ora = oh.OracleEngine(task)           # only its SETUP (tables, initial state, masks) is used below
b = ora.bodies[0]
D, M, bs = 3, b.M, ora.bs
nx = int(b.sizes[0])
lo, hi = nx * rank // world, nx * (rank + 1) // world
sizes = [hi - lo, int(b.sizes[1]), int(b.sizes[2])]
ctx = capi.Context(lib)
body = capi.CubicBody(ctx, D, M, sizes, [lo, 0, 0], ora.h, bs)
real = b.pde[b.real]                                       # initial state of the whole body
body.set_materials(b.U, b.U1, b.L, np.ascontiguousarray(b.table[b.real][lo:hi]))
body.upload(np.ascontiguousarray(real[lo:hi]), with_ghosts=False)
for ci, (d, area, vals) in enumerate(task["borders"][0]):
    names = sorted(vals, key=oh.QUANTITY_ORDER.index)
    codes = [oh.quantity_code(b.model, D, n) for n in names]
    sides = 3
    if d == 0:
        sides = (1 if rank == 0 else 0) | (2 if rank == world - 1 else 0)
    body.border_set_area(ci, d, area, codes, sides=sides)
tau = ora.tau
for step in range(task["steps"]):
    for stage in range(D):
        vals = []
        for (d, area, v) in task["borders"][0]:
            if d == stage:
                vals += [v[n](step * tau) for n in sorted(v, key=oh.QUANTITY_ORDER.index)]
        body.border_apply(stage, vals)
        if stage == 0:
            # my ghost planes <- neighbour's outermost real planes (a ContactCopier between slabs)
            for side, peer in ((0, rank - 1), (1, rank + 1)):
                if 0 <= peer < world:
                    mine = torch.from_numpy(body.halo_get(side))
                    theirs = torch.empty_like(mine)
                    if rank < peer:
                        dist.send(mine, peer); dist.recv(theirs, peer)
                    else:
                        dist.recv(theirs, peer); dist.send(mine, peer)
                    body.halo_put(side, theirs.numpy())
        body.stage(stage, tau)
out = torch.from_numpy(body.download(with_ghosts=False).reshape(-1))
parts = [torch.empty((nx * (r + 1) // world - nx * r // world) * sizes[1] * sizes[2] * M, dtype=torch.float64) for r in range(world)]
dist.all_gather(parts, out) if len(set(p.numel() for p in parts)) == 1 else None
if rank == 0:
    if len(set(p.numel() for p in parts)) != 1:
        raise SystemExit("uneven slabs are not gathered by this test")
    whole = torch.cat(parts).numpy().reshape(-1, M)
    ref = golden(name)["body0"]
    assert np.array_equal(whole, ref), np.abs(whole - ref).max()
    print("SLABS_OK", world)
dist.barrier()
dist.destroy_process_group()
'''


def test_two_slabs_equal_the_reference_bitwise(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SLABS_OK 2" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
