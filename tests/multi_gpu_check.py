"""Multi-GPU parity check, run under torchrun on a box with >= 2 GPUs (not collected by pytest):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29540 \
      tests/multi_gpu_check.py

Every rank owns one x-slab (host engine + C ABI + NCCL halo exchange).  Rank 0 gathers the slabs and requires
  * equality with the unmodified reference's output (tests/golden) bit for bit, and
  * equality with the same task run on ONE GPU at a size no fixture covers, bit for bit.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import numpy as np
import torch
import torch.distributed as dist

import gcm_b200
from gcm_b200 import capi
from helpers import golden
from scenarios import SCENARIOS, elastic3d_layers, ortho3d_contact


def new_nccl_id(lib, rank):
    """a fresh NCCL id for every engine (an id initialises exactly one communicator)"""
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = (capi.ctypes.c_ubyte * 128)()
        lib.check(lib.c.gcmb_comm_unique_id(capi.ctypes.cast(raw, capi.vp)))
        buf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().tolist())


def run_slabs(lib, text, rank, world, local, nccl_id=None):
    nccl_id = new_nccl_id(lib, rank)
    eng = capi.HostEngine(lib, text, device=local, slab_rank=rank, slab_count=world, nccl_id=nccl_id)
    eng.run()
    out = {}
    bid = 0
    while True:
        try:
            D, M, sizes, start = eng.body_info(bid)
        except capi.GcmError:
            break
        mine = torch.from_numpy(eng.body_pde(bid).reshape(-1)).cuda()
        counts = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(counts, torch.tensor([mine.numel()], device="cuda"))
        parts = [torch.empty(int(c.item()), dtype=torch.float64, device="cuda") for c in counts]
        # all_gather needs equal sizes: pad to the largest slab
        big = max(int(c.item()) for c in counts)
        padded = torch.zeros(big, dtype=torch.float64, device="cuda")
        padded[:mine.numel()] = mine
        gathered = [torch.empty(big, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(gathered, padded)
        out[bid] = torch.cat([g[:int(c.item())] for g, c in zip(gathered, counts)]).cpu().numpy().reshape(-1, M)
        bid += 1
    t, v = eng.seismogram()
    eng.close()
    return out, (t, v)


def check_fixtures(lib, rank, world, local, verbose=True):
    """decomposed runs against the reference fixtures; returns {"bitwise": bool, "cases": [...]} (rank 0 decides)"""
    cases, ok = [], True
    for name in ("elastic3d_layers", "elastic3d_ortho", "acoustic3d_free", "ortho3d_contact", "maxwell3d", "ortho3d_rotated_plies",
                 "elastic3d_layers_courant1", "ortho3d_contact_courant1", "elastic3d_contact_z"):
        nx = int(SCENARIOS[name].split("sizes")[1].split()[0])
        bs = int(SCENARIOS[name].split("border_size")[1].split()[0])
        if nx // world < bs:  # a slab must hold at least border_size planes (CubicGrid.hpp:186-199)
            continue
        got, seis = run_slabs(lib, SCENARIOS[name], rank, world, local)
        if rank == 0:
            g = golden(name)
            same = all(np.array_equal(arr, g["body%d" % bid]) for bid, arr in got.items())
            if "detector" in g.files:
                same = same and bool(np.allclose(seis[1], g["detector"][:, 1], rtol=5e-6, atol=1e-30))
            ok = ok and same
            cases.append({"task": name, "gpus": world, "bitwise_equal_to_reference": bool(same)})
            if verbose:
                print("multi-gpu == reference (bitwise):", name, "on", world, "GPUs", same, flush=True)
    # a contact normal to x cannot be decomposed along x: the engine must refuse it, not drop it silently
    text = SCENARIOS["ortho3d_contact"].replace("start 0 8 0", "start 16 0 0").replace("sizes 16 8 16", "sizes 16 16 16")
    refused = False
    try:
        eng = capi.HostEngine(lib, text, device=local, slab_rank=rank, slab_count=world, nccl_id=new_nccl_id(lib, rank))
        eng.close()
    except capi.GcmError as e:
        refused = e.code == -1
    if rank == 0:
        cases.append({"task": "two bodies in contact across x", "gpus": world, "refused_as_unsupported": refused})
        ok = ok and refused
    return {"bitwise": bool(ok), "cases": cases}


def main():
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    lib = gcm_b200.library()
    nccl_id = None
    os.chdir("/tmp")

    # (1) against the reference fixtures
    res = check_fixtures(lib, rank, world, local)
    if rank == 0:
        assert res["bitwise"], res

    # (2) against one GPU at a larger size (uneven slabs when world does not divide 100)
    text = elastic3d_layers(n=100, steps=12)
    got, seis = run_slabs(lib, text, rank, world, local, nccl_id)
    if rank == 0:
        one = capi.HostEngine(lib, text, device=local)
        one.run()
        ref = one.body_pde(0)
        t1, v1 = one.seismogram()
        one.close()
        assert np.array_equal(got[0], ref), np.abs(got[0] - ref).max()
        assert np.allclose(seis[1], v1, rtol=1e-6)
        print("multi-gpu == single gpu (bitwise): layered elastic 100^3, 12 steps, %d GPUs" % world, flush=True)
        print("MULTI_GPU_OK", world, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
