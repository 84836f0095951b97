#!/usr/bin/env python
"""Benchmark of the gcm_b200 hot path: GCM node-updates/s on the 3-D isotropic elastic layered medium of
BASELINE.json (configs[2]: 1024^3 per GPU, border size 2, free surface on top, surface seismogram).

  python bench.py --gpus N --steps K --warmup W            our CUDA engine (torchrun for N > 1)
  python bench.py --impl reference ...                     the reference's own CPU engine (oracle/_ref)

One "step" = one full time step (all three splitting stages + border fill [+ halo exchange]) of every node.
One JSON line is printed by rank 0; see DESIGN.md "Measurement" for how every field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GCM node-updates/sec (3D isotropic elastic, layered, fp64)"
BYTES_PER_NODE_STAGE = 2 * 9 * 8  # algorithmic: read M + write M doubles


def task_text(nx, ny, nz, steps, detector=True):
    """BASELINE config 3: layered isotropic medium, free surface on the top z face,
    P-wave pulse, detector disc on the top face.  nx is the GLOBAL x extent (slabs are cut by the engine)."""
    h = repr(1.0 / (nz - 1))
    lx = (nx - 1) / (nz - 1)
    t = f"""
dimensionality 3
courant 0.9
border_size 2
h {h} {h} {h}
steps {steps}
body 0 elastic isotropic sizes {nx} {ny} {nz} start 0 0 0
material default isotropic 1 2 0.8
material area box -10 0.2 -10 {lx + 10} 0.4 10 isotropic 0.5 2 0.8
material area box -10 0.4 -10 {lx + 10} 0.7 10 isotropic 2 2 0.8
material area box -10 0.7 -10 {lx + 10} 10 10 isotropic 4 2 0.8
initial wave P_FORWARD 2 PRESSURE 1 box -10 -10 0.3 {lx + 10} 10 0.6
border 0 2 infinite Sxz const 0 Syz const 0 Szz const 0
"""
    if detector:
        t += f"detector 0 Vz sphere 0.4 {lx / 2} 0.5 1.0 bench\n"
    return t


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(float(s[0])) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.samples[0][1])), "reasons": reasons,
                "samples": len(sm), "power_w_max": max(float(s[2]) for s in self.samples)}


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference_cpu(n, steps, text=None, what="layered elastic"):
    """The reference's own CPU engine (oracle/_ref/gcm_ref, unmodified sources) on a bounded sample."""
    exe = os.path.join(ROOT, "oracle", "_ref", "gcm_ref")
    kind = "reference"
    if not os.path.exists(exe):
        return None
    with tempfile.TemporaryDirectory() as tmp:
        tf = os.path.join(tmp, "task.txt")
        with open(tf, "w") as f:
            f.write(text if text is not None else task_text(n, n, n, steps, detector=False))
        subprocess.run([exe, tf, os.path.join(tmp, "out")], check=True, cwd=tmp)
        meta = dict(line.split()[:2] for line in open(os.path.join(tmp, "out.meta")) if not line.startswith("body"))
    seconds = float(meta["run_seconds"])
    done = int(float(meta["steps"]))
    return {"value": n ** 3 * done / seconds, "unit": "node-updates/s", "cores": 1, "kind": kind,
            "sample": "%d^3 nodes x %d steps of the same %s task, run() wall time %.2f s" % (n, done, what, seconds),
            "seconds": seconds, "steps": done}


def run_oracle_port_cpu(n, steps):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_host as oh
    eng = oh.OracleEngine(oh.parse_task(task_text(n, n, n, steps, detector=False)))
    t0 = time.perf_counter()
    eng.run()
    seconds = time.perf_counter() - t0
    return {"value": n ** 3 * eng.steps_done / seconds, "unit": "node-updates/s", "cores": 1, "kind": "port",
            "sample": "%d^3 nodes x %d steps of the same layered elastic task (C restatement)" % (n, eng.steps_done),
            "seconds": seconds, "steps": eng.steps_done}


def cpu_baseline(n, steps):
    r = run_reference_cpu(n, steps)
    return r if r is not None else run_oracle_port_cpu(n, steps)


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = args.ref_size
    t0 = time.perf_counter()
    base = cpu_baseline(n, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "node-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * base["seconds"] / max(1, base["steps"]), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "3D isotropic elastic layered medium, border size 2, free surface + P-wave pulse; "
                               "reference CPU engine on a %d^3 sample of the 1024^3-per-GPU task (1 thread: the "
                               "reference's stage loop is sequential)" % n,
                   "note": "CPU run has no warm-up phase; all steps timed"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "node-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))
    return 0


def pick_size(free_bytes, want):
    """largest cube edge (multiple of 128) whose two fp64 time layers + material ids fit"""
    for n in [want] + [s for s in (1024, 896, 768, 640, 512, 384, 256, 128) if s < want]:
        pitch = (16 + n + 2 + 15) // 16 * 16
        comp = (n + 4) * (n + 4) * pitch
        need = comp * (2 * 9 * 8 + 1) + (1 << 30)
        if need < free_bytes:
            return n
    return 64


SIMPLEX_TASK = """grid simplex
dimensionality 3
courant 0.7
steps {steps}
simplex_box {nx} {ny} {nz} 0 0 0 {h!r} jitter 0.3 seed 1
cavity box {c0!r} {c0!r} {c1!r} {c2!r} {c2!r} {c3!r}
body 0 elastic isotropic
material body 0 isotropic 7800 1.2e11 8.0e10
basis 1 0 0 0 1 0 0 0 1
border_condition infinite fixed_force const 0 const 0 const 0
initial quantity PRESSURE 1 sphere {r!r} {sx!r} {sx!r} {sz!r}
"""


def simplex_task(nx, ny, nz, h, steps=1000000):
    """a 4:4:1 plate of tetrahedra with an inner cavity (the geometry of
    meshes/layers_with_fracture.off, meshed by our box mesher), isotropic elastic, fixed identity calculation
    basis, zero fixed force on every border, pressure-sphere source"""
    lx, lz = nx * h, nz * h
    return SIMPLEX_TASK.format(steps=steps, nx=nx, ny=ny, nz=nz, h=h, c0=0.4 * lx, c2=0.6 * lx, c1=0.35 * lz, c3=0.65 * lz,
                               r=0.08 * lx, sx=0.3 * lx, sz=0.5 * lz)


def simplex_section(lib, device, steps, warmup, with_cpu):
    """the simplex (tetrahedral) path, SURVEY.md §8 rows a13-a21: vertex-updates/s through simplex::Engine"""
    import numpy as np
    from gcm_b200 import capi
    nx, ny, nz = (int(x) for x in os.environ.get("GCMB_BENCH_SIMPLEX_CUBES", "160,160,40").split(","))
    h = 0.16 / nx
    eng = capi.SimplexHostEngine(lib, simplex_task(nx, ny, nz, h), device=device)
    info = eng.simplex_body_info(0)
    ctxh = eng.context_handle()
    eng.advance(warmup)
    lib.check(lib.c.gcmb_sync(ctxh))
    launches0 = lib.c.gcmb_launch_count(ctxh)
    lib.check(lib.c.gcmb_timer_start(ctxh))
    t0 = time.perf_counter()
    eng.advance(steps)
    ms = capi.ctypes.c_float()
    lib.check(lib.c.gcmb_timer_stop(ctxh, capi.ctypes.byref(ms)))
    state = eng.simplex_pde(0)
    wall = time.perf_counter() - t0
    assert eng.errors() == 0 and np.isfinite(state).all() and np.abs(state).max() > 0
    out = {"metric": "simplex GCM vertex-updates/s (3-D isotropic elastic tetrahedra, SURVEY.md §8d C5)",
           "value": info["n_local"] * steps / (ms.value * 1e-3), "unit": "vertex-updates/s", "ms_per_step": ms.value / steps,
           "e2e": {"value": info["n_local"] * steps / wall, "unit": "vertex-updates/s", "d2h_bytes_per_step": info["n_local"] * 72 // steps,
                   "what": "simplex::Engine::run loop incl. border functors on the host and the final state read-back"},
           "gpu_launches": int(lib.c.gcmb_launch_count(ctxh) - launches0),
           "config": {"workload": "%dx%dx%d cubes of edge %g cut into 6 tetrahedra each, jitter 0.3, inner cavity: %d vertices; "
                                  "identity calculation basis, fixed zero force on all borders, Courant 0.7" % (nx, ny, nz, h, info["n_local"])},
           "parity": "bit-identical to the unmodified reference simplex engine built against a CGAL stand-in (tests/golden/simplex_*.npz)"}
    eng.close()
    if with_cpu:
        # CPU baseline on a bounded sample of the same task: the UNMODIFIED reference simplex engine
        # (oracle/_ref/gcm_ref_simplex, built against a CGAL stand-in) when it was built, else the C port
        sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
        cpu_steps = 10
        text = simplex_task(24, 24, 6, 0.16 / 24, steps=cpu_steps)
        small = capi.SimplexHostEngine(lib, text, device=device)
        i2 = small.simplex_body_info(0)
        tri = small.triangulation()
        exe = os.path.join(ROOT, "oracle", "_ref", "gcm_ref_simplex")
        if os.path.exists(exe):
            from simplex_helpers import run_reference_simplex
            ref, meta = run_reference_simplex(text, tri, tempfile.mkdtemp(prefix="gcmb_sx_"), [0], 9)
            small.run()
            same = bool(np.array_equal(small.simplex_pde(0), ref[0][1]) and small.info()[2] == meta["tau"])
            out["cpu_baseline"] = {"value": i2["n_local"] * cpu_steps / meta["run_seconds"], "unit": "vertex-updates/s", "cores": 1,
                                   "kind": "reference",
                                   "sample": "%d steps of the same task at 24x24x6 cubes (%d vertices) by the reference's "
                                             "simplex::Engine::run(), %.2f s" % (cpu_steps, i2["n_local"], meta["run_seconds"]),
                                   "gpu_equals_reference_bitwise": same}
        else:
            import simplex_host
            U, U1, L = small.simplex_matrices(0)
            nodes, normals = small.border_nodes(0, 0)
            pde0 = small.simplex_pde(0)
            tau = small.info()[2]
            ref, spent = simplex_host.run_single_body(tri, 0, 0, U, U1, L, np.eye(3), nodes, normals, np.zeros(len(nodes), dtype=np.int32),
                                                      np.zeros(1, dtype=np.int32), lambda t: np.zeros((1, 3)), pde0, tau, cpu_steps)
            small.advance(cpu_steps)
            same = bool(np.array_equal(small.simplex_pde(0), ref))
            out["cpu_baseline"] = {"value": i2["n_local"] * cpu_steps / spent, "unit": "vertex-updates/s", "cores": 1, "kind": "port",
                                   "sample": "%d steps of the same task at 24x24x6 cubes (%d vertices) through oracle/simplex_oracle.c" % (cpu_steps, i2["n_local"]),
                                   "gpu_equals_port_bitwise": same}
        small.close()
    return out


def rotated_section(lib, device, steps, warmup, with_cpu=True):
    """rotated orthotropic plies (BASELINE config 4 with the plies turned about the stacking axis):
    node-updates/s of the dense-eigen-system stage kernel through Engine::run's loop, state resident in HBM"""
    import torch
    from gcm_b200 import capi
    sys.path.insert(0, os.path.join(ROOT, "scripts", "gpu_runs"))
    import rotated_bench
    n = 512
    eng = capi.HostEngine(lib, rotated_bench.task(n), device=device)
    ctxh = eng.context_handle()
    eng.advance(warmup)
    lib.check(lib.c.gcmb_sync(ctxh))
    torch.cuda.synchronize()
    lib.check(lib.c.gcmb_timer_start(ctxh))
    eng.advance(steps)
    ms = capi.ctypes.c_float(0)
    lib.check(lib.c.gcmb_timer_stop(ctxh, capi.ctypes.byref(ms)))
    kernels = [eng.kernel_name(0, d) for d in range(3)]
    eng.close()
    per_s = n ** 3 * steps / (ms.value * 1e-3)
    peak, which = measured_hbm_peak()
    cpu = None
    if with_cpu:
        try:   # the unmodified reference engine (dense 9x9 products, 81 interpolations per node-stage) on a 40^3 sample
            cpu = run_reference_cpu(40, 3, text=rotated_bench.task(40, 3), what="rotated-plies")
        except Exception as e:
            cpu = {"error": "%s: %s" % (type(e).__name__, e)}
    return {"metric": "GCM node-updates/sec (3D rotated orthotropic elastic, two glued bodies, fp64)", "value": per_s, "cpu_baseline": cpu,
            "unit": "node-updates/s", "steps": steps, "warmup": warmup, "ms_per_step": ms.value / steps, "kernels": kernels,
            "config": {"workload": "two glued bodies %dx%dx%d, carbon-fibre plies (ndi.hpp:120-131) turned +-45 degrees, border size 2, "
                                   "Courant 0.9; state larger than L2" % (n, n // 2, n)},
            "roofline": {"bound": "hbm", "achieved": per_s * 3 * BYTES_PER_NODE_STAGE / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": per_s * 3 * BYTES_PER_NODE_STAGE / 1e9 / peak, "peak_source": which,
                         "note": "dense 9x9 eigen-systems: 81 limited interpolations + two dense mat-vecs per node-stage, ~800 fp64 "
                                 "instructions against 144 B; ncu (profiles/r1_dense_k0_one_ncu_384.csv): fp64 pipe 61 % active, "
                                 "DRAM traffic == algorithmic bytes"},
            "parity": "bit-identical to the unmodified reference engine (tests/golden/elastic3d_ortho_rotated.npz, ortho3d_rotated_plies.npz)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gcm_b200")
    ap.add_argument("--size", type=int, default=1024, help="cube edge per GPU")
    ap.add_argument("--ref-size", type=int, default=64, help="cube edge of the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-simplex", action="store_true", help="skip the secondary simplex-path measurement")
    ap.add_argument("--no-rotated", action="store_true", help="skip the secondary rotated-orthotropic measurement")
    ap.add_argument("--no-host-roundtrip", action="store_true", help="skip the host-resident-state measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)

    import numpy as np
    import torch
    import gcm_b200
    from gcm_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: gcm_b200 has no CPU path")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    nccl_id = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = gcm_b200.library()
    if world > 1:
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = (capi.ctypes.c_ubyte * 128)()
            lib.check(lib.c.gcmb_comm_unique_id(capi.ctypes.cast(raw, capi.vp)))
            buf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        nccl_id = bytes(buf.cpu().tolist())

    free, total = torch.cuda.mem_get_info()
    n = pick_size(free, args.size)
    if world > 1:
        t = torch.tensor([n], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        n = int(t.item())
    K, W = args.steps, max(3, args.warmup)
    nodes_per_gpu = n ** 3
    text = task_text(n * world, n, n, steps=10 ** 6, detector=True)

    os.chdir(tempfile.mkdtemp(prefix="gcmb_bench_"))  # snapshots/ of the SliceSnapshotter land here
    eng = capi.HostEngine(lib, text, device=local, slab_rank=rank, slab_count=world, nccl_id=nccl_id)
    ctxh = eng.context_handle()
    body = eng.body_handle(0)
    kernels = [lib.c.gcmb_cubic_stage_kernel_name(body, d).decode() for d in range(3)]
    tau = eng.info()[2]

    def barrier():
        lib.check(lib.c.gcmb_sync(ctxh))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- warm-up through the public API -------------------------------------------------------
    eng.advance(W)
    barrier()

    # ---- (1) device-timed region: K steps of the hot path, state resident in HBM -----------------
    sampler = ClockSampler(local)
    sampler.start()
    lib.check(lib.c.gcmb_profile_enable(ctxh, 1))
    launches0 = lib.c.gcmb_launch_count(ctxh)
    zeros3 = np.zeros(3)
    barrier()
    lib.check(lib.c.gcmb_timer_start(ctxh))
    for _ in range(K):
        for d in range(3):
            if d == 2:
                lib.check(lib.c.gcmb_cubic_border_apply(body, 2, 3, capi.dp(zeros3)))
            if d == 0 and world > 1:
                lib.check(lib.c.gcmb_cubic_halo_exchange(body))
            lib.check(lib.c.gcmb_cubic_stage(body, d, tau))
    ms = capi.ctypes.c_float()
    lib.check(lib.c.gcmb_timer_stop(ctxh, capi.ctypes.byref(ms)))
    barrier()
    dev_ms = float(ms.value)
    launches = lib.c.gcmb_launch_count(ctxh) - launches0
    prof_ms = np.zeros(8)
    prof_n = np.zeros(8, dtype=np.int64)
    lib.check(lib.c.gcmb_profile_get(ctxh, 8, capi.dp(prof_ms), prof_n.ctypes.data_as(capi.c_ll_p)))
    lib.check(lib.c.gcmb_profile_enable(ctxh, 0))

    # ---- (2) end to end through the reference-facing API: Engine loop with border functors evaluated on
    #          the host every stage and the surface seismogram read back (and written) every step --------
    barrier()
    t0 = time.perf_counter()
    eng.advance(K)
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag.set()
    sampler.join()
    times, values = eng.seismogram()
    assert len(values) >= K and np.all(np.isfinite(values)), "seismogram is not finite"

    if world > 1:
        t = torch.tensor([dev_ms, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(t[0].item()), float(t[1].item())

    total_nodes = nodes_per_gpu * world
    value = total_nodes * K / (dev_ms * 1e-3)
    e2e = total_nodes * K / e2e_s

    # ---- roofline of the dominant kernel (the slowest of the three stage kernels) ------------------
    peak, peak_src = measured_hbm_peak()
    stage_ms = [prof_ms[a] / K for a in range(3)]   # per stage and step (the x stage is three launches when the halo exchange overlaps it)
    dom = int(np.argmax(stage_ms))
    bytes_per_launch = BYTES_PER_NODE_STAGE * nodes_per_gpu
    achieved = bytes_per_launch / (stage_ms[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": kernels[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "per_stage_ms": {kernels[a]: stage_ms[a] for a in range(3)},
                "per_stage_gbs": {kernels[a]: bytes_per_launch / (stage_ms[a] * 1e-3) / 1e9 for a in range(3)},
                "whole_step_frac": value * 3 * BYTES_PER_NODE_STAGE / world / 1e9 / peak}
    traffic_file = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(traffic_file):
        try:
            t = json.load(open(traffic_file))
            # dram__bytes_read+write of one launch from the ncu --set full capture (taken at 512^3), scaled
            # to this run's launch by the node count
            roofline["traffic"] = t["ratio_to_algorithmic"] * bytes_per_launch
            roofline["traffic_source"] = t["source"]
        except Exception:
            pass

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "node-updates/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "3D isotropic elastic layered medium %dx%dx%d per GPU (x-slabs), border size 2, "
                                   "4 materials in y-layers, free surface on the top z face, P-wave pulse, surface "
                                   "seismogram; Courant 0.9" % (n, n, n),
                       "nodes_per_gpu": nodes_per_gpu, "global_grid": [n * world, n, n],
                       "cache": "inputs (%.1f GB per time layer) are far larger than L2, no flush needed"
                                % (nodes_per_gpu * 72 / 1e9),
                       "parallelism": "x-slab decomposition, NCCL halo exchange" if world > 1 else "single GPU",
                       "stage_kernels": kernels},
            "roofline": roofline,
            "e2e": {"value": e2e, "unit": "node-updates/s", "ms_per_step": 1e3 * e2e_s / K,
                    "h2d_bytes_per_step": 3 * 8, "d2h_bytes_per_step": 16 + 8 * n,
                    "what": "Engine time loop of the host layer (reference AbstractEngine::run body): border functors "
                            "evaluated on the host each stage and passed to the kernels, detector sum + z-axis line "
                            "read back and written to the SliceSnapshotter text files every step; the grid state is "
                            "created on the device from the Task's analytic areas, like the reference builds it from "
                            "the Task (no state array crosses PCIe)"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "device_bytes": int(lib.c.gcmb_device_bytes(ctxh)),
        }
        if not args.no_cpu_baseline:
            base = cpu_baseline(args.ref_size, 5)
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
    eng.close()
    if rank == 0 and world == 1 and not args.no_host_roundtrip:
        # ---- (3) the pessimistic bound of a drop-in that keeps the reference's HOST-resident mesh: the whole state
        #          (reference AoS layout with ghost nodes) goes up from pinned host memory before EVERY step and comes
        #          back after it.  Measured on a 512^3 body (9.8 GB each way) so that the pinned buffer fits any host.
        try:
            m = min(n, 512)
            small = capi.HostEngine(lib, task_text(m, m, m, steps=10 ** 6, detector=False), device=local)
            sbody, sctx = small.body_handle(0), small.context_handle()
            stau = small.info()[2]
            count = (m + 4) ** 3 * 9
            host = torch.empty(count, dtype=torch.float64, pin_memory=True)
            hp = capi.ctypes.c_void_p(host.data_ptr())
            lib.check(lib.c.gcmb_cubic_download_state(sbody, hp, 1))
            lib.check(lib.c.gcmb_sync(sctx))
            reps = 2
            t0 = time.perf_counter()
            for _ in range(reps):
                lib.check(lib.c.gcmb_cubic_upload_state(sbody, hp, 1))
                for d in range(3):
                    if d == 2:
                        lib.check(lib.c.gcmb_cubic_border_apply(sbody, 2, 3, capi.dp(zeros3)))
                    lib.check(lib.c.gcmb_cubic_stage(sbody, d, stau))
                lib.check(lib.c.gcmb_cubic_download_state(sbody, hp, 1))
            lib.check(lib.c.gcmb_sync(sctx))
            dt = (time.perf_counter() - t0) / reps
            assert bool(torch.isfinite(host[:: 4097]).all())
            line["e2e_host_state_every_step"] = {
                "value": m ** 3 / dt, "unit": "node-updates/s", "ms_per_step": 1e3 * dt,
                "h2d_bytes_per_step": count * 8, "d2h_bytes_per_step": count * 8,
                "what": "gcmb_cubic_upload_state from pinned host memory + one time step + gcmb_cubic_download_state, "
                        "%d^3 body: what a binding that leaves the mesh in host memory (reference DefaultMesh) would "
                        "pay per step; PCIe-bound, the reason INTEGRATION.md keeps the mesh resident in HBM" % m}
            small.close()
            del host
        except Exception as e:
            line["e2e_host_state_every_step"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank == 0:
        if world == 1 and not args.no_simplex:
            # secondary measurement (never the headline): the tetrahedral path of SURVEY.md §8 a13-a21
            try:
                line["simplex"] = simplex_section(lib, local, 20, 3, not args.no_cpu_baseline)
            except Exception as e:  # the headline line must survive a failure here
                line["simplex"] = {"error": "%s: %s" % (type(e).__name__, e)}
        if world == 1 and not args.no_rotated:
            # secondary measurement (never the headline): dense eigen-systems of rotated orthotropic materials
            try:
                line["rotated_orthotropic"] = rotated_section(lib, local, 5, 3, not args.no_cpu_baseline)
            except Exception as e:
                line["rotated_orthotropic"] = {"error": "%s: %s" % (type(e).__name__, e)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
