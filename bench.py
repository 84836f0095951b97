#!/usr/bin/env python
"""Benchmark of the gcm_b200 hot path: GCM node-updates/s on the 3-D isotropic elastic layered medium of
BASELINE.json (configs[2]: 1024^3 per GPU, border size 2, free surface on top, surface seismogram).

  python bench.py --gpus N --steps K --warmup W            our CUDA engine (torchrun for N > 1)
  python bench.py --impl reference ...                     the reference's own CPU engine (oracle/_ref)

One "step" = one full time step (all three splitting stages + border fill [+ halo exchange]) of every node.
One JSON line is printed by rank 0; see DESIGN.md "Measurement" for how every field is obtained.  Besides the headline the
line carries, each measured through the same engine loop: `config4` (BASELINE configs[3], at every N), and at N = 1 `config2`
(configs[1]), `fp32`, `fma`, `courant1` / `courant09_same_size`, `small_grids`, `simplex` (new random basis every step, the reference's
default) / `simplex_fixed_basis`, `rotated_orthotropic`, `e2e_host_state_every_step`; at N > 1 `parity_check`: fixtures of the
unmodified reference reproduced bit for bit by the decomposed engine in this very run.
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
basis {basis}
border_condition infinite fixed_force const 0 const 0 const 0
initial quantity PRESSURE 1 sphere {r!r} {sx!r} {sx!r} {sz!r}
"""


def simplex_task(nx, ny, nz, h, steps=1000000, basis="identity"):
    """a 4:4:1 plate of tetrahedra with an inner cavity (the geometry of
    meshes/layers_with_fracture.off, meshed by our box mesher), isotropic elastic, fixed identity calculation
    basis, zero fixed force on every border, pressure-sphere source"""
    lx, lz = nx * h, nz * h
    return SIMPLEX_TASK.format(steps=steps, nx=nx, ny=ny, nz=nz, h=h, c0=0.4 * lx, c2=0.6 * lx, c1=0.35 * lz, c3=0.65 * lz,
                               r=0.08 * lx, sx=0.3 * lx, sz=0.5 * lz, basis="1 0 0 0 1 0 0 0 1" if basis == "identity" else "random 1")


def simplex_section(lib, device, steps, warmup, with_cpu, basis="random"):
    """the simplex (tetrahedral) path, SURVEY.md §8 rows a13-a21: vertex-updates/s through simplex::Engine.
    basis = "random": a new random calculation basis every time step, the reference's default
    (engine/simplex/Engine.hpp:190-206) -- every characteristic foot is located afresh each step;
    basis = "identity": a fixed basis, where the located feet repeat and are cached from the third step on."""
    import numpy as np
    from gcm_b200 import capi
    nx, ny, nz = (int(x) for x in os.environ.get("GCMB_BENCH_SIMPLEX_CUBES", "160,160,40").split(","))
    h = 0.16 / nx
    eng = capi.SimplexHostEngine(lib, simplex_task(nx, ny, nz, h, basis=basis), device=device)
    info = eng.simplex_body_info(0)
    ctxh = eng.context_handle()
    eng.advance(warmup)
    lib.check(lib.c.gcmb_sync(ctxh))
    launches0 = lib.c.gcmb_launch_count(ctxh)
    lib.check(lib.c.gcmb_profile_enable(ctxh, 1))
    lib.check(lib.c.gcmb_timer_start(ctxh))
    t0 = time.perf_counter()
    eng.advance(steps)
    ms = capi.ctypes.c_float()
    lib.check(lib.c.gcmb_timer_stop(ctxh, capi.ctypes.byref(ms)))
    state = eng.simplex_pde(0)
    wall = time.perf_counter() - t0
    prof_ms = np.zeros(8)
    prof_n = np.zeros(8, dtype=np.int64)
    lib.check(lib.c.gcmb_profile_get(ctxh, 8, capi.dp(prof_ms), prof_n.ctypes.data_as(capi.c_ll_p)))
    lib.check(lib.c.gcmb_profile_enable(ctxh, 0))
    assert eng.errors() == 0 and np.isfinite(state).all() and np.abs(state).max() > 0
    out = {"metric": "simplex GCM vertex-updates/s (3-D isotropic elastic tetrahedra, SURVEY.md §8d C5)",
           "value": info["n_local"] * steps / (ms.value * 1e-3), "unit": "vertex-updates/s", "ms_per_step": ms.value / steps,
           "e2e": {"value": info["n_local"] * steps / wall, "unit": "vertex-updates/s", "d2h_bytes_per_step": info["n_local"] * 72 // steps,
                   "what": "simplex::Engine::run loop incl. border functors on the host and the final state read-back"},
           "gpu_launches": int(lib.c.gcmb_launch_count(ctxh) - launches0),
           "per_class_ms": dict(zip(("riemann_transforms", "gradient", "border_nodes", "border_correct", "contact", "inner_nodes", "setup", "ode"),
                                    (float(x) / steps for x in prof_ms))),
           "config": {"workload": "%dx%dx%d cubes of edge %g cut into 6 tetrahedra each, jitter 0.3, inner cavity: %d vertices; "
                                  "%s, fixed zero force on all borders, Courant 0.7" % (
                                      nx, ny, nz, h, info["n_local"], "a new random calculation basis every step (the reference's default: no foot can be cached)"
                                      if basis == "random" else "fixed identity calculation basis (located feet are cached from the third step on)")},
           "parity": "bit-identical to the unmodified reference simplex engine built against a CGAL stand-in (tests/golden/simplex_*.npz)"}
    eng.close()
    if with_cpu:
        # CPU baseline on a bounded sample of the same task: the UNMODIFIED reference simplex engine
        # (oracle/_ref/gcm_ref_simplex, built against a CGAL stand-in) when it was built, else the C port
        sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
        cpu_steps = 10
        text = simplex_task(24, 24, 6, 0.16 / 24, steps=cpu_steps, basis="identity")
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
    # the same with the FMA-contracted kernels (gcmb_set_fma): this stage is fp64-issue-bound, where contraction can pay
    fma = None
    try:
        eng = capi.HostEngine(lib, rotated_bench.task(n), device=device, fma=True)
        ctxh = eng.context_handle()
        eng.advance(warmup)
        lib.check(lib.c.gcmb_sync(ctxh))
        lib.check(lib.c.gcmb_timer_start(ctxh))
        eng.advance(steps)
        ms2 = capi.ctypes.c_float(0)
        lib.check(lib.c.gcmb_timer_stop(ctxh, capi.ctypes.byref(ms2)))
        fma = {"value": n ** 3 * steps / (ms2.value * 1e-3), "unit": "node-updates/s", "ms_per_step": ms2.value / steps,
               "kernels": [eng.kernel_name(0, d) for d in range(3)],
               "note": "within 1e-12 of the reference, not bit-identical (tests/test_gpu_parity.py::test_fma_kernels_within_stated_tolerance)"}
        eng.close()
    except Exception as e:
        fma = {"error": "%s: %s" % (type(e).__name__, e)}
    peak, which = measured_hbm_peak()
    cpu = None
    if with_cpu:
        try:   # the unmodified reference engine (dense 9x9 products, 81 interpolations per node-stage) on a 40^3 sample
            cpu = run_reference_cpu(40, 3, text=rotated_bench.task(40, 3), what="rotated-plies")
        except Exception as e:
            cpu = {"error": "%s: %s" % (type(e).__name__, e)}
    return {"metric": "GCM node-updates/sec (3D rotated orthotropic elastic, two glued bodies, fp64)", "value": per_s, "cpu_baseline": cpu, "fma": fma,
            "unit": "node-updates/s", "steps": steps, "warmup": warmup, "ms_per_step": ms.value / steps, "kernels": kernels,
            "config": {"workload": "two glued bodies %dx%dx%d, carbon-fibre plies (ndi.hpp:120-131) turned +-45 degrees, border size 2, "
                                   "Courant 0.9; state larger than L2" % (n, n // 2, n)},
            "roofline": {"bound": "hbm", "achieved": per_s * 3 * BYTES_PER_NODE_STAGE / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": per_s * 3 * BYTES_PER_NODE_STAGE / 1e9 / peak, "peak_source": which,
                         "note": "dense 9x9 eigen-systems: 81 limited interpolations + two dense mat-vecs per node-stage, ~800 fp64 "
                                 "instructions against 144 B; ncu (profiles/r1_dense_k0_one_ncu_384.csv): fp64 pipe 61 % active, "
                                 "DRAM traffic == algorithmic bytes"},
            "parity": "bit-identical to the unmodified reference engine (tests/golden/elastic3d_ortho_rotated.npz, ortho3d_rotated_plies.npz)"}


def make_engine(lib, text, local, rank, world, nccl_id_fn, **kw):
    from gcm_b200 import capi
    return capi.HostEngine(lib, text, device=local, slab_rank=rank, slab_count=world,
                           nccl_id=nccl_id_fn() if world > 1 else None, **kw)


def timed_engine_loop(lib, eng, K, W, barrier):
    """W warm-up steps, then K steps of Engine::run's loop timed on the device (CUDA events on the engine's stream) and by the
    host clock; per-kernel-class device times from the library's own event pairs"""
    import numpy as np
    from gcm_b200 import capi
    ctxh = eng.context_handle()
    eng.advance(W)
    barrier(ctxh)
    lib.check(lib.c.gcmb_profile_enable(ctxh, 1))
    launches0 = lib.c.gcmb_launch_count(ctxh)
    lib.check(lib.c.gcmb_timer_start(ctxh))
    t0 = time.perf_counter()
    eng.advance(K)
    ms = capi.ctypes.c_float()
    lib.check(lib.c.gcmb_timer_stop(ctxh, capi.ctypes.byref(ms)))
    barrier(ctxh)
    wall = time.perf_counter() - t0
    prof_ms = np.zeros(8)
    prof_n = np.zeros(8, dtype=np.int64)
    lib.check(lib.c.gcmb_profile_get(ctxh, 8, capi.dp(prof_ms), prof_n.ctypes.data_as(capi.c_ll_p)))
    lib.check(lib.c.gcmb_profile_enable(ctxh, 0))
    return {"dev_ms": float(ms.value), "wall_s": wall, "launches": int(lib.c.gcmb_launch_count(ctxh) - launches0),
            "class_ms": [float(x) / K for x in prof_ms], "class_launches": [int(x) // K for x in prof_n]}


def section(lib, text, nodes_per_gpu, bytes_per_node_update, K, W, local, rank, world, nccl_id_fn, barrier, reduce_max, what, **kw):
    """one secondary configuration measured like the headline: node-updates/s of the engine's time loop, device-timed"""
    eng = make_engine(lib, text, local, rank, world, nccl_id_fn, **kw)
    r = timed_engine_loop(lib, eng, K, W, barrier)
    n_bodies = 0
    while True:
        try:
            eng.body_info(n_bodies)
            n_bodies += 1
        except Exception:
            break
    kernels = [[eng.kernel_name(b, d) for d in range(3)] for b in range(min(2, n_bodies))]
    eng.close()
    dev_ms, wall = reduce_max(r["dev_ms"], r["wall_s"])
    peak, src = measured_hbm_peak()
    per_s = nodes_per_gpu * world * K / (dev_ms * 1e-3)
    return {"workload": what, "value": per_s, "unit": "node-updates/s", "ms_per_step": dev_ms / K, "steps": K, "warmup": W,
            "e2e": {"value": nodes_per_gpu * world * K / wall, "unit": "node-updates/s"},
            "kernels": kernels, "gpu_launches": r["launches"], "bytes_per_node_update": bytes_per_node_update,
            "roofline": {"bound": "hbm", "achieved": per_s / world * bytes_per_node_update / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": per_s / world * bytes_per_node_update / 1e9 / peak, "peak_source": src},
            "per_class_ms": dict(zip(("stage_x", "stage_y", "stage_z", "border", "contact", "ode", "transfer", "seismo"), r["class_ms"]))}


def c2_text(n):
    """BASELINE config 2: 3-D acoustic, point source, PRESSURE -> 0 on all six faces"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from scenarios import acoustic3d_free
    return acoustic3d_free(n, 10 ** 6).replace("sphere 0.2 0.5 0.5 0.5", "sphere 0.05 0.5 0.5 0.5")


def c4_text(nx, ny, nz):
    """BASELINE config 4: four stacked bodies nx x ny/4 x nz along y, the carbon-fibre composite of
    launcher/ndi.hpp:120-131 alternating with titanium written as an orthotropic material (ndi.hpp:136-159), automatic
    ADHESION contacts, fixed normal velocity on a disc of the top face; nx is the GLOBAL x extent"""
    sys.path.insert(0, os.path.join(ROOT, "scripts", "gpu_runs"))
    import c4_bench
    t = c4_bench.task(nz)
    q = ny // 4
    h = 1.0 / (nz - 1)
    lx = (nx - 1) * h
    out = []
    for line in t.splitlines():
        if line.startswith("body "):
            f = line.split()
            b = int(f[1])
            line = "body %d elastic orthotropic sizes %d %d %d start 0 %d 0" % (b, nx, q, nz, b * q)
        elif line.startswith("initial "):
            line = "initial quantity PRESSURE 1 sphere 0.2 %r 0.5 0.5" % (lx / 2)
        elif line.startswith("border "):
            line = "border 3 1 sphere 0.3 %r 1.0 0.5 Vy sin 1.0 5.0" % (lx / 2)
        out.append(line)
    return "\n".join(out) + "\n"


def multi_gpu_parity(lib, rank, world, local):
    """bitwise comparison of decomposed runs with the reference's fixtures (tests/golden), inside the bench"""
    sys.path[:0] = [os.path.join(ROOT, "tests")]
    import multi_gpu_check as mg
    return mg.check_fixtures(lib, rank, world, local, verbose=False)


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
    ap.add_argument("--no-sections", action="store_true", help="headline only (kernel experiments)")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)

    # the contract is ONE line on stdout, but libraries write there too (NCCL's "NCCL version ..." banner is a plain printf):
    # file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

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
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = gcm_b200.library()

    def nccl_id_fn():
        """a fresh NCCL id for every engine (an id initialises exactly one communicator)"""
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = (capi.ctypes.c_ubyte * 128)()
            lib.check(lib.c.gcmb_comm_unique_id(capi.ctypes.cast(raw, capi.vp)))
            buf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        return bytes(buf.cpu().tolist())

    def barrier(ctxh=None):
        if ctxh is not None:
            lib.check(lib.c.gcmb_sync(ctxh))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def reduce_max(*values):
        if world == 1:
            return values
        t = torch.tensor(list(values), device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return tuple(float(x) for x in t.tolist())

    os.chdir(tempfile.mkdtemp(prefix="gcmb_bench_"))  # snapshots/ of the SliceSnapshotter land here

    # ---- multi-GPU: the decomposed engine reproduces the reference's fixtures bit for bit, checked in this very run ----
    parity = None
    if world > 1:
        parity = multi_gpu_parity(lib, rank, world, local)

    free, total = torch.cuda.mem_get_info()
    n = pick_size(free, args.size)
    if world > 1:
        t = torch.tensor([n], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        n = int(t.item())
    K, W = args.steps, max(3, args.warmup)
    nodes_per_gpu = n ** 3
    text = task_text(n * world, n, n, steps=10 ** 6, detector=True)

    t_create = time.perf_counter()
    eng = make_engine(lib, text, local, rank, world, nccl_id_fn)
    ctxh = eng.context_handle()
    barrier(ctxh)
    create_s = time.perf_counter() - t_create
    body = eng.body_handle(0)
    tau = eng.info()[2]

    # ---- warm-up through the public API -------------------------------------------------------
    eng.advance(W)
    barrier(ctxh)
    kernels = [lib.c.gcmb_cubic_stage_kernel_name(body, d).decode() for d in range(3)]

    # ---- (1) device-timed region: K steps of the hot path through the C ABI, state resident in HBM -----------------
    sampler = ClockSampler(local)
    sampler.start()
    lib.check(lib.c.gcmb_profile_enable(ctxh, 1))
    launches0 = lib.c.gcmb_launch_count(ctxh)
    zeros3 = np.zeros(3)
    fused = capi.ctypes.c_int(0)
    fused_steps = 0
    barrier(ctxh)
    lib.check(lib.c.gcmb_timer_start(ctxh))
    for _ in range(K):
        # cubic::Engine::nextTimeStep (engine/cubic/Engine.cpp:92-121) for this task: [halo] stage x; stage y; free-surface
        # ghost fill + stage z
        if world > 1:
            lib.check(lib.c.gcmb_cubic_halo_exchange(body))
        lib.check(lib.c.gcmb_cubic_stage(body, 0, tau))
        lib.check(lib.c.gcmb_cubic_stage(body, 1, tau))
        # free surface + stage z in one call: the tile kernel mirrors the ghost nodes inside its shared-memory rows
        lib.check(lib.c.gcmb_cubic_stage_with_border(body, 2, tau, 3, capi.dp(zeros3), capi.ctypes.byref(fused)))
        fused_steps += fused.value
    ms = capi.ctypes.c_float()
    lib.check(lib.c.gcmb_timer_stop(ctxh, capi.ctypes.byref(ms)))
    barrier(ctxh)
    dev_ms = float(ms.value)
    launches = lib.c.gcmb_launch_count(ctxh) - launches0
    prof_ms = np.zeros(8)
    prof_n = np.zeros(8, dtype=np.int64)
    lib.check(lib.c.gcmb_profile_get(ctxh, 8, capi.dp(prof_ms), prof_n.ctypes.data_as(capi.c_ll_p)))
    lib.check(lib.c.gcmb_profile_enable(ctxh, 0))

    # ---- (2) end to end through the reference-facing API: Engine loop with border functors evaluated on
    #          the host every stage and the surface seismogram read back (and written) every step --------
    barrier(ctxh)
    t0 = time.perf_counter()
    eng.advance(K)
    barrier(ctxh)
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag.set()
    sampler.join()
    times, values = eng.seismogram()
    assert len(values) >= K and np.all(np.isfinite(values)), "seismogram is not finite"
    # ---- (3) the whole run a user of the reference launcher sees: engine construction (grid state built on the device from
    #          the Task) + the K steps above + the final state read back to pinned host memory -------------------------------
    t0 = time.perf_counter()
    final_bytes = 0
    if world == 1:
        host = torch.empty(n * n * 9, dtype=torch.float64, pin_memory=True)   # one x-plane at a time through the box read-back
        lo = np.zeros(3, dtype=np.int32)
        ext = np.array([1, n, n], dtype=np.int32)
        lib.check(lib.c.gcmb_cubic_download_box_begin(body, capi.ip(lo), capi.ip(ext), capi.ctypes.c_void_p(host.data_ptr())))
        lib.check(lib.c.gcmb_cubic_download_box_end(body))   # untimed: creates the side stream and the staging buffer
        t0 = time.perf_counter()
        for x in range(0, n, max(1, n // 8)):      # a bounded sample of planes: the figure is extrapolated, and says so
            lo[0] = x
            lib.check(lib.c.gcmb_cubic_download_box_begin(body, capi.ip(lo), capi.ip(ext), capi.ctypes.c_void_p(host.data_ptr())))
            lib.check(lib.c.gcmb_cubic_download_box_end(body))
            final_bytes += host.numel() * 8
        assert bool(torch.isfinite(host[::4097]).all())
    sample_s = time.perf_counter() - t0
    download_s = sample_s * (n ** 3 * 72 / final_bytes) if final_bytes else None

    dev_ms, e2e_s = reduce_max(dev_ms, e2e_s)
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
                "border_fill_ms": prof_ms[3] / K, "border_fill_inside_stage_z": bool(fused_steps == K),
                "whole_step_frac": value * 3 * BYTES_PER_NODE_STAGE / world / 1e9 / peak}
    traffic_file = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(traffic_file):
        try:
            t = json.load(open(traffic_file))
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at this size, from the committed
            # `ncu --set full` capture named in traffic_source (the bench itself is never run under a profiler)
            if t.get("nodes_per_launch") == nodes_per_gpu:
                roofline["traffic"] = t["dram_bytes_per_launch"]
                roofline["traffic_source"] = t["source"]
        except Exception:
            pass

    line = None
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
            "e2e_full_run": {"construction_s": create_s, "steps_s": e2e_s, "final_download_s": download_s,
                             "value": total_nodes * K / (create_s + e2e_s + (download_s or 0.0)), "unit": "node-updates/s",
                             "what": "createEngine(task) (allocation, materials, initial state and masks built on the device) + %d "
                                     "steps + read-back of the whole final state to pinned host memory (extrapolated from every "
                                     "%d-th x-plane); a run of this length is dominated by set-up, the reference's is not" % (K, max(1, n // 8))},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "device_bytes": int(lib.c.gcmb_device_bytes(ctxh)),
        }
        if parity is not None:
            line["parity_check"] = parity
        if not args.no_cpu_baseline:
            base = cpu_baseline(args.ref_size, 5)
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
    eng.close()
    if rank == 0 and world == 1 and not args.no_host_roundtrip and not args.no_sections:
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

    # ---- secondary configurations (never the headline), every one through the same engine loop ----------------------------
    def run_section(name, fn):
        try:
            out = fn()
        except Exception as e:  # the headline line must survive a failure here
            out = {"error": "%s: %s" % (type(e).__name__, e)}
        if rank == 0:
            line[name] = out

    common = dict(local=local, rank=rank, world=world, nccl_id_fn=nccl_id_fn, barrier=barrier, reduce_max=reduce_max)
    if not args.no_sections:
        # BASELINE config 4 (configs[3]): four glued orthotropic bodies per GPU, x-slabs, all four halo exchanges in one group
        n4 = n
        run_section("config4", lambda: section(
            lib, c4_text(n4 * world, n4, n4), n4 ** 3, 432, 5, 3,
            what="BASELINE config 4: four glued bodies %dx%dx%d per GPU (x-slabs over %d GPUs), composite / titanium-as-orthotropic, "
                 "ADHESION contacts, velocity disc, fp64, border size 2" % (n4, n4 // 4, n4, world), **common))
    if world == 1 and not args.no_sections:
        run_section("config2", lambda: section(
            lib, c2_text(512), 512 ** 3, 192, 10, 3,
            what="BASELINE config 2: 3-D acoustic 512^3, PRESSURE -> 0 on six faces, fp64, border size 2 (2.1 GB per layer: larger than L2)", **common))
        run_section("fp32", lambda: dict(section(
            lib, task_text(n, n, n, steps=10 ** 6, detector=True), n ** 3, 216, 10, 3,
            what="the headline task in fp32 (gcmb_create(..., 4); reference: LIBGCM_DOUBLE_PRECISION off, Types.hpp:8-14): "
                 "216 B per node-update", real_bytes=4, **common),
            parity="max|g-r|/max|r| against the fp64 run after 1 / 10 / 100 steps: tests/test_gpu_parity.py::test_fp32_variant_bound_over_steps"))
        run_section("fma", lambda: dict(section(
            lib, task_text(n, n, n, steps=10 ** 6, detector=True), n ** 3, 432, 5, 3,
            what="the headline task with the FMA-contracted fp64 kernels (gcmb_set_fma): <= 1e-12 of the reference, not bit-identical",
            fma=True, **common)))
        m1 = min(n, 512)
        run_section("courant1", lambda: section(
            lib, task_text(m1, m1, m1, steps=10 ** 6, detector=True).replace("courant 0.9", "courant 1.0"), m1 ** 3, 432, 10, 3,
            what="the headline medium at %d^3 with the reference launcher's Courant number 1 (src/launcher/main.cpp:82,197,...): feet in "
                 "the second cell, specialised kernels with run-time foot cells" % m1, **common))
        run_section("courant09_same_size", lambda: section(
            lib, task_text(m1, m1, m1, steps=10 ** 6, detector=True), m1 ** 3, 432, 10, 3,
            what="the headline medium at %d^3, Courant 0.9 (the comparison for courant1)" % m1, **common))
        # sizes the reference's own tasks have (its tests and launcher run 10^4 - 10^7 nodes): the launches are shaped for
        # about four waves of blocks whatever the size (stage_inst.cu launch_march / launch_ztile)
        def small_grids():
            out = {}
            for m in (64, 128, 256):
                r = section(lib, task_text(m, m, m, steps=10 ** 6, detector=True), m ** 3, 432, 100, 10,
                            what="the headline medium at %d^3" % m, **common)
                out["%d^3" % m] = {"value": r["value"], "unit": r["unit"], "ms_per_step": r["ms_per_step"],
                                   "of_hbm_roofline": r["roofline"]["frac"], "e2e": r["e2e"]["value"]}
            return out
        run_section("small_grids", small_grids)
    if rank == 0:
        if world == 1 and not args.no_simplex and not args.no_sections:
            # secondary measurement (never the headline): the tetrahedral path of SURVEY.md §8 a13-a21
            for key, basis in (("simplex", "random"), ("simplex_fixed_basis", "identity")):
                try:
                    line[key] = simplex_section(lib, local, 20, 3, (not args.no_cpu_baseline) and basis == "random", basis=basis)
                except Exception as e:  # the headline line must survive a failure here
                    line[key] = {"error": "%s: %s" % (type(e).__name__, e)}
        if world == 1 and not args.no_rotated and not args.no_sections:
            # secondary measurement (never the headline): dense eigen-systems of rotated orthotropic materials
            try:
                line["rotated_orthotropic"] = rotated_section(lib, local, 5, 3, not args.no_cpu_baseline)
            except Exception as e:
                line["rotated_orthotropic"] = {"error": "%s: %s" % (type(e).__name__, e)}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
