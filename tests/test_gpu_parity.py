"""Parity of the CUDA path (through the C ABI, on a real GPU) against
  * outputs of the UNMODIFIED reference committed under tests/golden (bit for bit),
  * the CPU oracle (oracle/cubic_oracle.c, itself pinned bit-for-bit to the reference) on seeded random inputs,
  * size-independent properties at sizes the oracle cannot reach.
Tolerance stated by north_star: <= 1e-12 relative in fp64.  We require bitwise equality wherever the
reference result is available and assert the 1e-12 bound as well."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

import gcm_b200
from gcm_b200 import capi
from helpers import compare_with_golden, run_engine
from scenarios import SCENARIOS, acoustic3d_free, elastic3d_iso, elastic3d_layers

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    lib = gcm_b200.library()
    assert lib.cuda_path.endswith("gcm_b200/libgcm_b200.so")  # the CUDA library, nothing else
    return lib


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_cuda_engine_matches_reference_bitwise(lib, name):
    eng, worst = compare_with_golden(lib, name, SCENARIOS[name])
    assert worst == 0.0
    eng.close()


VARIANT_ENVS = [{"GCMB_FORCE_DENSE": "1"}, {"GCMB_FORCE_DENSE": "1", "GCMB_DENSE_LITERAL": "1"}, {"GCMB_MARCH_SEG": "5"}, {"GCMB_MARCH_SEG": "0"},
                {"GCMB_ZTILE_ROWS": "7"}, {"GCMB_STAGE_IMPL": "2"}, {"GCMB_STAGE_IMPL": "3"}, {"GCMB_STAGE_IMPL": "3", "GCMB_MARCH_SEG": "5", "GCMB_ZTILE_ROWS": "3"},
                {"GCMB_FUSED_BORDER": "1"}, {"GCMB_FUSED_BORDER": "1", "GCMB_STAGE_IMPL": "3"}]


@pytest.mark.parametrize("env", VARIANT_ENVS)
def test_kernel_variants_match_reference(env):
    """dense kernels, cp.async (LDGSTS) and bulk-copy (TMA) pipelines, odd marching segments and row counts, fused and
    separate ghost fill: every variant reproduces the reference bits -- on the fixtures and, for sizes with full warps
    and several z chunks, on random states against the oracle."""
    code = ("import sys; sys.path[:0] = [%r, %r, %r]\n"
            "import gcm_b200\n"
            "from helpers import compare_with_golden, random_stage_check, fused_border_check\n"
            "from scenarios import SCENARIOS\n"
            "lib = gcm_b200.library()\n"
            "for n in ('elastic3d_layers', 'ortho3d_contact', 'acoustic3d_free', 'elastic2d_ortho', 'acoustic2d_border1', 'maxwell3d', 'ortho3d_rotated_plies',\n"
            "          'elastic3d_ortho_rotated', 'elastic3d_layers_courant1', 'acoustic3d_courant1', 'elastic2d_courant1', 'ortho3d_contact_courant1',\n"
            "          'elastic3d_ortho_rotated_courant1', 'elastic3d_layers_bs3_courant25', 'elastic2d_bs3_courant15', 'acoustic1d'):\n"
            "    compare_with_golden(lib, n, SCENARIOS[n])[0].close()\n"
            "random_stage_check(lib, ((3, (19, 13, 37), 'elastic', 2), (3, (7, 9, 300), 'acoustic', 2), (3, (5, 40, 700), 'elastic', 2),\n"
            "                         (2, (23, 131), 'elastic', 2), (2, (70, 515), 'acoustic', 1), (3, (6, 11, 130), 'elastic', 3), (1, (1000,), 'acoustic', 2)))\n"
            "fused_border_check(lib)\n"
            % (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_reference_engine_with_gpu_backend_matches_reference_bitwise(name, tmp_path):
    """The drop-in, run: oracle/_ref/gcm_ref_gpu is the UNMODIFIED reference cubic::Engine<D> (Engine.cpp, AbstractEngine.cpp,
    models, materials, snapshotters compiled where they lie) whose factory hands out the GPU-backed mesh / GCM / border /
    contact / ODE objects of integration/GpuBackend.hpp.  Its results must equal those of the all-CPU reference build
    (tests/golden, made by oracle/_ref/gcm_ref) bit for bit: states of every body, time step, step count, seismogram."""
    import oracle_host as oh
    from helpers import golden
    exe = os.path.join(ROOT, "oracle", "_ref", "gcm_ref_gpu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/gcm_ref_gpu is built where /root/reference exists (make -C oracle ref_gpu)")
    g = golden(name)
    out = oh.run_reference(SCENARIOS[name], str(tmp_path), exe_name="gcm_ref_gpu")
    assert out["meta"]["tau"] == float(g["tau"]) and out["meta"]["time"] == float(g["time"]) and int(out["meta"]["steps"]) == int(g["steps"])
    bid = 0
    while "body%d" % bid in g.files:
        assert np.array_equal(out[bid], g["body%d" % bid]), (name, bid, np.abs(out[bid] - g["body%d" % bid]).max())
        bid += 1
    if "detector" in g.files:
        assert np.array_equal(out["detector"], g["detector"])  # both written by the reference's own SliceSnapshotter


def test_kernel_names_report_what_was_launched(lib):
    """the Courant number of the reference launcher (1) and border size 3 run the specialised kernels, and
    gcmb_cubic_stage_kernel_name says so"""
    for name, want in (("elastic3d_layers_courant1", "sparse:elastic3d_iso_%s/bs2+k0"), ("acoustic3d_courant1", "sparse:acoustic3d_%s/bs2+k0"),
                       ("elastic3d_iso_bs3", "sparse:elastic3d_iso_%s/bs3+k0"), ("elastic3d_layers", "sparse:elastic3d_iso_%s/bs2"),
                       ("elastic3d_ortho_rotated_courant1", "dense_k0_one:M9/bs2+k0")):
        eng = capi.HostEngine(lib, SCENARIOS[name])
        assert [eng.kernel_name(0, d) for d in range(3)] == ["unset"] * 3
        eng.advance(1)
        assert [eng.kernel_name(0, d) for d in range(3)] == [want % c if "%s" in want else want for c in "xyz"]
        eng.close()


def test_fused_border_fill_equals_separate_fill():
    """GCMB_FUSED_BORDER=1 (off by default: it does not pay): the marching stage writes the z ghosts itself"""
    code = ("import sys; sys.path[:0] = [%r, %r, %r]\n"
            "import gcm_b200\n"
            "from helpers import fused_border_check\n"
            "n = fused_border_check(gcm_b200.library())\n"
            "assert n >= 4, n\n" % (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")))
    for impl in ("2", "3"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GCMB_FUSED_BORDER="1", GCMB_STAGE_IMPL=impl), capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("impl", ["0", "2", "3"])
def test_border_inside_contiguous_stage_equals_separate_fill(impl):
    """gcmb_cubic_stage_with_border: the tile kernels of the contiguous axis (LDGSTS and bulk-copy pipelines, fp64 and fp32)
    mirror the ghost nodes inside their shared-memory rows; real nodes equal border fill + stage bit for bit"""
    code = ("import sys; sys.path[:0] = [%r, %r, %r]\n"
            "import gcm_b200\n"
            "from helpers import tile_border_check\n"
            "lib = gcm_b200.library()\n"
            "assert tile_border_check(lib, 8) == 14 and tile_border_check(lib, 4) == 14\n" % (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GCMB_STAGE_IMPL=impl), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


def test_separate_border_fill_still_matches_reference():
    """GCMB_ZTILE_BORDER=0: the sector-wide fill kernel of the z faces (the path bodies with a contact across z take)"""
    code = ("import sys; sys.path[:0] = [%r, %r, %r]\n"
            "import gcm_b200\n"
            "from helpers import compare_with_golden\n"
            "from scenarios import SCENARIOS\n"
            "lib = gcm_b200.library()\n"
            "for n in ('elastic3d_layers', 'acoustic3d_free', 'elastic3d_layers_courant1', 'elastic3d_layers_bs3_courant25', 'acoustic1d'):\n"
            "    compare_with_golden(lib, n, SCENARIOS[n])[0].close()\n" % (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GCMB_ZTILE_BORDER="0"), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_fma_kernels_within_stated_tolerance(lib, name):
    """gcmb_set_fma: the stage kernels compiled WITH contraction agree with the reference to north_star's 1e-12 (relative to
    the largest value of the field); bit-exactness is the default mode's property, not this one's"""
    from helpers import golden
    g = golden(name)
    eng = capi.HostEngine(lib, SCENARIOS[name], fma=True).run()
    assert eng.info()[0] == int(g["steps"])
    bid = 0
    while "body%d" % bid in g.files:
        ref, got = g["body%d" % bid], eng.body_pde(bid)
        assert np.abs(ref - got).max() <= 1e-12 * np.abs(ref).max(), (name, np.abs(ref - got).max() / np.abs(ref).max())
        bid += 1
    assert all("+fma" in eng.kernel_name(0, d) for d in range(eng.body_info(0)[0]))
    eng.close()


@pytest.mark.parametrize("maker,kw", [(elastic3d_layers, dict(n=48)), (acoustic3d_free, dict(n=48))])
def test_fp32_variant_bound_over_steps(lib, maker, kw):
    """gcmb_create(..., 4): max|g - r| / max|r| against the fp64 run (== the reference, bitwise) after 1, 10 and 100 steps
    (SURVEY.md 8d parity protocol; the reference's own fp32 tolerance is 1e-3, util/infrastructure/Types.hpp:13)."""
    f64 = capi.HostEngine(lib, maker(steps=1000, **kw))
    f32 = capi.HostEngine(lib, maker(steps=1000, **kw), real_bytes=4)
    assert f32.info()[2] == f64.info()[2]  # the same time step
    done, bounds = 0, {}
    for n in (1, 10, 100):
        f64.advance(n - done)
        f32.advance(n - done)
        done = n
        r, g = f64.body_pde(0), f32.body_pde(0)
        bounds[n] = float(np.abs(g - r).max() / np.abs(r).max())
    print("fp32 bound", maker.__name__, bounds)
    assert bounds[1] <= 2e-6 and bounds[10] <= 1e-5 and bounds[100] <= 1e-4, bounds
    assert "/f32" in f32.kernel_name(0, 0)
    f64.close()
    f32.close()


def test_fp32_random_state_single_stages_close_to_oracle(lib):
    from helpers import random_stage_check
    random_stage_check(lib, ((3, (19, 13, 37), "elastic", 2), (3, (5, 40, 300), "acoustic", 2), (2, (23, 131), "elastic", 1),
                             (3, (6, 11, 130), "elastic", 3)), real_bytes=4)


def test_async_box_download_matches_state(lib):
    ctx = capi.Context(lib)
    rng = np.random.default_rng(5)
    body = capi.CubicBody(ctx, 3, 9, (33, 20, 70), [0, 0, 0], [0.1, 0.1, 0.1], 2)
    full = rng.normal(size=(37, 24, 74, 9))
    body.upload(full, with_ghosts=True)
    assert np.array_equal(body.download_box((3, -2, 10), (7, 24, 33)), full[5:12, 0:24, 12:45])
    assert np.array_equal(body.download_box((0, 0, 0), (33, 20, 70)), full[2:-2, 2:-2, 2:-2])
    body.close()
    ctx.close()


def test_vtk_snapshots_on_gpu_match_reference_fields(lib, tmp_path, monkeypatch):
    """VtkSnapshotter (util/snapshot/VtkSnapshotter.hpp:20-61) on the GPU engine, read back asynchronously while the time
    loop goes on: every .vts of the run is parsed and its float32 fields are compared with (a) the state the engine held at
    that step and (b) the values the reference's snapshotter writes, i.e. float32 of the reference's own state at that step
    (fixtures of the unmodified engine stopped at each snapshot step)."""
    from helpers import golden, read_vtk_appended
    monkeypatch.chdir(tmp_path)
    base = SCENARIOS["elastic3d_layers"]          # 20^3, 5 steps in the fixture
    g = golden("elastic3d_layers")
    text = base + "vtk every 1 PRESSURE Sxy\noutput vtkrun\n"
    eng = capi.HostEngine(lib, text)
    D, M, sizes, _ = eng.body_info(0)
    states = {0: eng.body_pde(0).copy()}
    for step in range(1, 6):
        eng.advance(1)                            # writes snapshot `step` asynchronously ...
        states[step] = eng.body_pde(0).copy()     # ... while we already use the engine again
    eng.close()                                   # the last snapshot is written when the engine ends
    files = sorted(os.listdir(tmp_path / "snapshots" / "vtkrun" / "vtk"))
    assert files == ["mesh0core00snap%04d.vts" % s for s in range(1, 6)], files
    nx, ny, nz = (int(s) for s in sizes)
    for step in range(1, 6):
        v = read_vtk_appended(tmp_path / "snapshots" / "vtkrun" / "vtk" / ("mesh0core00snap%04d.vts" % step))
        u = states[step].reshape(nx, ny, nz, M).transpose(2, 1, 0, 3).reshape(-1, M)   # VTK order: x fastest
        assert np.array_equal(v["Velocity"], u[:, :3].astype(np.float32))
        assert np.array_equal(v["Sxy"], u[:, 4].astype(np.float32))
        assert np.array_equal(v["pressure"], (-(u[:, 3] + u[:, 6] + u[:, 8]) / 3).astype(np.float32))
    # the last step is the reference fixture's final state: the file holds float32 of the REFERENCE's values
    ref = g["body0"].reshape(nx, ny, nz, M).transpose(2, 1, 0, 3).reshape(-1, M)
    assert np.array_equal(v["Velocity"], ref[:, :3].astype(np.float32))
    assert np.array_equal(v["pressure"], (-(ref[:, 3] + ref[:, 6] + ref[:, 8]) / 3).astype(np.float32))


def test_full_size_anchor_1024(lib):
    """The BASELINE headline task at its full size (1024^3, offsets beyond 2^33 bytes) against thin bodies cut out of the
    same medium: after 3 steps every value depends on initial data at most 3 * border_size nodes away along each axis, so
    the centre line of a 17-node-thick body placed inside the big one must reproduce the big run BIT FOR BIT, and the thin
    bodies are small enough for the CPU oracle (pinned to the reference), which they must equal as well."""
    import torch
    sys.path.insert(0, ROOT)
    import bench
    import oracle_host as oh
    free, _ = torch.cuda.mem_get_info()
    n = bench.pick_size(free, 1024)
    if n < 1024:
        pytest.skip("needs a GPU with room for the 1024^3 state (%d fits)" % n)
    steps = 3
    big = capi.HostEngine(lib, bench.task_text(n, n, n, steps, detector=False)).run()
    ctxh, bodyh = big.context_handle(), big.body_handle(0)

    def column(lo, ext):
        out = np.empty(tuple(ext) + (9,))
        lo_a, ext_a = np.array(lo, dtype=np.int32), np.array(ext, dtype=np.int32)
        lib.check(lib.c.gcmb_cubic_download_box_begin(bodyh, capi.ip(lo_a), capi.ip(ext_a), out.ctypes.data_as(capi.vp)))
        lib.check(lib.c.gcmb_cubic_download_box_end(bodyh))
        return out

    try:
        # thin bodies: 17 x 17 x n along z (crossing the pulse and the free surface), 17 x n x 17 along y (crossing all four
        # material layers), n x 17 x 17 along x -- placed away from the big body's faces except along their long axis
        for sizes, start in (((17, 17, n), (600, 300, 0)), ((17, n, 17), (900, 0, 400)), ((n, 17, 17), (0, 700, 500))):
            text = bench.task_text(n, n, n, steps, detector=False).replace(
                "sizes %d %d %d start 0 0 0" % (n, n, n), "sizes %d %d %d start %d %d %d" % (sizes + start))
            assert "start %d" % start[0] in text
            thin = capi.HostEngine(lib, text).run()
            got = thin.body_pde(0).reshape(sizes + (9,))
            ora = oh.run_task_text(text)
            assert thin.info()[0] == ora.steps_done == steps
            assert np.array_equal(got, ora.real_nodes(0).reshape(sizes + (9,))), "thin body differs from the CPU oracle"
            long_axis = int(np.argmax(sizes))
            lo = [start[i] + (8 if i != long_axis else 0) for i in range(3)]
            ext = [1 if i != long_axis else n for i in range(3)]
            line_big = column(lo, ext).reshape(n, 9)
            sel = [8, 8, 8]
            sel[long_axis] = slice(None)
            line_thin = got[tuple(sel)]
            assert np.abs(line_thin).max() > 0
            assert np.array_equal(line_big, line_thin), (sizes, np.abs(line_big - line_thin).max())
            thin.close()
    finally:
        big.close()


def _oracle_engine(text):
    import oracle_host as oh
    return oh.run_task_text(text)


@pytest.mark.parametrize("maker,kw", [(elastic3d_iso, dict(n=48, steps=6)), (acoustic3d_free, dict(n=64, steps=8)),
                                      (elastic3d_layers, dict(n=40, steps=10))])
def test_cuda_engine_matches_oracle_at_medium_size(lib, maker, kw):
    """sizes beyond the committed fixtures: CUDA vs the (reference-pinned) CPU oracle, bitwise."""
    text = maker(**kw)
    ora = _oracle_engine(text)
    eng = run_engine(lib, text)
    got = eng.body_pde(0)
    ref = ora.real_nodes(0)
    assert eng.info()[0] == ora.steps_done
    assert np.array_equal(ref, got), np.abs(ref - got).max()
    eng.close()


def test_random_state_single_stages_match_oracle(lib):
    """C-ABI level: random state with ghosts, random material map, each direction, vs gcmo_stage."""
    from helpers import random_stage_check
    random_stage_check(lib, ((3, (19, 13, 37), "elastic", 2), (3, (9, 17, 150), "acoustic", 2),
                             (3, (5, 40, 700), "elastic", 2), (3, (300, 3, 33), "elastic", 1),
                             (2, (23, 131), "elastic", 2), (2, (33, 40), "elastic", 1), (2, (70, 515), "acoustic", 1),
                             (1, (300,), "elastic", 2), (1, (1000,), "acoustic", 2)))


def test_many_materials_stay_on_the_specialised_kernels(lib):
    """255 materials in one body: the packed tables of all of them sit in dynamic shared memory of the same kernels"""
    from helpers import random_stage_check
    random_stage_check(lib, ((3, (19, 13, 300), "elastic", 2), (3, (9, 40, 130), "acoustic", 2), (2, (23, 515), "elastic", 2)), n_materials=255)
    random_stage_check(lib, ((3, (19, 13, 300), "elastic", 2),), n_materials=40)


def test_ghost_layers_roundtrip_and_checksum(lib):
    ctx = capi.Context(lib)
    rng = np.random.default_rng(3)
    body = capi.CubicBody(ctx, 3, 9, (33, 20, 70), [0, 0, 0], [0.1, 0.1, 0.1], 2)
    full = rng.normal(size=(37, 24, 74, 9))
    body.upload(full, with_ghosts=True)
    assert np.array_equal(body.download(with_ghosts=True), full)
    real = full[2:-2, 2:-2, 2:-2]
    assert abs(body.checksum() - (real * np.arange(1, 10)).sum()) < 1e-9
    body.close()
    ctx.close()


def test_full_size_properties_acoustic_512(lib):
    """BASELINE config 2 at its full size (512^3 acoustic, free surfaces): properties the scheme guarantees.
    (a) the centred point source stays mirror-symmetric in x, y and z; (b) the same run on a 64^3 corner
    problem equals the oracle (covered above); (c) total checksum is finite and reproducible run to run."""
    text = acoustic3d_free(n=512, steps=3)
    eng = run_engine(lib, text)
    a = eng.body_pde(0).reshape(512, 512, 512, 4)
    p = a[..., 3]
    scale = np.abs(p).max()
    assert scale > 0
    for axis in range(3):
        assert np.abs(p - np.flip(p, axis=axis)).max() <= 1e-12 * scale
    v0 = a[..., 0]
    assert np.abs(v0 + np.flip(v0, axis=0)).max() <= 1e-12 * max(np.abs(v0).max(), 1e-300)
    chk1 = float((a * np.arange(1, 5)).sum())
    eng.close()
    del a, p, v0
    eng2 = run_engine(lib, text)
    b = eng2.body_pde(0)
    chk2 = float((b.reshape(512, 512, 512, 4) * np.arange(1, 5)).sum())
    assert chk1 == chk2
    eng2.close()


def test_uniform_state_is_a_fixed_point(lib):
    """a spatially constant PDE vector is reproduced by every stage (interpolation of constants is exact)."""
    ctx = capi.Context(lib)
    U, U1, L = capi.host_matrices(lib, "elastic", 3, ("isotropic", 2.0, 3.0, 1.5))
    body = capi.CubicBody(ctx, 3, 9, (40, 36, 200), [0, 0, 0], [0.1, 0.1, 0.1], 2)
    body.set_materials(U[None], U1[None], L[None])
    vec = np.arange(1.0, 10.0)
    state = np.broadcast_to(vec, (44, 40, 204, 9)).copy()
    body.upload(state, with_ghosts=True)
    tau = 0.9 * 0.1 / np.abs(L).max()
    for s in range(3):
        body.stage(s, tau)
    out = body.download(with_ghosts=False)
    # only the first stage sees constant ghosts (the other time layer's ghosts are zero, like the
    # reference's): nodes within the stencil of the y/z faces are excluded
    assert np.abs(out[:, 2:-2, 2:-2] - vec).max() <= 1e-13 * 9
    body.close()
    ctx.close()


def test_reference_run_statement(lib):
    """src/test/sequence/TestEngine.cpp:91-136 on the CUDA engine"""
    from reference_engine_cases import run_statement
    run_statement(lib)


@pytest.mark.parametrize("vary", ["rho", "E"])
def test_reference_two_layers(lib, vary):
    """src/test/sequence/TestEngine.cpp:139-296 on the CUDA engine, and bitwise equal to the oracle run"""
    import oracle_host as oh
    from reference_engine_cases import two_layers
    for (steps, rs, rs_theory, rv, rv_theory, init, reflect) in two_layers(lib, vary):
        assert abs(rs - rs_theory) < 1e-2 and abs(rv - rv_theory) < 1e-2, (rs, rs_theory, rv, rv_theory)


# ---- simplex path on the device; oracle: oracle/simplex_oracle.c (pinned by tests/golden/simplex_*.npz) --------
def test_simplex_vertex_info_gpu(lib):
    import simplex_cases
    simplex_cases.check_vertex_info(lib)
    simplex_cases.check_vertex_info(lib, "regular")


def test_simplex_cell_location_protocol_gpu(lib):
    """the full TestLineWalkSearch3D.cpp:120-154 protocol: every vertex x 16x16 directions x 9 lengths"""
    import simplex_cases
    hist = simplex_cases.check_locate_protocol(lib, "jitter_void", n_dirs=16, lengths=9)
    assert hist[3] > 0 and hist[0] > 0
    simplex_cases.check_locate_protocol(lib, "regular", n_dirs=16, lengths=9)


def test_simplex_cell_location_big_mesh_gpu(lib):
    import numpy as np
    import simplex_cases
    rng = np.random.default_rng(0)
    m = simplex_cases.make_mesh(lib, "big")
    vs = np.sort(rng.choice(m.n_local, 600, replace=False))
    simplex_cases.check_locate_protocol(lib, "big", n_dirs=8, lengths=6, vertices=vs)


def test_simplex_direction_masks_keep_the_answer_gpu(lib):
    import simplex_cases
    simplex_cases.check_direction_masks(lib)


def test_simplex_gradient_gpu(lib):
    import simplex_cases
    simplex_cases.check_gradient(lib)
    simplex_cases.check_gradient(lib, "regular")


@pytest.mark.parametrize("model", [0, 1])
def test_simplex_time_steps_gpu(lib, model):
    import simplex_cases
    simplex_cases.check_stage(lib, model, steps=4)


@pytest.mark.parametrize("model", [0, 1])
def test_simplex_zero_stays_zero_gpu(lib, model):
    import simplex_cases
    simplex_cases.check_stage(lib, model, kind="regular", steps=3, zero=True)


@pytest.mark.parametrize("model", [0, 1])
@pytest.mark.parametrize("kind", ["layers", "layers_void"])
def test_simplex_two_bodies_in_contact_gpu(lib, model, kind):
    import simplex_cases
    simplex_cases.check_two_bodies(lib, model, steps=3, kind=kind)


@pytest.mark.parametrize("model,bodies,basis,cavity", [(0, 2, "identity", True), (1, 2, "rotated", True), (0, 1, "random", True),
                                                      (0, 2, "random", False), (1, 1, "identity", False)])
def test_simplex_engine_gpu(lib, model, bodies, basis, cavity):
    """simplex::Engine of the host layer on the GPU == the oracle driven in the reference's order, bit for bit"""
    import simplex_cases
    simplex_cases.check_engine(lib, model, bodies=bodies, basis=basis, cavity=cavity, steps=4)


@pytest.mark.parametrize("model", [0, 1])
def test_simplex_pde_vectors_gpu(lib, model):
    """GcmType::ADVECT_PDE_VECTORS: two bodies in contact, then the engine with a random basis"""
    import simplex_cases
    simplex_cases.check_two_bodies(lib, model, steps=3, kind="layers_void", gcm_type=1)
    simplex_cases.check_engine(lib, model, bodies=2, basis="random", cavity=True, steps=3, gcm_type=1)


@pytest.mark.parametrize("task", ["cubic2d", "cubic3d", "acoustic", "ndi_empty", "ndi", "titan", "cubeAcs", "cubeEls"])
def test_launcher_gpu(task, tmp_path):
    """gcm_b200/gcmb_exe --task <id>: the reference launcher's cubic demo tasks (src/launcher/main.cpp:332-467) on
    and of ndi.hpp:162-317 on the GPU against the unmodified reference's step count, end time and the checksum of every body"""
    import json
    import re
    exe = os.path.join(ROOT, "gcm_b200", "gcmb_exe")
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "launcher_tasks.json")))[task]
    out = subprocess.run([exe, "--task", task, "-q"], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert out.returncode == 0, out.stderr
    steps, time = re.search(r"steps = (\d+), time = (\S+)", out.stdout).groups()
    assert int(steps) == gold["steps"] and float(time) == gold["time"]
    for body, want in gold.get("bodies", {"0": gold}).items():
        checksum = float(re.search(r"body %s (?:vertices = \d+ )?checksum = (\S+)" % body, out.stdout).group(1))
        assert abs(checksum - want["checksum"]) <= 1e-10 * want["abs_sum"], (task, body)


def test_launcher_simplex_plate_gpu(tmp_path):
    import re
    exe = os.path.join(ROOT, "gcm_b200", "gcmb_exe")
    out = subprocess.run([exe, "--task", "simplex_plate", "-q"], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert out.returncode == 0, out.stderr
    assert re.search(r"steps = 50,", out.stdout)
    assert "would have thrown on = 0" in out.stdout
    checksum = float(re.search(r"checksum = (\S+)", out.stdout).group(1))
    assert np.isfinite(checksum) and checksum != 0


import simplex_cases as _sx  # noqa: E402


@pytest.mark.parametrize("name", _sx.GOLDEN_SIMPLEX + _sx.GOLDEN_SIMPLEX_LOCAL)
def test_simplex_cuda_engine_matches_reference_bitwise(lib, name, tmp_path):
    """the CUDA simplex path == the UNMODIFIED reference simplex engine (compiled against a CGAL stand-in,
    tests/golden/make_simplex_golden.py): time step, step count and every PDE value of every body"""
    _sx.check_engine_against_reference(lib, name, tmp_path)


def test_simplex_cuda_cell_location_matches_reference(lib):
    _sx.check_locate_against_reference(lib)
