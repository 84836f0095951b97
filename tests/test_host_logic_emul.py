"""Host logic + kernel index math on the build machine (no GPU).

Runs the product's host engine (gcm_b200/host) against tests/emul's stepping harness — the product's own
CUDA sources executed thread by thread on the CPU — and requires bit-for-bit agreement with the outputs of
the unmodified reference (tests/golden).  This is NOT the parity claim for the CUDA path (that is
tests/test_gpu_parity.py, marked gpu); it proves the host layer, the table builder, the sparsity-pattern
selection and the per-thread index arithmetic before GPU time is spent."""
import os

import numpy as np
import pytest

from gcm_b200 import capi
from helpers import compare_with_golden, emul_library, golden, run_engine
from scenarios import SCENARIOS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    return emul_library()


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_engine_matches_reference_bitwise(lib, name):
    eng, _ = compare_with_golden(lib, name, SCENARIOS[name])
    eng.close()


EXPECTED_KERNELS = {
    "elastic3d_iso": ["sparse:elastic3d_iso_x/bs2", "sparse:elastic3d_iso_y/bs2", "sparse:elastic3d_iso_z/bs2"],
    "elastic3d_iso_bs1": ["sparse:elastic3d_iso_x/bs1", "sparse:elastic3d_iso_y/bs1", "sparse:elastic3d_iso_z/bs1"],
    "elastic3d_layers": ["sparse:elastic3d_iso_x/bs2", "sparse:elastic3d_iso_y/bs2", "sparse:elastic3d_iso_z/bs2"],
    "elastic3d_ortho": ["sparse:elastic3d_ortho_x/bs2", "sparse:elastic3d_ortho_y/bs2", "sparse:elastic3d_ortho_z/bs2"],
    "acoustic3d_free": ["sparse:acoustic3d_x/bs2", "sparse:acoustic3d_y/bs2", "sparse:acoustic3d_z/bs2"],
    "elastic2d_pwave": ["sparse:elastic2d_iso_x/bs2", "sparse:elastic2d_iso_y/bs2"],
    "elastic2d_courant45": ["dense:M5", "dense:M5"],  # border size 5: no specialised kernel
    "elastic1d": ["sparse:elastic1d_iso_x/bs2"],
    # rotated axes: no zero of U is structural; one material: the coefficients travel as kernel parameters
    "elastic3d_ortho_rotated": ["dense_k0_one:M9/bs2", "dense_k0_one:M9/bs2", "dense_k0_one:M9/bs2"],
    # Courant number 1 (feet in the second cell) and border size 3 stay on the specialised kernels
    "elastic3d_layers_courant1": ["sparse:elastic3d_iso_x/bs2+k0", "sparse:elastic3d_iso_y/bs2+k0", "sparse:elastic3d_iso_z/bs2+k0"],
    "acoustic3d_courant1": ["sparse:acoustic3d_x/bs2+k0", "sparse:acoustic3d_y/bs2+k0", "sparse:acoustic3d_z/bs2+k0"],
    "elastic2d_courant1": ["sparse:elastic2d_iso_x/bs2+k0", "sparse:elastic2d_iso_y/bs2+k0"],
    "ortho3d_contact_courant1": ["sparse:elastic3d_ortho_x/bs2+k0", "sparse:elastic3d_ortho_y/bs2+k0", "sparse:elastic3d_ortho_z/bs2+k0"],
    "elastic3d_ortho_rotated_courant1": ["dense_k0_one:M9/bs2+k0", "dense_k0_one:M9/bs2+k0", "dense_k0_one:M9/bs2+k0"],
    "elastic3d_iso_bs3": ["sparse:elastic3d_iso_x/bs3+k0", "sparse:elastic3d_iso_y/bs3+k0", "sparse:elastic3d_iso_z/bs3+k0"],
    "elastic2d_bs3_courant15": ["sparse:elastic2d_iso_x/bs3+k0", "sparse:elastic2d_iso_y/bs3+k0"],
}


@pytest.mark.parametrize("name", ["elastic3d_layers", "acoustic3d_free", "elastic2d_pwave", "ortho3d_contact", "elastic3d_layers_courant1",
                                  "elastic2d_bs3_courant15", "elastic3d_ortho_rotated", "acoustic2d_border1", "maxwell3d", "elastic1d"])
def test_fp32_engine_close_to_reference(lib, name):
    """real_bytes = 4 (the reference's `real` = float build, util/infrastructure/Types.hpp:8-14): the same engine in
    single precision stays within a few float roundings of the fp64 reference over these short runs"""
    g = golden(name)
    eng = capi.HostEngine(lib, SCENARIOS[name], real_bytes=4).run()
    try:
        assert eng.info()[0] == int(g["steps"]) and eng.info()[2] == float(g["tau"])
        bid = 0
        while "body%d" % bid in g.files:
            ref, got = g["body%d" % bid], eng.body_pde(bid)
            assert np.abs(ref - got).max() <= 5e-5 * np.abs(ref).max()
            bid += 1
        assert all(eng.kernel_name(0, d).endswith("/f32") for d in range(eng.body_info(0)[0]))
    finally:
        eng.close()


def test_async_box_download(lib):
    ctx = capi.Context(lib)
    rng = np.random.default_rng(5)
    for real_bytes in (8, 4):
        c = capi.Context(lib, real_bytes=real_bytes)
        body = capi.CubicBody(c, 3, 9, (9, 6, 20), [0, 0, 0], [0.1, 0.1, 0.1], 2)
        full = rng.normal(size=(13, 10, 24, 9)).astype(c.real)
        body.upload(full, with_ghosts=True)
        assert np.array_equal(body.download_box((3, -2, 10), (5, 10, 7)), full[5:10, 0:10, 12:19])
        body.close()
        c.close()
    ctx.close()


@pytest.mark.parametrize("name", sorted(EXPECTED_KERNELS))
def test_sparsity_pattern_selection(lib, name):
    from gcm_b200 import capi
    eng = capi.HostEngine(lib, SCENARIOS[name])
    try:
        D = eng.body_info(0)[0]
        assert [eng.kernel_name(0, d) for d in range(D)] == ["unset"] * D
        eng.advance(1)  # the name is that of the kernel actually launched
        assert [eng.kernel_name(0, d) for d in range(D)] == EXPECTED_KERNELS[name]
    finally:
        eng.close()


@pytest.mark.parametrize("env", [{"GCMB_FORCE_DENSE": "1"}, {"GCMB_FORCE_DENSE": "1", "GCMB_DENSE_LITERAL": "1"}, {"GCMB_MARCH_SEG": "7"}, {"GCMB_ZTILE_ROWS": "5"}])
@pytest.mark.parametrize("name", ["elastic3d_layers", "ortho3d_contact", "acoustic2d_border1", "elastic2d_ortho", "ortho3d_rotated_plies"])
def test_kernel_variants_agree(name, env):
    """dense kernel, direct kernel and short marching segments give the same bits (fresh process: the
    variant is read from the environment once)."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from helpers import compare_with_golden, emul_library\n"
            "from scenarios import SCENARIOS\n"
            "compare_with_golden(emul_library(), %r, SCENARIOS[%r])\n"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)),
               os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"), name, name))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("real_bytes", [8, 4])
def test_border_inside_contiguous_stage_equals_separate_fill(lib, real_bytes):
    """gcmb_cubic_stage_with_border == border fill + stage on the real nodes, bit for bit (row lengths around the tile and warp
    boundaries, one-sided and repeated conditions)"""
    from helpers import tile_border_check
    assert tile_border_check(lib, real_bytes) == 14


def test_deferred_border_fill_keeps_the_reference_order(lib):
    """gcmb_cubic_border_apply defers the fill of the last direction's faces to the stage kernel; whoever looks at the ghost
    nodes before the stage must find them filled, a second request replaces the first, and a stage of ANOTHER direction in
    between does not lose the condition"""
    rng = np.random.default_rng(3)
    ctx = capi.Context(lib)
    D, M, sizes, bs = 3, 9, (4, 5, 40), 2
    ms = capi.host_matrices(lib, "elastic", D, ("isotropic", 2.0, 3.0, 1.0))
    U, U1, Lm = (np.ascontiguousarray(m[None]) for m in ms)
    h = np.array([1.0, 1.1, 0.9])
    tau = 0.4 * h.min() / np.abs(Lm).max()
    state = rng.normal(size=tuple(s + 2 * bs for s in sizes) + (M,))
    table = np.zeros(sizes, dtype=np.uint8)

    def body_with(values_list, between=None):
        """border_apply(z) for every entry of values_list, then `between`, then download with ghosts"""
        b = capi.CubicBody(ctx, D, M, sizes, [0] * D, h, bs)
        b.set_materials(U, U1, Lm, table)
        b.border_set_area(0, 2, ("infinite",), [5, 7, 8])
        b.upload(state, with_ghosts=True)
        for v in values_list:
            b.border_apply(2, v)
        if between:
            between(b)
        out = b.download(with_ghosts=True)
        b.close()
        return out

    # what the fill kernel leaves (BorderConditions.hpp:97-114): ghost(-a) = inner(+a), the set components -inner + 2 b
    plain_lib_env = state.copy()
    n2 = sizes[2]
    for a in range(1, bs + 1):
        for ghost, inner in ((bs - a, bs + a), (bs + n2 - 1 + a, bs + n2 - 1 - a)):
            row = state[bs:-bs, bs:-bs, inner, :].copy()
            for c, val in zip((5, 7, 8), (0.1, 0.2, 0.3)):
                row[..., c] = -row[..., c] + 2 * val
            plain_lib_env[bs:-bs, bs:-bs, ghost, :] = row
    # a reader of the ghost nodes runs the deferred fill first
    assert np.array_equal(body_with([[0.1, 0.2, 0.3]]), plain_lib_env)
    # the second request wins, as two fill kernels one after the other would
    assert np.array_equal(body_with([[9.0, 9.0, 9.0], [0.1, 0.2, 0.3]]), plain_lib_env)
    # stage x in between: the fill happens before it (the ghost nodes of the layer it leaves behind are those of the request)
    a = body_with([[0.1, 0.2, 0.3]], between=lambda b: (b.stage(0, tau), b.stage(2, tau)))
    b2 = capi.CubicBody(ctx, D, M, sizes, [0] * D, h, bs)
    b2.set_materials(U, U1, Lm, table)
    b2.border_set_area(0, 2, ("infinite",), [5, 7, 8])
    b2.upload(plain_lib_env, with_ghosts=True)     # the state with the ghost nodes already filled
    b2.stage(0, tau); b2.stage(2, tau)
    real = tuple(slice(bs, bs + s) for s in sizes)
    assert np.array_equal(a[real], b2.download(with_ghosts=True)[real])
    b2.close()
    ctx.close()


@pytest.fixture(scope="module")
def binding_on_the_harness(lib):
    """oracle/_ref/gcm_ref_gpu_emul: the UNMODIFIED reference cubic::Engine<D> with integration/GpuBackend.hpp as its backend, linked
    against the stepping harness instead of the CUDA library (built where /root/reference exists)"""
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "gcm_ref_gpu_emul")
    if os.path.isdir("/root/reference/src"):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref_gpu_emul"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/gcm_ref_gpu_emul is built where /root/reference exists (make -C oracle ref_gpu_emul)")
    return exe


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_reference_engine_with_gcm_b200_backend_on_the_harness(binding_on_the_harness, name, tmp_path):
    """The drop-in without a GPU: the reference's own engine loop drives gcm_b200 through the binding a maintainer would add
    (GpuMesh / GpuGcm / GpuBorder / GpuContact / GpuMaxwellOde), the library's host logic and kernel bodies run on the
    stepping harness, and the result must equal the all-CPU reference (tests/golden) bit for bit.  The binding registers every
    border condition as face masks: whole-face masks take the fill inside the stage kernel (gcmb_cubic_border_set)."""
    import oracle_host as oh
    g = golden(name)
    out = oh.run_reference(SCENARIOS[name], str(tmp_path), exe_name="gcm_ref_gpu_emul")
    assert out["meta"]["tau"] == float(g["tau"]) and out["meta"]["time"] == float(g["time"]) and int(out["meta"]["steps"]) == int(g["steps"])
    bid = 0
    while "body%d" % bid in g.files:
        assert np.array_equal(out[bid], g["body%d" % bid]), (name, bid)
        bid += 1
    if "detector" in g.files:
        assert np.array_equal(out["detector"], g["detector"])


def test_whole_face_host_masks_count_as_whole_faces(lib):
    """the reference-side binding registers every condition as face masks (integration/GpuBackend.hpp): masks that select the
    whole face must take the same path as an infinite area -- the fill inside the stage kernel -- and give the same bits;
    a mask with a hole goes through the fill kernel"""
    rng = np.random.default_rng(4)
    ctx = capi.Context(lib)
    D, M, sizes, bs = 3, 9, (4, 5, 40), 2
    ms = capi.host_matrices(lib, "elastic", D, ("isotropic", 2.0, 3.0, 1.0))
    U, U1, Lm = (np.ascontiguousarray(m[None]) for m in ms)
    h = np.array([1.0, 1.1, 0.9])
    tau = 0.4 * h.min() / np.abs(Lm).max()
    state = rng.normal(size=tuple(s + 2 * bs for s in sizes) + (M,))
    table = np.zeros(sizes, dtype=np.uint8)
    face = sizes[0] * sizes[1]
    out = {}
    for kind in ("area", "full_masks", "holed_masks"):
        b = capi.CubicBody(ctx, D, M, sizes, [0] * D, h, bs)
        b.set_materials(U, U1, Lm, table)
        if kind == "area":
            b.border_set_area(0, 2, ("infinite",), [5, 7, 8])
        else:
            mask = np.ones(face, dtype=np.uint8)
            if kind == "holed_masks":
                mask[3] = 0
            b.border_set(0, 2, mask, mask, [5, 7, 8])
        b.upload(state, with_ghosts=True)
        took = b.stage_with_border(2, tau, [0.1, 0.2, 0.3])
        assert took == (kind != "holed_masks"), kind
        out[kind] = b.download()
        b.close()
    assert np.array_equal(out["area"], out["full_masks"])
    assert not np.array_equal(out["area"], out["holed_masks"])
    ctx.close()


def test_adhesion_contact_equals_single_body(lib):
    """test/sequence/TestEngine.cpp:27-87 through our engine: two glued bodies == one body, bitwise."""
    two = run_engine(lib, SCENARIOS["adhesion2d_two"])
    one = run_engine(lib, SCENARIOS["adhesion2d_one"])
    full = one.body_pde(0).reshape(11, 26, 5)
    assert np.array_equal(full[:, :13], two.body_pde(0).reshape(11, 13, 5))
    assert np.array_equal(full[:, 13:], two.body_pde(1).reshape(11, 13, 5))
    two.close()
    one.close()


def test_upload_download_roundtrip(lib):
    from gcm_b200 import capi
    ctx = capi.Context(lib)
    rng = np.random.default_rng(0)
    for D, sizes, M, bs in ((3, (5, 4, 7), 9, 2), (2, (6, 9), 5, 3), (1, (11,), 2, 1)):
        body = capi.CubicBody(ctx, D, M, sizes, [0] * D, [0.1] * D, bs)
        full = rng.normal(size=tuple(s + 2 * bs for s in sizes) + (M,))
        body.upload(full, with_ghosts=True)
        assert np.array_equal(body.download(with_ghosts=True), full)
        real = tuple(slice(bs, bs + s) for s in sizes)
        assert np.array_equal(body.download(with_ghosts=False), full[real])
        chk = body.checksum()
        want = (full[real] * np.arange(1, M + 1)).sum()
        assert abs(chk - want) <= 1e-12 * np.abs(full).sum()
        body.close()
    ctx.close()


def test_courant_above_border_size_is_rejected(lib):
    """EqualDistanceLineInterpolator.hpp:22 asserts k <= size-1: the reference throws, so do we."""
    from gcm_b200 import capi
    bad = SCENARIOS["elastic2d_pwave"].replace("courant 0.9", "courant 3.5")
    eng = capi.HostEngine(lib, bad)
    with pytest.raises(capi.GcmError):
        eng.run()
    eng.close()


def test_rotated_orthotropic_error_behaviour(lib):
    """ElasticModel3D.cpp:151-283 through GslUtils.hpp:163-204: a characteristic cubic with complex roots is rejected
    (the reference's fallback needs gsl_poly_complex_solve_cubic, which neither the oracle build nor the product
    provides), and rotated axes exist in 3-D only -- the same on the product's host and in the oracle."""
    import oracle_host as oh
    from gcm_b200 import capi
    from scenarios import elastic2d_ortho, elastic3d_ortho_rotated
    complex_roots = elastic3d_ortho_rotated(8, 1).replace("orthotropic 4 360 70 70 180 70 90 10 20 30", "orthotropic 4 4 2 2 4 2 4 1 1 1")
    assert complex_roots != elastic3d_ortho_rotated(8, 1)
    flat = elastic2d_ortho()
    line = [l for l in flat.splitlines() if l.startswith("material")][0]
    flat = flat.replace(line, line + " angles 0.1 0 0")
    for text in (complex_roots, flat):
        with pytest.raises(capi.GcmError):
            capi.HostEngine(lib, text)
        with pytest.raises(ValueError):
            oh.run_task_text(text)


def test_random_state_single_stages_match_oracle(lib):
    """multi-chunk rows, ragged sizes, first-order border: stepping harness vs gcmo_stage, bitwise"""
    from helpers import random_stage_check
    random_stage_check(lib, ((3, (7, 6, 37), "elastic", 2), (3, (3, 35, 300), "elastic", 2), (3, (260, 2, 9), "acoustic", 1),
                             (2, (9, 515), "elastic", 1), (2, (300, 8), "acoustic", 2), (1, (600,), "elastic", 2)))


def test_reference_run_statement(lib):
    """src/test/sequence/TestEngine.cpp:91-136 on our engine (dense kernel: border size 5, Courant 4.5)"""
    from reference_engine_cases import run_statement
    assert run_statement(lib) == ["dense:M5", "dense:M5"]


@pytest.mark.parametrize("vary", ["rho", "E"])
def test_reference_two_layers(lib, vary):
    """src/test/sequence/TestEngine.cpp:139-296 on our engine; the oracle must see the same numbers bit for bit,
    and the reference's own 1e-2 assertions on the reflection coefficients must hold."""
    import oracle_host as oh
    from reference_engine_cases import two_layers
    for (steps, rs, rs_theory, rv, rv_theory, init, reflect) in two_layers(lib, vary):
        assert abs(rs - rs_theory) < 1e-2 and abs(rv - rv_theory) < 1e-2, (rs, rs_theory, rv, rv_theory)


# ---- simplex path on the stepping harness (oracle: oracle/simplex_oracle.c, pinned by tests/golden/simplex_*.npz) -----------------
def test_simplex_vertex_info(lib):
    import simplex_cases
    simplex_cases.check_vertex_info(lib)
    simplex_cases.check_vertex_info(lib, "regular")


def test_simplex_cell_location_protocol(lib):
    """src/test/sequence/TestLineWalkSearch3D.cpp:120-154 protocol, integer-exact vs the restatement"""
    import simplex_cases
    simplex_cases.check_locate_protocol(lib, "jitter_void", n_dirs=8, lengths=5)
    simplex_cases.check_locate_protocol(lib, "regular", n_dirs=8, lengths=5)


def test_simplex_direction_masks_keep_the_answer(lib):
    import simplex_cases
    simplex_cases.check_direction_masks(lib)


def test_simplex_gradient(lib):
    import simplex_cases
    simplex_cases.check_gradient(lib)


@pytest.mark.parametrize("model", [0, 1])
def test_simplex_time_steps(lib, model):
    """4 steps: the cache of located feet is keyed in step 1, filled in step 2 and used from step 3 on"""
    import simplex_cases
    simplex_cases.check_stage(lib, model, steps=4)


@pytest.mark.parametrize("model", [0, 1])
def test_simplex_zero_stays_zero(lib, model):
    """src/test/sequence/TestSimplexGcm.cpp:29-67"""
    import simplex_cases
    simplex_cases.check_stage(lib, model, kind="regular", steps=2, zero=True)


@pytest.mark.parametrize("model", [0, 1])
def test_simplex_two_bodies_in_contact(lib, model):
    import simplex_cases
    simplex_cases.check_two_bodies(lib, model, steps=2)


@pytest.mark.parametrize("model,bodies,basis", [(0, 2, "identity"), (1, 2, "rotated"), (0, 1, "random")])
def test_simplex_engine(lib, model, bodies, basis):
    """simplex::Engine of the host layer (task text -> run) == the oracle driven in the reference's order"""
    import simplex_cases
    simplex_cases.check_engine(lib, model, bodies=bodies, basis=basis, steps=2)


@pytest.mark.parametrize("model", [0, 1])
def test_simplex_pde_vectors_two_bodies(lib, model):
    """GcmType::ADVECT_PDE_VECTORS (engine/simplex/GridCharacteristicMethodInPdeVectors.hpp) with contacts"""
    import simplex_cases
    simplex_cases.check_two_bodies(lib, model, steps=2, kind="layers_void", gcm_type=1)


def test_simplex_pde_vectors_engine(lib):
    import simplex_cases
    simplex_cases.check_engine(lib, 0, bodies=2, basis="rotated", steps=4, gcm_type=1)


@pytest.mark.parametrize("task", ["cubic2d", "acoustic", "ndi_empty", "ndi", "cubeAcs", "cubeEls"])
def test_launcher_command_line(task, tmp_path):
    """gcmb_exe --task <id> (src/launcher/main.cpp:22-71) on the shipped demo tasks: step count, end time and state
    checksum against the unmodified reference's (tests/golden/launcher_tasks.json; cubeAcs / cubeEls, main.cpp:547-640,
    against the reference's simplex engine on the same triangulation)"""
    import json
    import re
    import subprocess
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
    import build_emul
    exe = build_emul.build_launcher_emul()
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "launcher_tasks.json")))[task]
    out = subprocess.run([exe, "--task", task, "-q"], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert out.returncode == 0, out.stderr
    steps, time = re.search(r"steps = (\d+), time = (\S+)", out.stdout).groups()
    assert int(steps) == gold["steps"] and float(time) == gold["time"]
    for body, want in gold.get("bodies", {"0": gold}).items():
        checksum = float(re.search(r"body %s (?:vertices = \d+ )?checksum = (\S+)" % body, out.stdout).group(1))
        assert abs(checksum - want["checksum"]) <= 1e-10 * want["abs_sum"], (task, body)
    bad = subprocess.run([exe, "--task", "no_such_task"], capture_output=True, text=True)
    assert bad.returncode != 0 and "Invalid task file" in bad.stderr


def test_vtk_snapshots_cubic(lib, tmp_path, monkeypatch):
    """VtkSnapshotter fields and file names (util/snapshot/VtkSnapshotter.hpp:20-61, Snapshotter.hpp:53-68) for a
    layered 3-D elastic body: .vts in VTK point order with Velocity, the requested scalars and material_index"""
    from helpers import read_vtk_appended
    from scenarios import elastic3d_layers
    monkeypatch.chdir(tmp_path)
    text = elastic3d_layers(n=10, steps=4) + "vtk every 2 PRESSURE Sxy\noutput run1\n"
    eng = capi.HostEngine(lib, text).run()
    D, M, sizes, _ = eng.body_info(0)
    u = eng.body_pde(0).reshape(tuple(sizes) + (M,))
    files = sorted(os.listdir(tmp_path / "snapshots" / "run1" / "vtk"))
    # numberOfSnaps = 4 snapshots of stepsPerSnap = 2 steps each, plus the initial one (AbstractEngine.cpp:30-46)
    assert files == ["mesh0core00snap%04d.vts" % s for s in (0, 2, 4, 6, 8)]
    v = read_vtk_appended(tmp_path / "snapshots" / "run1" / "vtk" / files[-1])
    nx, ny, nz = (int(s) for s in sizes)
    assert 'WholeExtent="0 %d 0 %d 0 %d"' % (nx - 1, ny - 1, nz - 1) in v["header"]
    vtk_order = u.transpose(2, 1, 0, 3).reshape(-1, M)          # x fastest
    assert np.array_equal(v["Velocity"], vtk_order[:, :3].astype(np.float32))
    assert np.array_equal(v["Sxy"], vtk_order[:, 4].astype(np.float32))
    pressure = -(vtk_order[:, 3] + vtk_order[:, 6] + vtk_order[:, 8]) / 3
    assert np.array_equal(v["pressure"], pressure.astype(np.float32))
    assert v["Points"].shape == (nx * ny * nz, 3) and v["Points"][1, 0] > v["Points"][0, 0] and v["Points"][1, 1] == v["Points"][0, 1]
    assert set(np.unique(v["material_index"])) <= {0.0, 1.0, 2.0, 3.0} and len(np.unique(v["material_index"])) >= 1
    eng.close()


def test_vtk_snapshots_simplex(lib, tmp_path, monkeypatch):
    from helpers import read_vtk_appended
    import simplex_cases
    monkeypatch.chdir(tmp_path)
    text, _ = simplex_cases.engine_scenario(0, bodies=2, steps=2)
    eng = capi.SimplexHostEngine(lib, text + "vtk PRESSURE\n").run()
    tri = eng.triangulation()
    for body in (0, 1):
        v = read_vtk_appended(tmp_path / "snapshots" / "vtk" / ("mesh%dcore00snap0002.vtu" % body))
        u = eng.simplex_pde(body)
        assert np.array_equal(v["Velocity"], u[:, :3].astype(np.float32))
        assert np.array_equal(v["pressure"], (-(u[:, 3] + u[:, 6] + u[:, 8]) / 3).astype(np.float32))
        n_cells = int((tri["cell_grid"] == body).sum())
        assert v["connectivity"].shape == (4 * n_cells,) and v["connectivity"].max() == len(u) - 1
        assert np.array_equal(v["offsets"], 4 * np.arange(1, n_cells + 1)) and (v["types"] == 10).all()
    eng.close()


def test_inm_mesh_file_round_trip(lib, tmp_path):
    """INM mesh files (grid/simplex/mesh_loaders/InmMeshLoader.hpp:96-168): a two-body box mesh with a cavity saved
    in that format and loaded back gives the same bodies, borders, contacts and, after 2 steps, the same bits"""
    import simplex_cases
    text, _ = simplex_cases.engine_scenario(0, bodies=2, steps=2)
    # one engine at a time: like the reference's, the Clock is a process-wide static (engine/GlobalVariables.hpp:16-40)
    a = capi.SimplexHostEngine(lib, text)
    path = tmp_path / "mesh.out"
    a.save_inm(path)
    head = open(path).read().split("\n")
    ta = a.triangulation()
    assert int(head[0]) == len(ta["xyz"]) and open(path).read().rstrip().endswith("\n0")
    a.run()
    ref = {body: a.simplex_pde(body) for body in (0, 1)}
    ref_contacts = a.contact_nodes(0, 1)[0]
    assert a.errors() == 0
    a.close()
    keep = [ln for ln in text.split("\n") if not ln.startswith(("simplex_box", "region", "cavity"))]
    b = capi.SimplexHostEngine(lib, "\n".join(keep) + "\nsimplex_mesh %s\n" % path)
    tb = b.triangulation()
    assert np.array_equal(ta["xyz"], tb["xyz"])
    assert (ta["cell_grid"] >= 0).sum() == len(tb["cell_grid"])       # empty cells are not stored in the file
    b.run()
    assert np.array_equal(ref_contacts, b.contact_nodes(0, 1)[0])
    for body in (0, 1):
        assert np.abs(ref[body]).max() > 0.1 and np.array_equal(ref[body], b.simplex_pde(body))
    assert b.errors() == 0
    b.close()


# ---- simplex path against the UNMODIFIED reference engine (fixtures: tests/golden/make_simplex_golden.py) ----------
import simplex_cases as _sx  # noqa: E402


@pytest.mark.parametrize("name", _sx.GOLDEN_SIMPLEX + _sx.GOLDEN_SIMPLEX_LOCAL)
def test_simplex_engine_matches_reference_bitwise(lib, name, tmp_path):
    _sx.check_engine_against_reference(lib, name, tmp_path)


@pytest.mark.parametrize("name", _sx.GOLDEN_SIMPLEX)
def test_simplex_oracle_matches_reference_bitwise(lib, name, tmp_path):
    _sx.check_oracle_against_reference(lib, name, tmp_path)


def test_simplex_cell_location_matches_reference(lib):
    _sx.check_locate_against_reference(lib, with_oracle=True)


@pytest.mark.parametrize("name", _sx.GOLDEN_SIMPLEX[:5])   # the fixtures made before the host had its own clean-up
def test_simplex_mesh_cleanup_matches_reference(lib, name):
    _sx.check_mesh_cleanup_against_reference(lib, name)


def test_slabs_refuse_contact_across_x(lib):
    """two bodies touching across x cannot be cut into x-slabs; the engine says so instead of dropping the contact"""
    text = SCENARIOS["ortho3d_contact"].replace("start 0 8 0", "start 16 0 0")
    with pytest.raises(capi.GcmError) as e:
        capi.HostEngine(lib, text, slab_rank=0, slab_count=2, nccl_id=bytes(128))
    assert e.value.code == -1 and "contact normal to x" in str(e.value)
    capi.HostEngine(lib, text).advance(1).close()  # undecomposed: fine


def test_inm_loader_reads_the_reference_fixture(lib):
    """src/test/sequence/TestInmMeshLoader.cpp:19-68 on our loader: the same file (tests/golden/testInmLoader.out is a byte copy of
    /root/reference/meshes/testInmLoader.out, compared here when the reference tree is present), the same assertions"""
    fixture = os.path.join(ROOT, "tests", "golden", "testInmLoader.out")
    original = "/root/reference/meshes/testInmLoader.out"
    if os.path.exists(original):
        assert open(original, "rb").read() == open(fixture, "rb").read()
    points, cells, materials = capi.inm_read(lib, fixture)
    assert len(points) == 12 and len(cells) == 3
    x = {3: -2.583210754394531250e+01, 4: 2.400143432617187500e+01, 5: 2.306465148925781250e+01, 6: 2.364865112304687500e+01,
         7: 2.377413177490234375e+01, 8: -7.073544311523437500e+01, 9: -6.941218566894531250e+01, 10: -6.654748535156250000e+01,
         11: -6.843022155761718750e+01}
    y = [4.283624267578125000e+01, 4.416368865966796875e+01, 4.302170562744140625e+01, 4.141203308105468750e+01, 1.096292877197265625e+01,
         1.127018737792968750e+01, 1.141680145263671875e+01, 1.079985809326171875e+01, 3.833699035644531250e+01, 4.111538696289062500e+01,
         4.088376617431640625e+01, 3.788503265380859375e+01]
    z = [1.406894775390625000e+03, 1.404953125000000000e+03, 1.405088378906250000e+03, 1.405093261718750000e+03, 1.388465698242187500e+03,
         1.388329711914062500e+03, 1.388676269531250000e+03, 1.387921020507812500e+03, 1.369011962890625000e+03, 1.365369384765625000e+03,
         1.370376220703125000e+03, 1.366383422851562500e+03]
    for i, v in x.items():
        assert points[i, 0] == v
    assert list(points[:, 1]) == y and list(points[:, 2]) == z
    # InmMeshLoader::Cell holds the file's 1-based vertex numbers: {1,2,3,4} -> 4, {5,6,7,8} -> 5, {9,10,11,12} -> 1
    got = {tuple(sorted(int(v) + 1 for v in c)): int(m) for c, m in zip(cells, materials)}
    assert got == {(1, 2, 3, 4): 4, (5, 6, 7, 8): 5, (9, 10, 11, 12): 1}
    # the denominator of Task::SimplexGrid::scale is applied to the points
    half, _, _ = capi.inm_read(lib, fixture, scale=2.0)
    assert np.array_equal(half, points / 2.0)


def test_ndi_task_as_written_is_refused_like_the_reference_does(lib):
    """src/launcher/ndi.hpp:241-242 starts the sample one node inside the prism; cubic::Engine throws "Bodies must not intersect"
    (engine/cubic/Engine.cpp:57) and so does ours -- the shipped ndi.task moves the sample one node down"""
    text = open(os.path.join(ROOT, "gcm_b200", "tasks", "ndi.task")).read().replace("start 0 -31", "start 0 -30")
    with pytest.raises(capi.GcmError) as e:
        capi.HostEngine(lib, text)
    assert e.value.code == 4 and "Bodies must not intersect" in str(e.value)  # gcm::Exception::BAD_MESH
    import oracle_host as oh
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "gcm_ref")):
        import subprocess
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            open(os.path.join(tmp, "t.txt"), "w").write(text)
            r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "gcm_ref"), os.path.join(tmp, "t.txt"), os.path.join(tmp, "out")],
                               capture_output=True, text=True, cwd=tmp)
        assert r.returncode != 0 and "Bodies must not intersect" in r.stderr + r.stdout


def test_many_materials_stay_on_the_specialised_kernels(lib):
    """the coefficient tables of ALL materials of a body live in dynamic shared memory: a body with 60 materials (round 1 switched
    to a slower kernel beyond 16) runs the same kernels and reproduces the oracle bit for bit"""
    from helpers import random_stage_check
    random_stage_check(lib, ((3, (7, 6, 40), "elastic", 2), (2, (9, 70), "acoustic", 2)), n_materials=60)
