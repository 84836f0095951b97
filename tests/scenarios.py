"""Task files (plain text) used by the parity tests.

Each entry is small enough for the CPU oracle to finish in seconds.  tests/golden/make_golden.py runs the
UNMODIFIED reference (oracle/_ref/gcm_ref) on every one of them and commits the raw fp64 results as
tests/golden/<name>.npz; the same text drives the C oracle and the CUDA engine.
"""


def _h(n, length=1.0):
    return repr(length / (n - 1))


def elastic3d_iso(n=20, steps=4, bs=2, courant=0.9):
    h = _h(n)
    return f"""
dimensionality 3
courant {courant}
border_size {bs}
h {h} {h} {h}
steps {steps}
body 0 elastic isotropic sizes {n} {n} {n} start 0 0 0
material default isotropic 4 2 1
initial quantity PRESSURE 1 sphere 0.3 0.5 0.5 0.5
"""


def acoustic3d_free(n=20, steps=5, courant=0.9, bs=2):
    # BASELINE config 2 at reduced size: point source, PRESSURE -> 0 on all six faces
    h = _h(n)
    return f"""
dimensionality 3
courant {courant}
border_size {bs}
h {h} {h} {h}
steps {steps}
body 0 acoustic isotropic sizes {n} {n} {n} start 0 0 0
material default isotropic 1 1 0
initial quantity PRESSURE 1 sphere 0.2 0.5 0.5 0.5
border 0 0 infinite PRESSURE const 0
border 0 1 infinite PRESSURE const 0
border 0 2 infinite PRESSURE const 0
"""


def elastic3d_layers(n=20, steps=5, courant=0.9, bs=2):
    # BASELINE config 3 at reduced size: y-layered medium, free surface on top (z right), detector disc
    h = _h(n)
    return f"""
dimensionality 3
courant {courant}
border_size {bs}
h {h} {h} {h}
steps {steps}
body 0 elastic isotropic sizes {n} {n} {n} start 0 0 0
material default isotropic 1 2 0.8
material area box -10 0.2 -10 10 0.4 10 isotropic 0.5 2 0.8
material area box -10 0.4 -10 10 0.7 10 isotropic 2 2 0.8
material area box -10 0.7 -10 10 10 10 isotropic 4 2 0.8
initial wave P_FORWARD 2 PRESSURE 1 box -10 -10 0.3 10 10 0.6
border 0 2 infinite Sxz const 0 Syz const 0 Szz const 0
detector 0 Vz sphere 0.4 0.5 0.5 1.0
"""


def elastic3d_ortho(n=18, steps=4):
    h = _h(n)
    return f"""
dimensionality 3
courant 0.9
border_size 2
h {h} {h} {h}
steps {steps}
body 0 elastic orthotropic sizes {n} {n} {n} start 0 0 0
material default orthotropic 4 360 70 70 180 70 90 10 10 10
initial quantity PRESSURE 1 sphere 0.3 0.5 0.5 0.5
"""


def elastic3d_ortho_rotated(n=14, steps=4, courant=0.9):
    # rotated axes of the material (ElasticModel3D.cpp:151-283): dense eigen-systems from the cubic's roots
    h = _h(n)
    return f"""
dimensionality 3
courant {courant}
border_size 2
h {h} {h} {h}
steps {steps}
body 0 elastic orthotropic sizes {n} {n} {n} start 0 0 0
material default orthotropic 4 360 70 70 180 70 90 10 20 30 angles 0.3 -0.5 1.1
initial quantity PRESSURE 1 sphere 0.3 0.5 0.5 0.5
"""


def ortho3d_contact(n=16, steps=5, courant=0.9, bs=2):
    # BASELINE config 4 at reduced size: two orthotropic bodies glued along y + fixed velocity on a disc
    h = _h(n)
    half = n // 2
    return f"""
dimensionality 3
courant {courant}
border_size {bs}
h {h} {h} {h}
steps {steps}
body 0 elastic orthotropic sizes {n} {half} {n} start 0 0 0
body 1 elastic orthotropic sizes {n} {half} {n} start 0 {half} 0
material body 0 orthotropic 1580 10.30e9 6.96e9 6.96e9 23.25e9 6.96e9 10.30e9 5.01e9 1.67e9 5.01e9
material body 1 orthotropic 4 360 70 70 180 70 90 10 10 10
initial quantity PRESSURE 1 sphere 0.35 0.5 0.5 0.5
border 1 1 sphere 0.3 0.5 1.0 0.5 Vy sin 1.0 5.0
"""


def elastic3d_contact_z(n=14, steps=5, courant=0.9, bs=2):
    # two bodies glued across the LAST direction, both with a free-surface condition on their z faces: the reference applies
    # the border condition to the whole face first and the contact copy then overwrites the ghost nodes of the glued face
    # (cubic/Engine.cpp:94-107) -- the order a deferred border fill has to keep
    h = _h(n)
    half = n // 2
    return f"""
dimensionality 3
courant {courant}
border_size {bs}
h {h} {h} {h}
steps {steps}
body 0 elastic isotropic sizes {n} {n} {half} start 0 0 0
body 1 elastic isotropic sizes {n} {n} {half} start 0 0 {half}
material body 0 isotropic 1 2 0.8
material body 1 isotropic 3 5 1.5
initial quantity PRESSURE 1 sphere 0.3 0.5 0.5 0.35
border 0 2 infinite Sxz const 0 Syz const 0 Szz const 0
border 1 2 infinite Sxz const 0 Syz const 0 Szz sin 1.0 4.0
"""


def ortho3d_rotated_plies(n=14, steps=4):
    # BASELINE config 4 with the plies turned +-45 degrees about y (the stacking axis): rotated orthotropic materials
    # in two glued bodies, one of them with a Maxwell relaxation time
    h = _h(n)
    half = n // 2
    ply = "1580 10.30e9 6.96e9 6.96e9 23.25e9 6.96e9 10.30e9 5.01e9 1.67e9 5.01e9"
    return f"""
dimensionality 3
courant 0.9
border_size 2
h {h} {h} {h}
steps {steps}
body 0 elastic orthotropic sizes {n} {half} {n} start 0 0 0
body 1 elastic orthotropic sizes {n} {half} {n} start 0 {half} 0
material body 0 orthotropic {ply} angles 0 0.7853981633974483 0
material body 1 orthotropic {ply} angles 0 -0.7853981633974483 0
initial quantity PRESSURE 1 sphere 0.35 0.5 0.5 0.5
border 1 1 sphere 0.3 0.5 1.0 0.5 Vy sin 1.0 5.0
"""


def elastic2d_pwave(n=40, steps=8, courant=0.9, bs=2):
    # BASELINE config 1 (200x200, 20 steps) at reduced size
    h = repr(4.0 / (n - 1))
    return f"""
dimensionality 2
courant {courant}
border_size {bs}
h {h} {h}
steps {steps}
body 0 elastic isotropic sizes {n} {n} start 0 0
material default isotropic 4 2 1
initial wave P_FORWARD 0 PRESSURE 1 box 1 -10 -10 2 10 10
"""


def elastic2d_full():
    # BASELINE config 1 exactly (checksum recorded in BASELINE.md / SURVEY.md §8c)
    return elastic2d_pwave(200, 20)


def elastic2d_courant45(n=30, steps=4):
    # the regime of test/sequence/TestEngine.cpp:91-136: border size 5, Courant 4.5, S-wave packet
    h = _h(n, 3.0)
    return f"""
dimensionality 2
courant 4.5
border_size 5
h {h} {h}
steps {steps}
body 0 elastic isotropic sizes {n} {n} start 0 0
material default isotropic 4 2 0.5
initial wave S1_FORWARD 0 Vy 1 box 0.6 -10 -10 1.4 10 10
"""


def elastic2d_ortho(n=24, steps=5):
    h = _h(n)
    return f"""
dimensionality 2
courant 0.8
border_size 2
h {h} {h}
steps {steps}
body 0 elastic orthotropic sizes {n} {n} start 0 0
material default orthotropic 3 360 70 70 180 70 90 10 10 25
initial quantity PRESSURE 1 sphere 0.3 0.5 0.5 0
border 0 0 infinite Sxx const 0 Sxy const 0
"""


def acoustic2d_border1(n=24, steps=6):
    # first-order scheme (border size 1) + partial-area border condition
    h = _h(n)
    return f"""
dimensionality 2
courant 0.7
border_size 1
h {h} {h}
steps {steps}
body 0 acoustic isotropic sizes {n} {n} start 0 0
material default isotropic 2 3 0
material area sphere 0.25 0.6 0.5 0 isotropic 1 8 0
initial quantity PRESSURE 1 sphere 0.2 0.4 0.5 0
border 0 1 box 0.2 -10 -10 0.8 10 10 Vy const 0.5
border 0 0 infinite PRESSURE sin 2.0 30.0
"""


def elastic1d(n=50, steps=10):
    h = _h(n)
    return f"""
dimensionality 1
courant 0.6
border_size 2
h {h}
steps {steps}
body 0 elastic isotropic sizes {n} start 0
material default isotropic 4 2 1
material area box 0.5 -10 -10 10 10 10 isotropic 1 3 2
initial wave P_FORWARD 0 Vx 1 box 0.1 -10 -10 0.3 10 10
border 0 0 infinite Vx const 0
"""


def acoustic1d(n=40, steps=10):
    h = _h(n)
    return f"""
dimensionality 1
courant 1.0
border_size 2
h {h}
steps {steps}
body 0 acoustic isotropic sizes {n} start 0
material default isotropic 1 4 0
initial quantity PRESSURE 1 box 0.3 -10 -10 0.6 10 10
"""


def maxwell3d(n=12, steps=4):
    h = _h(n)
    return f"""
dimensionality 3
courant 0.9
border_size 2
h {h} {h} {h}
steps {steps}
body 0 elastic isotropic sizes {n} {n} {n} start 0 0 0 ode maxwell
material default isotropic 4 2 1 tau0 0.05
material area box -10 -10 0.5 10 10 10 isotropic 2 3 1 tau0 0.2
initial quantity PRESSURE 1 sphere 0.3 0.5 0.5 0.5
"""


def adhesion2d(steps=12):
    # test/sequence/TestEngine.cpp:27-87 at reduced size: two glued bodies == one body
    two = f"""
dimensionality 2
courant 0.9
border_size 2
h 0.5 0.25
steps {steps}
body 0 elastic isotropic sizes 11 13 start 0 0
body 1 elastic isotropic sizes 11 13 start 0 13
material default isotropic 4 2 0.5
initial wave P_FORWARD 1 PRESSURE 1 box -10 0.4 -10 10 1.6 10
"""
    one = f"""
dimensionality 2
courant 0.9
border_size 2
h 0.5 0.25
steps {steps}
body 0 elastic isotropic sizes 11 26 start 0 0
material default isotropic 4 2 0.5
initial wave P_FORWARD 1 PRESSURE 1 box -10 0.4 -10 10 1.6 10
"""
    return two, one


def elastic2d_bs3(n=26, steps=6, courant=1.5):
    # the border size of the reference's engine tests (test/sequence/TestEngine.cpp:105,168,247) with feet in the
    # second cell: two materials, a border condition on each direction
    h = _h(n)
    return f"""
dimensionality 2
courant {courant}
border_size 3
h {h} {h}
steps {steps}
body 0 elastic isotropic sizes {n} {n + 3} start 0 0
material default isotropic 4 2 1
material area box -10 0.5 -10 10 10 10 isotropic 1 3 0.5
initial quantity PRESSURE 1 sphere 0.25 0.5 0.5 0
border 0 0 infinite Sxx const 0 Sxy const 0
border 0 1 box 0.2 -10 -10 0.8 10 10 Vy const 0.5
"""


SCENARIOS = {
    "elastic3d_iso": elastic3d_iso(),
    "elastic3d_iso_bs1": elastic3d_iso(14, 4, bs=1, courant=0.8),
    "acoustic3d_free": acoustic3d_free(),
    "elastic3d_layers": elastic3d_layers(),
    "elastic3d_ortho": elastic3d_ortho(),
    "elastic3d_ortho_rotated": elastic3d_ortho_rotated(),
    "ortho3d_contact": ortho3d_contact(),
    "ortho3d_rotated_plies": ortho3d_rotated_plies(),
    "elastic3d_contact_z": elastic3d_contact_z(),
    "elastic2d_pwave": elastic2d_pwave(),
    "elastic2d_courant45": elastic2d_courant45(),
    "elastic2d_ortho": elastic2d_ortho(),
    "acoustic2d_border1": acoustic2d_border1(),
    "elastic1d": elastic1d(),
    "acoustic1d": acoustic1d(),
    "maxwell3d": maxwell3d(),
    "adhesion2d_two": adhesion2d()[0],
    "adhesion2d_one": adhesion2d()[1],
    # Courant number 1, the value the reference launcher uses almost everywhere (src/launcher/main.cpp:82,197,272,395,
    # 438,477,580): the fastest wave's foot lies exactly one cell away (foot cell 1)
    "elastic3d_layers_courant1": elastic3d_layers(16, 4, courant=1.0),
    "acoustic3d_courant1": acoustic3d_free(16, 4, courant=1.0),
    "elastic2d_courant1": elastic2d_pwave(32, 6, courant=1.0),
    "ortho3d_contact_courant1": ortho3d_contact(12, 4, courant=1.0),
    "elastic3d_ortho_rotated_courant1": elastic3d_ortho_rotated(12, 3, courant=1.0),
    "elastic3d_layers_courant17": elastic3d_layers(14, 3, courant=1.7),
    # border size 3 (test/sequence/TestEngine.cpp:105,168,247)
    "elastic3d_iso_bs3": elastic3d_iso(14, 3, bs=3, courant=0.9),
    "elastic3d_layers_bs3_courant25": elastic3d_layers(14, 3, courant=2.5, bs=3),
    "acoustic3d_bs3_courant1": acoustic3d_free(14, 3, courant=1.0, bs=3),
    "elastic2d_bs3_courant15": elastic2d_bs3(),
}
