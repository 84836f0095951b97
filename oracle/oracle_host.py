"""TEST INFRASTRUCTURE ONLY — numpy/ctypes host around oracle/cubic_oracle.c.

Restates the *setup and time loop* of the reference's cubic engine so that the C restatement of the
hot path (cubic_oracle.c) can be run on a whole task and pinned against the unmodified reference
(oracle/_ref/gcm_ref).  The product (gcm_b200/) never imports this module.

Reference lines followed (relative to /root/reference/src/libgcm):
  task parsing        our own plain-text task format, same one oracle/ref_driver.cpp reads
  materials           util/task/MaterialsCondition.hpp:23-36,61-96
  initial conditions  util/task/InitialCondition.hpp:23-88
  areas               util/math/Area.hpp:23-123
  contacts            engine/cubic/Engine.cpp:40-87, util/math/AABB.hpp
  time loop           engine/AbstractEngine.cpp:18-46, engine/cubic/Engine.cpp:92-140
"""
import ctypes
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

QUANTITY_ORDER = ["VELOCITY", "FORCE", "Vx", "Vy", "Vz", "Sxx", "Sxy", "Sxz", "Syy", "Syz", "Szz",
                  "RHO", "PRESSURE", "DAMAGE_MEASURE"]  # util/Enum.hpp:28-50 (std::map order)
Q_PRESSURE_TRACE = -1


def build_oracle():
    """Compile oracle/_ref/libgcm_oracle.so (and gcm_ref when the reference tree is present)."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-j8", "-C", HERE, "ref", "ref_simplex"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "_ref", "libgcm_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = ctypes.CDLL(path)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        bp = ctypes.POINTER(ctypes.c_uint8)
        L.gcmo_pde_size.restype = ctypes.c_int
        L.gcmo_all_nodes.restype = ctypes.c_size_t
        L.gcmo_all_nodes.argtypes = [ctypes.c_int, ip, ctypes.c_int]
        L.gcmo_elastic_isotropic.argtypes = [ctypes.c_int] + [ctypes.c_double] * 3 + [dp] * 3
        L.gcmo_elastic_orthotropic.argtypes = [ctypes.c_int, ctypes.c_double, dp, dp, dp, dp]
        L.gcmo_elastic_orthotropic_rotated.argtypes = [ctypes.c_double, dp, dp, dp, dp, dp]
        L.gcmo_elastic_orthotropic_rotated.restype = ctypes.c_int
        L.gcmo_acoustic.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, dp, dp, dp]
        L.gcmo_minmax_interpolate.argtypes = [ctypes.c_int, ctypes.c_int, dp, ctypes.c_double, dp]
        L.gcmo_interpolate.argtypes = [ctypes.c_int, ctypes.c_int, dp, ctypes.c_double, dp]
        L.gcmo_stage.argtypes = [ctypes.c_int, ctypes.c_int, ip, ctypes.c_int, dp, ctypes.c_int,
                                 ctypes.c_double, ctypes.c_int, dp, dp, dp, bp, dp, dp]
        L.gcmo_border_apply.argtypes = [ctypes.c_int, ctypes.c_int, ip, ctypes.c_int, ctypes.c_int,
                                        bp, bp, ctypes.c_int, ip, dp, dp]
        L.gcmo_contact_copy.argtypes = [ctypes.c_int, ctypes.c_int, ip, ip, ctypes.c_int, ip, ip, ip, dp, dp]
        L.gcmo_ode_maxwell.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, ctypes.c_int, dp, bp, dp]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _bp(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))


# --------------------------------------------------------------------------------------------
# task files
# --------------------------------------------------------------------------------------------
class Tokens:
    def __init__(self, words):
        self.t, self.pos = words, 0

    def done(self):
        return self.pos >= len(self.t)

    def next(self):
        self.pos += 1
        return self.t[self.pos - 1]

    def peek(self):
        return None if self.done() else self.t[self.pos]

    def num(self):
        return float(self.next())

    def inum(self):
        return int(self.next())


def parse_area(tk):
    kind = tk.next()
    if kind == "infinite":
        return ("infinite",)
    if kind == "box":
        return ("box", [tk.num() for _ in range(3)], [tk.num() for _ in range(3)])
    if kind == "sphere":
        return ("sphere", tk.num(), [tk.num() for _ in range(3)])
    if kind == "cylinder":
        return ("cylinder", tk.num(), [tk.num() for _ in range(3)], [tk.num() for _ in range(3)])
    raise ValueError("unknown area " + kind)


def parse_material(tk):
    kind = tk.next()
    if kind == "isotropic":
        m = {"kind": kind, "rho": tk.num(), "lambda": tk.num(), "mu": tk.num(), "tau0": 0.0}
    elif kind == "orthotropic":
        m = {"kind": kind, "rho": tk.num(), "c": [tk.num() for _ in range(9)], "tau0": 0.0, "angles": [0.0, 0.0, 0.0]}
        if tk.peek() == "angles":
            tk.next()
            m["angles"] = [tk.num() for _ in range(3)]
    else:
        raise ValueError("unknown material " + kind)
    if tk.peek() == "tau0":
        tk.next()
        m["tau0"] = tk.num()
    return m


def parse_time_dependency(tk):
    kind = tk.next()
    if kind == "const":
        c = tk.num()
        return lambda t: c
    if kind == "sin":
        amp, omega = tk.num(), tk.num()
        return lambda t: amp * math.sin(omega * t)
    if kind == "gauss":
        amp, t0, tau = tk.num(), tk.num(), tk.num()
        return lambda t: amp * math.exp(-(t - t0) * (t - t0) / (2 * tau * tau))
    if kind == "until":
        t1, value = tk.num(), tk.num()
        return lambda t: value if t < t1 else 0.0
    raise ValueError("unknown time dependency " + kind)


def parse_task(text):
    task = {"bodies": {}, "mat_default": None, "mat_areas": [], "mat_bodies": {}, "initial": [],
            "borders": {}, "detector": None, "required_time": None, "steps": 0}
    for line in text.splitlines():
        line = line.split("#")[0]
        words = line.split()
        if not words:
            continue
        tk = Tokens(words)
        key = tk.next()
        if key == "dimensionality":
            task["D"] = tk.inum()
        elif key == "courant":
            task["courant"] = tk.num()
        elif key == "border_size":
            task["bs"] = tk.inum()
        elif key == "h":
            task["h"] = [float(w) for w in words[1:]]
        elif key == "steps":
            task["steps"] = tk.inum()
        elif key == "required_time":
            task["steps"] = 0
            task["required_time"] = tk.num()
        elif key == "body":
            bid = tk.inum()
            body = {"model": tk.next(), "material": tk.next(), "odes": []}
            D = task["D"]
            while not tk.done():
                sub = tk.next()
                if sub == "sizes":
                    body["sizes"] = [tk.inum() for _ in range(D)]
                elif sub == "start":
                    body["start"] = [tk.inum() for _ in range(D)]
                elif sub == "ode":
                    body["odes"].append(tk.next())
            task["bodies"][bid] = body
        elif key == "material":
            how = tk.next()
            if how == "default":
                task["mat_default"] = parse_material(tk)
            elif how == "area":
                area = parse_area(tk)
                task["mat_areas"].append((area, parse_material(tk)))
            elif how == "body":
                bid = tk.inum()
                task["mat_bodies"][bid] = parse_material(tk)
        elif key == "initial":
            what = tk.next()
            if what == "quantity":
                q, v = tk.next(), tk.num()
                task["initial"].append(("quantity", q, v, parse_area(tk)))
            elif what == "wave":
                w, d, q, v = tk.next(), tk.inum(), tk.next(), tk.num()
                task["initial"].append(("wave", w, d, q, v, parse_area(tk)))
        elif key == "border":
            bid, d = tk.inum(), tk.inum()
            area = parse_area(tk)
            vals = {}
            while not tk.done():
                q = tk.next()
                vals[q] = parse_time_dependency(tk)
            task["borders"].setdefault(bid, []).append((d, area, vals))
        elif key == "detector":
            task["detector"] = (tk.inum(), tk.next(), parse_area(tk))
        else:
            raise ValueError("unknown key " + key)
    return task


# --------------------------------------------------------------------------------------------
# areas (util/math/Area.hpp) evaluated on arrays of coordinates [..., 3]
# --------------------------------------------------------------------------------------------
def area_contains(area, x, y, z):
    kind = area[0]
    if kind == "infinite":
        return np.ones(np.broadcast(x, y, z).shape, dtype=bool)
    if kind == "box":
        lo, hi = area[1], area[2]
        ok = np.ones(np.broadcast(x, y, z).shape, dtype=bool)
        for c, a, b in ((x, lo[0], hi[0]), (y, lo[1], hi[1]), (z, lo[2], hi[2])):
            ok &= ~((c <= a) | (c >= b))
        return ok
    if kind == "sphere":
        r, c = area[1], area[2]
        dx, dy, dz = x - c[0], y - c[1], z - c[2]
        return np.sqrt(dx * dx + dy * dy + dz * dz) < r
    if kind == "cylinder":
        r, b, e = area[1], np.array(area[2]), np.array(area[3])
        ax = e - b
        ax = ax / math.sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2])
        pb = [x - b[0], y - b[1], z - b[2]]
        pe = [x - e[0], y - e[1], z - e[2]]
        d1 = pb[0] * ax[0] + pb[1] * ax[1] + pb[2] * ax[2]
        d2 = pe[0] * ax[0] + pe[1] * ax[1] + pe[2] * ax[2]
        between = ~(d1 * d2 >= 0)
        return between & ((pb[0] * pb[0] + pb[1] * pb[1] + pb[2] * pb[2]) - d1 * d1 < r * r)
    raise ValueError(kind)


def sym_index(D, i, j):
    if i > j:
        i, j = j, i
    return i * D - ((i - 1) * i) // 2 + j - i


def quantity_code(model, D, name):
    """component index, or Q_PRESSURE_TRACE (VelocitySigmaVariables.cpp / AcousticVariables.cpp maps)."""
    if name in ("Vx", "Vy", "Vz"):
        i = "xyz".index(name[1])
        assert i < D
        return i
    if model == "acoustic":
        assert name == "PRESSURE"
        return D
    if name == "PRESSURE":
        return Q_PRESSURE_TRACE
    i, j = "xyz".index(name[1]), "xyz".index(name[2])
    assert i < D and j < D
    return D + sym_index(D, i, j)


WAVE_COLUMNS = {  # rheology/models/Model.cpp:6-63
    ("elastic", "isotropic"): {"P_FORWARD": 0, "P_BACKWARD": 1, "S1_FORWARD": 2, "S1_BACKWARD": 3,
                               "S2_FORWARD": 4, "S2_BACKWARD": 5},
    ("elastic", "orthotropic", 3): {"P_FORWARD": 5, "P_BACKWARD": 4, "S1_FORWARD": 1, "S1_BACKWARD": 0,
                                    "S2_FORWARD": 3, "S2_BACKWARD": 2},
    ("elastic", "orthotropic", 2): {"P_FORWARD": 3, "P_BACKWARD": 2, "S1_FORWARD": 1, "S1_BACKWARD": 0},
    ("acoustic", "isotropic"): {"P_FORWARD": 0, "P_BACKWARD": 1},
}


def matrices_for(model, D, mat):
    L = lib()
    M = L.gcmo_pde_size(1 if model == "acoustic" else 0, D)
    U = np.zeros((D, M, M))
    U1 = np.zeros((D, M, M))
    Lm = np.zeros((D, M))
    if model == "acoustic":
        L.gcmo_acoustic(D, mat["rho"], mat["lambda"], _dp(U), _dp(U1), _dp(Lm))
    elif mat["kind"] == "isotropic":
        L.gcmo_elastic_isotropic(D, mat["rho"], mat["lambda"], mat["mu"], _dp(U), _dp(U1), _dp(Lm))
    else:
        c = np.array(mat["c"], dtype=np.float64)
        if any(a != 0 for a in mat.get("angles", [0, 0, 0])):   # ElasticModel3D.cpp:151 (3-D only, like the reference)
            if D != 3:
                raise ValueError("rotated orthotropic materials exist in 3-D only")
            angles = np.array(mat["angles"], dtype=np.float64)
            if L.gcmo_elastic_orthotropic_rotated(mat["rho"], _dp(c), _dp(angles), _dp(U), _dp(U1), _dp(Lm)) != 0:
                raise ValueError("rotated orthotropic material: no real eigen-system")
        else:
            L.gcmo_elastic_orthotropic(D, mat["rho"], _dp(c), _dp(U), _dp(U1), _dp(Lm))
    return U, U1, Lm


class Body:
    pass


class OracleEngine:
    """Restatement of cubic::Engine<D> construction + run on top of the C oracle."""

    def __init__(self, task):
        self.task = task
        D, bs = task["D"], task["bs"]
        self.D, self.bs = D, bs
        self.h = np.array(task["h"], dtype=np.float64)
        self.courant = task["courant"]
        self.bodies = {}
        for bid in sorted(task["bodies"]):
            self.bodies[bid] = self._make_body(bid, task["bodies"][bid])
        self._make_contacts()
        self.time = 0.0
        self.tau = self.estimate_time_step()
        if task["steps"] > 0:
            self.required_time = self.tau * task["steps"] * 1
        else:
            self.required_time = task["required_time"]
        self.steps_done = 0
        self.seismo = []

    # ---- setup ----
    def _conditions(self, bid, body):
        t = self.task
        if t["mat_bodies"]:
            return [(("infinite",), t["mat_bodies"][bid])]
        return [(("infinite",), t["mat_default"])] + list(t["mat_areas"])

    def _coords(self, b, full=False):
        """coordinates of real nodes as broadcastable arrays x,y,z (CubicGrid::coords)."""
        D = self.D
        axes = []
        for i in range(3):
            if i < D:
                c = (b.start[i] * self.h[i]) + np.arange(b.sizes[i], dtype=np.float64) * self.h[i]
            else:
                c = np.zeros(1)
            shape = [1, 1, 1]
            shape[i] = c.size
            axes.append(c.reshape(shape))
        return axes

    def _make_body(self, bid, spec):
        L = lib()
        D, bs = self.D, self.bs
        b = Body()
        b.id, b.model, b.mat_kind = bid, spec["model"], spec["material"]
        b.sizes = np.array(spec["sizes"], dtype=np.int32)
        b.start = np.array(spec["start"], dtype=np.int32)
        b.odes = spec["odes"]
        b.M = L.gcmo_pde_size(1 if b.model == "acoustic" else 0, D)
        b.full = tuple(int(s) + 2 * bs for s in b.sizes)
        b.real = tuple(slice(bs, bs + int(s)) for s in b.sizes)
        conds = self._conditions(bid, spec)
        b.materials = [m for _, m in conds]
        mats = [matrices_for(b.model, D, m) for _, m in conds]
        b.U = np.ascontiguousarray(np.stack([m[0] for m in mats]))
        b.U1 = np.ascontiguousarray(np.stack([m[1] for m in mats]))
        b.L = np.ascontiguousarray(np.stack([m[2] for m in mats]))
        b.max_eig = 0.0
        for m in mats:
            b.max_eig = max(b.max_eig, float(np.abs(m[2]).max()))
        x, y, z = self._coords(b)
        shape3 = tuple(int(s) for s in b.sizes) + (1,) * (3 - D)
        table = np.zeros(shape3, dtype=np.uint8)
        for i, (area, _) in enumerate(conds):
            table[np.broadcast_to(area_contains(area, x, y, z), shape3)] = i
        b.table = np.zeros(b.full, dtype=np.uint8)
        b.table[b.real] = table.reshape(tuple(int(s) for s in b.sizes))
        # initial conditions (InitialCondition.hpp)
        pde = np.zeros(shape3 + (b.M,))
        for ic in self.task["initial"]:
            vec = np.zeros(b.M)
            if ic[0] == "quantity":
                _, q, v, area = ic
                code = quantity_code(b.model, D, q)
                if code >= 0:
                    vec[code] = v
                else:
                    for i in range(D):
                        vec[D + sym_index(D, i, i)] = -v
            else:
                _, w, d, q, v, area = ic
                key = (b.model, b.mat_kind) if b.mat_kind == "isotropic" else (b.model, b.mat_kind, D)
                col = WAVE_COLUMNS[key][w]
                vec = b.U1[0][d][:, col].copy()
                code = quantity_code(b.model, D, q)
                if code >= 0:
                    cur = vec[code]
                else:
                    tr = 0.0
                    for i in range(D):
                        tr += vec[D + sym_index(D, i, i)]
                    cur = -tr / D
                vec = vec * (v / cur)
            mask = np.broadcast_to(area_contains(area, x, y, z), shape3)
            pde[mask] += vec
        b.pde = np.zeros(b.full + (b.M,))
        b.pde[b.real] = pde.reshape(tuple(int(s) for s in b.sizes) + (b.M,))
        b.pde_new = np.zeros_like(b.pde)
        # border conditions (BorderConditions.hpp:46-78)
        b.borders = []
        for (d, area, vals) in self.task["borders"].get(bid, []):
            names = sorted(vals, key=QUANTITY_ORDER.index)
            codes = np.array([quantity_code(b.model, D, n) for n in names], dtype=np.int32)
            masks = []
            for side in (0, 1):
                sl = [x, y, z]
                idx = 0 if side == 0 else int(b.sizes[d]) - 1
                face = []
                for i in range(3):
                    if i == d:
                        face.append(np.take(sl[i], [idx], axis=i))
                    else:
                        face.append(sl[i])
                shape_face = list(shape3)
                shape_face[d] = 1
                m = np.broadcast_to(area_contains(area, *face), tuple(shape_face))
                masks.append(np.ascontiguousarray(m.reshape(-1).astype(np.uint8)))
            b.borders.append((d, masks[0], masks[1], codes, [vals[n] for n in names]))
        return b

    def _make_contacts(self):
        D, bs = self.D, self.bs
        for b in self.bodies.values():
            b.contacts = []
        for b in self.bodies.values():
            for o in self.bodies.values():
                if o.id == b.id:
                    continue
                amin, amax = b.start, b.start + b.sizes - 1
                bmin, bmax = o.start, o.start + o.sizes - 1
                imin, imax = np.maximum(amin, bmin), np.minimum(amax, bmax)
                sz = imax - imin
                if np.all(sz >= 0):
                    raise ValueError("Bodies must not intersect")
                axis = 0
                for i in range(1, D):
                    if sz[i] < sz[axis]:
                        axis = i
                if sz[axis] != -1:
                    continue
                lo, hi = imin.copy(), imax.copy()
                if b.start[axis] > o.start[axis]:
                    lo[axis] -= bs
                else:
                    hi[axis] += bs
                extent = (hi - lo + 1).astype(np.int32)
                b.contacts.append((o.id, axis, (lo - b.start).astype(np.int32),
                                   (lo - o.start).astype(np.int32), extent))

    def estimate_time_step(self):
        max_eig = 0.0
        for b in self.bodies.values():
            if b.max_eig > max_eig:
                max_eig = b.max_eig
        return self.courant * float(self.h.min()) / max_eig

    # ---- time loop ----
    def next_time_step(self):
        L = lib()
        D, bs = self.D, self.bs
        for stage in range(D):
            for b in self.bodies.values():
                for (d, lm, rm, codes, fns) in b.borders:
                    if d != stage:
                        continue
                    vals = np.array([f(self.time) for f in fns], dtype=np.float64)
                    L.gcmo_border_apply(D, b.M, _ip(b.sizes), bs, d, _bp(lm), _bp(rm),
                                        len(codes), _ip(codes), _dp(vals), _dp(b.pde))
            for b in self.bodies.values():
                for (oid, axis, amin, bmin, extent) in b.contacts:
                    if axis != stage:
                        continue
                    o = self.bodies[oid]
                    L.gcmo_contact_copy(D, b.M, _ip(b.sizes), _ip(o.sizes), bs, _ip(amin), _ip(bmin),
                                        _ip(extent), _dp(b.pde), _dp(o.pde))
            for b in self.bodies.values():
                rc = L.gcmo_stage(D, b.M, _ip(b.sizes), bs, _dp(self.h), stage, self.tau,
                                  len(b.materials), _dp(b.U), _dp(b.U1), _dp(b.L), _bp(b.table),
                                  _dp(b.pde), _dp(b.pde_new))
                if rc:
                    raise RuntimeError("interpolation out of range (reference would throw), rc=%d" % rc)
                b.pde, b.pde_new = b.pde_new, b.pde
        for b in self.bodies.values():
            for ode in b.odes:
                decay = np.array([math.exp(-self.tau / m["tau0"]) for m in b.materials])
                L.gcmo_ode_maxwell(D, b.M, 1 if b.model == "acoustic" else 0, _ip(b.sizes), bs,
                                   _dp(decay), _bp(b.table), _dp(b.pde))

    def detector_value(self):
        """SliceSnapshotter.hpp:60-77: mean of one quantity over the right border of the last axis."""
        det = self.task["detector"]
        if det is None:
            return None
        bid, q, area = det
        b = self.bodies[bid]
        D = self.D
        d = D - 1
        x, y, z = self._coords(b)
        sl = [x, y, z]
        sl[d] = np.take(sl[d], [int(b.sizes[d]) - 1], axis=d)
        shape3 = [int(s) for s in b.sizes] + [1] * (3 - D)
        shape3[d] = 1
        mask = np.broadcast_to(area_contains(area, *sl), tuple(shape3)).reshape(-1)
        real = b.pde[b.real]
        face = np.take(real, [int(b.sizes[d]) - 1], axis=d).reshape(-1, b.M)
        code = quantity_code(b.model, D, q)
        if code >= 0:
            vals = face[:, code]
        else:
            tr = np.zeros(face.shape[0])
            for i in range(D):
                tr = tr + face[:, D + sym_index(D, i, i)]
            vals = -tr / D
        sel = vals[mask.astype(bool)]
        acc = 0.0
        for v in sel:  # std::accumulate order
            acc += float(v)
        return np.float32(acc / len(sel))

    def run(self):
        self.seismo.append((self.time, self.detector_value()))
        while self.time < self.required_time:
            self.tau = self.estimate_time_step()
            self.next_time_step()
            self.steps_done += 1
            self.time += self.tau
            self.seismo.append((self.time, self.detector_value()))
        return self

    def real_nodes(self, bid):
        b = self.bodies[bid]
        return np.ascontiguousarray(b.pde[b.real]).reshape(-1, b.M)


def run_task_text(text):
    return OracleEngine(parse_task(text)).run()


def run_reference(task_text, workdir, matrices=False, exe_name="gcm_ref"):
    """Run the UNMODIFIED reference (oracle/_ref/gcm_ref) on a task; returns dict body id -> array.
    exe_name="gcm_ref_gpu": the same unmodified engine with gcm_b200 plugged in as its backend (integration/)."""
    exe = os.path.join(HERE, "_ref", exe_name)
    os.makedirs(workdir, exist_ok=True)
    tf = os.path.join(workdir, "task.txt")
    with open(tf, "w") as f:
        f.write(task_text)
    prefix = os.path.join(workdir, "out")
    for sub in ("zaxis", "detector"):  # SliceSnapshotter writes here (Snapshotter.hpp:53-68)
        os.makedirs(os.path.join(workdir, "snapshots", sub), exist_ok=True)
    cmd = [exe, tf, prefix] + (["--matrices"] if matrices else [])
    subprocess.run(cmd, check=True, cwd=workdir)
    meta = {}
    out = {}
    with open(prefix + ".meta") as f:
        for line in f:
            w = line.split()
            if w[0] == "body":
                bid, M = int(w[1]), int(w[3])
                out[bid] = np.fromfile("%s.body%d.f64" % (prefix, bid)).reshape(-1, M)
                if matrices:
                    out[("mat", bid)] = np.fromfile("%s.mat%d.f64" % (prefix, bid))
            else:
                meta[w[0]] = float(w[1])
    out["meta"] = meta
    det = os.path.join(workdir, "snapshots", "detector")
    files = sorted(os.listdir(det)) if os.path.isdir(det) else []
    if files:  # the last file holds the whole history: columns time, value (float32 printed)
        out["detector"] = np.loadtxt(os.path.join(det, files[-1]), ndmin=2)
    return out
