"""Helpers of the simplex-path tests: box meshes from the product's mesher, the oracle's view of a body."""
import ctypes
import os

import numpy as np

import oracle_host as oh

ip = ctypes.POINTER(ctypes.c_int)
dp = ctypes.POINTER(ctypes.c_double)


class GcmoTri(ctypes.Structure):
    _fields_ = [("nV", ctypes.c_int), ("nC", ctypes.c_int), ("xyz", dp), ("cell_v", ip), ("cell_n", ip),
                ("cell_grid", ip), ("inc_off", ip), ("inc_cell", ip), ("grid_id", ctypes.c_int),
                ("local_of", ip), ("global_of", ip), ("n_local", ctypes.c_int)]


def _i(a):
    return a.ctypes.data_as(ip)


def _d(a):
    return a.ctypes.data_as(dp)


class Mesh:
    """flat triangulation arrays (see gcm_b200/host/simplex_mesh.cpp)"""

    def __init__(self, host_lib, nx, ny, nz, origin=(0.0, 0.0, 0.0), h=1.0, jitter=0.0, seed=1, void_box=None, grid_id=0):
        f = host_lib.gcmb_host_simplex_box_mesh
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_int] * 3 + [dp, ctypes.c_double, ctypes.c_double, ctypes.c_uint, dp, ctypes.c_int, ip,
                                           dp, ip, ip, ip, ip, ip]
        o = np.array(origin, dtype=np.float64)
        vb = None if void_box is None else np.array(void_box, dtype=np.float64)
        sizes = np.zeros(3, dtype=np.int32)
        f(nx, ny, nz, _d(o), h, jitter, seed, None if vb is None else _d(vb), grid_id, _i(sizes), None, None, None, None, None, None)
        self.nV, self.nC, ninc = (int(s) for s in sizes)
        self.xyz = np.zeros((self.nV, 3))
        self.cell_v = np.zeros((self.nC, 4), dtype=np.int32)
        self.cell_n = np.zeros((self.nC, 4), dtype=np.int32)
        self.cell_grid = np.zeros(self.nC, dtype=np.int32)
        self.inc_off = np.zeros(self.nV + 1, dtype=np.int32)
        self.inc_cell = np.zeros(ninc, dtype=np.int32)
        f(nx, ny, nz, _d(o), h, jitter, seed, None if vb is None else _d(vb), grid_id, _i(sizes), _d(self.xyz),
          _i(self.cell_v), _i(self.cell_n), _i(self.cell_grid), _i(self.inc_off), _i(self.inc_cell))
        self.grid_id = grid_id
        self._index_body()

    @classmethod
    def from_arrays(cls, tri, grid_id):
        """a body's view of a triangulation given as the dict SimplexHostEngine.triangulation() returns"""
        m = cls.__new__(cls)
        for k, v in tri.items():
            setattr(m, k, v)
        m.nV, m.nC = len(m.xyz), len(m.cell_v)
        m.grid_id = grid_id
        m._index_body()
        return m

    def _index_body(self):
        used = np.zeros(self.nV, dtype=bool)
        used[self.cell_v[self.cell_grid == self.grid_id].ravel()] = True
        self.global_of = np.nonzero(used)[0].astype(np.int32)
        self.local_of = np.full(self.nV, -1, dtype=np.int32)
        self.local_of[self.global_of] = np.arange(len(self.global_of), dtype=np.int32)
        self.n_local = len(self.global_of)

    def retag(self, grid_of_centroid):
        """give every non-empty cell the body id grid_of_centroid(centroid) (topology does not depend on it)"""
        cent = self.xyz[self.cell_v].mean(axis=1)
        for c in range(self.nC):
            if self.cell_grid[c] >= 0:
                self.cell_grid[c] = grid_of_centroid(cent[c])
        self._index_body()

    def view(self, grid_id):
        """the same triangulation seen by another body"""
        import copy
        m = copy.copy(self)
        m.grid_id = grid_id
        m._index_body()
        return m

    def incident_grids(self):
        """Triangulation::incidentGridsIds of every global vertex; hull vertices see EmptySpace (-1)"""
        out = [set() for _ in range(self.nV)]
        for c in range(self.nC):
            for k in range(4):
                out[self.cell_v[c, k]].add(int(self.cell_grid[c]))
                if self.cell_n[c, k] < 0:
                    for j in range(4):
                        if j != k:
                            out[self.cell_v[c, j]].add(-1)
        return out

    def oracle_view(self):
        t = GcmoTri(self.nV, self.nC, _d(self.xyz), _i(self.cell_v), _i(self.cell_n), _i(self.cell_grid),
                    _i(self.inc_off), _i(self.inc_cell), self.grid_id, _i(self.local_of), _i(self.global_of), self.n_local)
        return t

    def local_xyz(self):
        return self.xyz[self.global_of]


def oracle():
    L = oh.lib()
    tp = ctypes.POINTER(GcmoTri)
    L.gcmo_simplex_locate.argtypes = [tp, ctypes.c_int, dp, ip]
    L.gcmo_simplex_border_state.argtypes = [tp, ctypes.c_int]
    L.gcmo_simplex_normal.argtypes = [tp, ctypes.c_int, ctypes.c_int, dp]
    L.gcmo_simplex_neighbors.argtypes = [tp, ctypes.c_int, ip, ctypes.c_int]
    L.gcmo_simplex_gradient.argtypes = [tp, ctypes.c_int, dp, dp]
    L.gcmo_simplex_hybrid_interpolate.argtypes = [tp, ctypes.c_int, dp, dp, ip, ctypes.c_int, dp, ip]
    L.gcmo_simplex_hybrid_interpolate.restype = ctypes.c_double
    L.gcmo_barycentric4.argtypes = [dp] * 6
    L.gcmo_oriented_volume.argtypes = [dp] * 4
    L.gcmo_oriented_volume.restype = ctypes.c_double
    L.gcmo_simplex_stage.argtypes = [tp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, dp, dp, dp, dp,
                                     ctypes.c_int, ip, dp, ip, ctypes.c_int, ip, dp, dp, dp]
    L.gcmo_simplex_plain_border.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, dp, ip, ip, dp, dp]
    L.gcmo_sx_begin.argtypes = [tp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, dp, dp, dp, dp, dp, dp, ctypes.c_int]
    L.gcmo_sx_begin.restype = ctypes.c_void_p
    L.gcmo_sx_nodes.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.gcmo_sx_nodes.restype = None
    L.gcmo_sx_border_correct.argtypes = [ctypes.c_void_p, ctypes.c_int, ip, dp, ip, ctypes.c_int, ip, dp]
    L.gcmo_sx_border_correct.restype = None
    L.gcmo_sx_contact_correct.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ip, ip, dp]
    L.gcmo_sx_contact_correct.restype = None
    L.gcmo_sx_end.argtypes = [ctypes.c_void_p]
    L.gcmo_simplex_contact_normal.argtypes = [tp, ctypes.c_int, ctypes.c_int, dp]
    L.gcmo_simplex_plain_contact.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, ip, dp, dp, dp]
    L.gcmo_simplex_plain_contact.restype = None
    L.gcmo_simplex_heights.argtypes = [tp, dp]
    L.gcmo_simplex_heights.restype = None
    return L


def directions(n=16):
    """the 16 x 16 directions x 9 lengths protocol of src/test/sequence/TestLineWalkSearch3D.cpp:120-154"""
    out = []
    for a in range(n):
        for b in range(n):
            phi, teta = 2 * np.pi * a / n, np.pi * (b + 0.5) / n
            out.append((np.sin(teta) * np.cos(phi), np.sin(teta) * np.sin(phi), np.cos(teta)))
    return np.array(out)


def oracle_locate_all(L, mesh, vertices, shifts):
    t = mesh.oracle_view()
    out = np.zeros((len(vertices), 5), dtype=np.int32)
    errs = 0
    for i, (v, s) in enumerate(zip(vertices, shifts)):
        sv = np.ascontiguousarray(s, dtype=np.float64)
        errs += L.gcmo_simplex_locate(ctypes.byref(t), int(v), _d(sv), _i(out[i]))
    return out, errs


class SimplexBody:
    """gcmb_simplex_* through ctypes"""

    def __init__(self, lib, ctx, mesh, model):
        from gcm_b200 import capi
        self.lib, self.mesh, self.model = lib, mesh, model
        self.handle = capi.vp()
        lib.check(lib.c.gcmb_simplex_body_create(ctx.handle, model, mesh.nV, mesh.nC, _d(mesh.xyz), _i(mesh.cell_v),
                                                 _i(mesh.cell_n), _i(mesh.cell_grid), _i(mesh.inc_off), _i(mesh.inc_cell),
                                                 mesh.grid_id, ctypes.byref(self.handle)))
        n, M, nb = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        lib.check(lib.c.gcmb_simplex_info(self.handle, ctypes.byref(n), ctypes.byref(M), ctypes.byref(nb)))
        self.n, self.M, self.n_border_vertices = n.value, M.value, nb.value

    def close(self):
        if self.handle:
            self.lib.c.gcmb_simplex_body_destroy(self.handle)
            self.handle = None

    def vertices(self):
        g = np.zeros(self.n, dtype=np.int32)
        st = np.zeros(self.n, dtype=np.uint8)
        bn = np.zeros((self.n, 3))
        cn = np.zeros((self.n, 3))
        self.lib.check(self.lib.c.gcmb_simplex_vertices(self.handle, _i(g), st.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), _d(bn), _d(cn)))
        return g, st, bn, cn

    def locate(self, vertices, shifts):
        v = np.ascontiguousarray(vertices, dtype=np.int32)
        s = np.ascontiguousarray(shifts, dtype=np.float64)
        out = np.zeros((len(v), 5), dtype=np.int32)
        self.lib.check(self.lib.c.gcmb_simplex_locate(self.handle, len(v), _i(v), _d(s), _i(out)))
        return out

    def errors(self):
        c = ctypes.c_int()
        self.lib.check(self.lib.c.gcmb_simplex_errors(self.handle, ctypes.byref(c)))
        return c.value

    def set_material(self, U, U1, L, basis):
        U, U1, L, basis = (np.ascontiguousarray(a, dtype=np.float64) for a in (U, U1, L, basis))
        self.lib.check(self.lib.c.gcmb_simplex_set_material(self.handle, _d(U), _d(U1), _d(L), _d(basis)))

    def upload(self, pde):
        p = np.ascontiguousarray(pde, dtype=np.float64)
        self.lib.check(self.lib.c.gcmb_simplex_upload_state(self.handle, _d(p)))

    def download(self):
        out = np.zeros((self.n, self.M))
        self.lib.check(self.lib.c.gcmb_simplex_download_state(self.handle, _d(out)))
        return out

    def border_set(self, types, nodes, normals, cond_of_node):
        t = np.ascontiguousarray(types, dtype=np.int32)
        nd = np.ascontiguousarray(nodes, dtype=np.int32)
        nr = np.ascontiguousarray(normals, dtype=np.float64)
        cd = np.ascontiguousarray(cond_of_node, dtype=np.int32)
        self.lib.check(self.lib.c.gcmb_simplex_border_set(self.handle, len(t), _i(t), len(nd), _i(nd), _d(nr), _i(cd)))

    def plain_border(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        self.lib.check(self.lib.c.gcmb_simplex_plain_border(self.handle, _d(v)))

    def stage(self, s, tau, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        self.lib.check(self.lib.c.gcmb_simplex_stage(self.handle, s, tau, _d(v)))

    def set_gcm_type(self, gcm_type):
        self.lib.check(self.lib.c.gcmb_simplex_set_gcm_type(self.handle, gcm_type))

    def before_stage(self, s, tau):
        self.lib.check(self.lib.c.gcmb_simplex_before_stage(self.handle, s, tau))

    def border_contact_stage(self):
        self.lib.check(self.lib.c.gcmb_simplex_border_contact_stage(self.handle))

    def border_correct(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        self.lib.check(self.lib.c.gcmb_simplex_border_correct(self.handle, _d(v)))

    def inner_stage(self):
        self.lib.check(self.lib.c.gcmb_simplex_inner_stage(self.handle))

    def after_stage(self):
        self.lib.check(self.lib.c.gcmb_simplex_after_stage(self.handle))

    def contact_normals(self, neighbor):
        out = np.zeros((self.n, 3))
        self.lib.check(self.lib.c.gcmb_simplex_contact_normals(self.handle, neighbor, _d(out)))
        return out

    def gradient(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        g = np.zeros((self.n, 3, self.M))
        self.lib.check(self.lib.c.gcmb_simplex_gradient(self.handle, _d(v), _d(g)))
        return g


class SimplexContact:
    def __init__(self, lib, a, b, node_a, node_b, normals):
        from gcm_b200 import capi
        self.lib = lib
        self.handle = capi.vp()
        na = np.ascontiguousarray(node_a, dtype=np.int32)
        nb = np.ascontiguousarray(node_b, dtype=np.int32)
        nr = np.ascontiguousarray(normals, dtype=np.float64)
        lib.check(lib.c.gcmb_simplex_contact_create(a.handle, b.handle, len(na), _i(na), _i(nb), _d(nr), ctypes.byref(self.handle)))

    def close(self):
        if self.handle:
            self.lib.c.gcmb_simplex_contact_destroy(self.handle)
            self.handle = None

    def plain(self):
        self.lib.check(self.lib.c.gcmb_simplex_contact_plain(self.handle))

    def correct(self):
        self.lib.check(self.lib.c.gcmb_simplex_contact_correct(self.handle))


def protocol_queries(mesh, n_dirs=16, lengths=9, scale=1.0, vertices=None):
    """every vertex x n_dirs^2 directions x `lengths` lengths (TestLineWalkSearch3D.cpp:120-154)"""
    dirs = directions(n_dirs)
    lens = scale * (0.15 + 0.35 * np.arange(lengths))
    vs = np.arange(mesh.n_local) if vertices is None else np.asarray(vertices)
    v = np.repeat(vs, len(dirs) * len(lens)).astype(np.int32)
    sh = (dirs[None, :, None, :] * lens[None, None, :, None]).reshape(1, -1, 3)
    sh = np.broadcast_to(sh, (len(vs), sh.shape[1], 3)).reshape(-1, 3).copy()
    return v, sh


def write_flat_dump(tri, path):
    """the mesh file of oracle/_ref/gcm_ref_simplex (oracle/shim/CGAL/flat_triangulation_3.h): "nV nC", points,
    then per cell 4 vertices, 4 neighbours, grid id"""
    with open(path, "w") as f:
        f.write("%d %d\n" % (len(tri["xyz"]), len(tri["cell_v"])))
        for p in tri["xyz"]:
            f.write("%.17e %.17e %.17e\n" % tuple(p))
        for v, n, g in zip(tri["cell_v"], tri["cell_n"], tri["cell_grid"]):
            f.write("%d %d %d %d %d %d %d %d %d\n" % (tuple(v) + tuple(n) + (g,)))


def run_reference_simplex(task_text, tri, workdir, bodies, M):
    """the UNMODIFIED reference simplex engine (with the CGAL stand-in) on the given triangulation: returns
    {body: (coords [n,3], pde [n,M])}, meta"""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(oh.__file__)), "_ref", "gcm_ref_simplex")
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "-j8", "-C", os.path.dirname(exe) + "/..", "ref_simplex"], check=True)
    os.makedirs(os.path.join(workdir, "snapshots"), exist_ok=True)
    mesh = os.path.join(workdir, "mesh.flat")
    write_flat_dump(tri, mesh)
    task = os.path.join(workdir, "task.txt")
    open(task, "w").write(task_text + "\nsimplex_flat %s\n" % mesh)
    r = subprocess.run([exe, task, os.path.join(workdir, "out")], cwd=workdir, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcm_ref_simplex failed: " + r.stdout[-2000:] + r.stderr[-2000:])
    meta = {}
    for line in open(os.path.join(workdir, "out.meta")):
        w = line.split()
        if w[0] == "body":
            meta[("body", int(w[1]))] = {w[i]: float(w[i + 1]) for i in range(2, len(w), 2)}
        else:
            meta[w[0]] = float(w[1])
    meta["cell_grid"] = np.loadtxt(os.path.join(workdir, "out.cells"), dtype=np.int64).astype(np.int32)
    out = {}
    for b in bodies:
        raw = np.fromfile(os.path.join(workdir, "out.body%d.f64" % b)).reshape(-1, 3 + M)
        out[b] = (raw[:, :3].copy(), raw[:, 3:].copy())
    return out, meta


def write_inm_all_cells(tri, cell_grid, path):
    """INM-format mesh file that keeps EVERY cell (empty ones with material -1) in the triangulation's order, so that
    a body built from it sees the same cells, neighbours and incident lists as the flat dump"""
    with open(path, "w") as f:
        f.write("%d\n" % len(tri["xyz"]))
        for p in tri["xyz"]:
            f.write("%.17e %.17e %.17e\n" % tuple(p))
        f.write("%d\n" % len(tri["cell_v"]))
        for v, g in zip(tri["cell_v"], cell_grid):
            f.write("%d %d %d %d %d\n" % (v[0] + 1, v[1] + 1, v[2] + 1, v[3] + 1, g))
        f.write("0\n")
