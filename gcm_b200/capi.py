"""ctypes binding of the C ABI (include/gcm_b200.h) and of the host layer's C entry points.

`Library()` loads gcm_b200/libgcm_b200.so + libgcm_b200_host.so and fails loudly when they are missing:
there is no Python or CPU implementation behind these calls.  (The host-logic tests pass explicit paths
to load their stepping harness instead; the product never does.)
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

c_int_p = ctypes.POINTER(ctypes.c_int)
c_double_p = ctypes.POINTER(ctypes.c_double)
c_u8_p = ctypes.POINTER(ctypes.c_uint8)
c_ll_p = ctypes.POINTER(ctypes.c_longlong)
c_float_p = ctypes.POINTER(ctypes.c_float)
vp = ctypes.c_void_p

Q_PRESSURE_TRACE = -1

# every symbol include/gcm_b200.h declares: name -> (restype, argtypes)
C_ABI = {
    "gcmb_last_error": (ctypes.c_char_p, []),
    "gcmb_version": (ctypes.c_char_p, []),
    "gcmb_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]),
    "gcmb_destroy": (None, [vp]),
    "gcmb_set_stream": (ctypes.c_int, [vp, vp]),
    "gcmb_sync": (ctypes.c_int, [vp]),
    "gcmb_timer_start": (ctypes.c_int, [vp]),
    "gcmb_timer_stop": (ctypes.c_int, [vp, c_float_p]),
    "gcmb_profile_enable": (ctypes.c_int, [vp, ctypes.c_int]),
    "gcmb_profile_get": (ctypes.c_int, [vp, ctypes.c_int, c_double_p, c_ll_p]),
    "gcmb_launch_count": (ctypes.c_longlong, [vp]),
    "gcmb_device_bytes": (ctypes.c_size_t, [vp]),
    "gcmb_cubic_body_create": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, c_int_p, c_int_p, c_double_p,
                                              ctypes.c_int, ctypes.POINTER(vp)]),
    "gcmb_cubic_body_destroy": (None, [vp]),
    "gcmb_cubic_upload_state": (ctypes.c_int, [vp, vp, ctypes.c_int]),
    "gcmb_cubic_download_state": (ctypes.c_int, [vp, vp, ctypes.c_int]),
    "gcmb_cubic_set_materials": (ctypes.c_int, [vp, ctypes.c_int, c_double_p, c_double_p, c_double_p, c_u8_p]),
    "gcmb_cubic_assign_table_in_area": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, c_double_p]),
    "gcmb_cubic_add_vector_in_area": (ctypes.c_int, [vp, c_double_p, ctypes.c_int, c_double_p]),
    "gcmb_cubic_border_set": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, c_u8_p, c_u8_p, ctypes.c_int, c_int_p]),
    "gcmb_cubic_border_set_area": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                  c_double_p, ctypes.c_int, c_int_p]),
    "gcmb_cubic_border_apply": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, c_double_p]),
    "gcmb_cubic_contact_apply": (ctypes.c_int, [vp, vp, c_int_p, c_int_p, c_int_p]),
    "gcmb_cubic_stage": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_double]),
    "gcmb_cubic_ode_maxwell": (ctypes.c_int, [vp, c_double_p]),
    "gcmb_cubic_detector_set_area": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, c_double_p]),
    "gcmb_cubic_detector_set_mask": (ctypes.c_int, [vp, ctypes.c_int, c_u8_p]),
    "gcmb_cubic_seismo": (ctypes.c_int, [vp, c_double_p, c_ll_p, ctypes.c_int, c_double_p, ctypes.c_int]),
    "gcmb_comm_unique_id": (ctypes.c_int, [vp]),
    "gcmb_comm_init": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, vp]),
    "gcmb_cubic_halo_exchange": (ctypes.c_int, [vp]),
    "gcmb_cubic_halo_bytes": (ctypes.c_size_t, [vp]),
    "gcmb_cubic_halo_get": (ctypes.c_int, [vp, ctypes.c_int, vp]),
    "gcmb_cubic_halo_put": (ctypes.c_int, [vp, ctypes.c_int, vp]),
    "gcmb_comm_allreduce_sum": (ctypes.c_int, [vp, c_double_p, ctypes.c_int]),
    "gcmb_cubic_checksum": (ctypes.c_int, [vp, c_double_p]),
    "gcmb_cubic_download_tables": (ctypes.c_int, [vp, c_u8_p]),
    "gcmb_simplex_body_create": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p, c_int_p, c_int_p,
                                                c_int_p, c_int_p, c_int_p, ctypes.c_int, ctypes.POINTER(vp)]),
    "gcmb_simplex_body_destroy": (None, [vp]),
    "gcmb_simplex_info": (ctypes.c_int, [vp, c_int_p, c_int_p, c_int_p]),
    "gcmb_simplex_vertices": (ctypes.c_int, [vp, c_int_p, c_u8_p, c_double_p, c_double_p]),
    "gcmb_simplex_locate": (ctypes.c_int, [vp, ctypes.c_int, c_int_p, c_double_p, c_int_p]),
    "gcmb_simplex_errors": (ctypes.c_int, [vp, c_int_p]),
    "gcmb_simplex_set_material": (ctypes.c_int, [vp, c_double_p, c_double_p, c_double_p, c_double_p]),
    "gcmb_simplex_upload_state": (ctypes.c_int, [vp, c_double_p]),
    "gcmb_simplex_download_state": (ctypes.c_int, [vp, c_double_p]),
    "gcmb_simplex_border_set": (ctypes.c_int, [vp, ctypes.c_int, c_int_p, ctypes.c_int, c_int_p, c_double_p, c_int_p]),
    "gcmb_simplex_plain_border": (ctypes.c_int, [vp, c_double_p]),
    "gcmb_simplex_stage": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_double, c_double_p]),
    "gcmb_simplex_gradient": (ctypes.c_int, [vp, c_double_p, c_double_p]),
    "gcmb_simplex_set_gcm_type": (ctypes.c_int, [vp, ctypes.c_int]),
    "gcmb_simplex_ode_maxwell": (ctypes.c_int, [vp, ctypes.c_double]),
    "gcmb_simplex_set_splitting": (ctypes.c_int, [vp, ctypes.c_int]),
    "gcmb_simplex_set_local_bases": (ctypes.c_int, [vp, ctypes.c_int, c_int_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "gcmb_simplex_average_layers": (ctypes.c_int, [vp]),
    "gcmb_simplex_before_stage": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_double]),
    "gcmb_simplex_border_contact_stage": (ctypes.c_int, [vp]),
    "gcmb_simplex_border_correct": (ctypes.c_int, [vp, c_double_p]),
    "gcmb_simplex_inner_stage": (ctypes.c_int, [vp]),
    "gcmb_simplex_after_stage": (ctypes.c_int, [vp]),
    "gcmb_simplex_contact_normals": (ctypes.c_int, [vp, ctypes.c_int, c_double_p]),
    "gcmb_simplex_contact_create": (ctypes.c_int, [vp, vp, ctypes.c_int, c_int_p, c_int_p, c_double_p, ctypes.POINTER(vp)]),
    "gcmb_simplex_contact_destroy": (None, [vp]),
    "gcmb_simplex_contact_plain": (ctypes.c_int, [vp]),
    "gcmb_simplex_contact_correct": (ctypes.c_int, [vp]),
    "gcmb_cubic_stage_kernel_name": (ctypes.c_char_p, [vp, ctypes.c_int]),
    "gcmb_real_bytes": (ctypes.c_int, [vp]),
    "gcmb_set_fma": (ctypes.c_int, [vp, ctypes.c_int]),
    "gcmb_cubic_download_box_begin": (ctypes.c_int, [vp, c_int_p, c_int_p, vp]),
    "gcmb_cubic_download_box_end": (ctypes.c_int, [vp]),
    "gcmb_host_alloc_pinned": (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(vp)]),
    "gcmb_host_free_pinned": (None, [vp]),
    "gcmb_cubic_stage_fill_next_border": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                                         c_double_p, c_int_p]),
    "gcmb_cubic_stage_with_border": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_double, ctypes.c_int, c_double_p, c_int_p]),
    "gcmb_cubic_seismo_at": (ctypes.c_int, [vp, c_double_p, c_ll_p, ctypes.c_int, c_double_p, ctypes.c_int, c_int_p]),
    "gcmb_halo_exchange_bodies": (ctypes.c_int, [ctypes.POINTER(vp), ctypes.c_int]),
    "gcmb_triangle_interpolate": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_int_p]),
    "gcmb_cubic_seismo_begin": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, c_int_p]),
    "gcmb_cubic_seismo_end": (ctypes.c_int, [vp, c_double_p, c_ll_p, c_double_p, ctypes.c_int]),
}

HOST_ABI = {
    "gcmb_host_last_error": (ctypes.c_char_p, []),
    "gcmb_host_engine_create": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp,
                                               ctypes.POINTER(vp)]),
    "gcmb_host_engine_create2": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int,
                                                ctypes.c_int, ctypes.POINTER(vp)]),
    "gcmb_host_engine_destroy": (None, [vp]),
    "gcmb_host_engine_run": (ctypes.c_int, [vp]),
    "gcmb_host_engine_advance": (ctypes.c_int, [vp, ctypes.c_int]),
    "gcmb_host_engine_info": (ctypes.c_int, [vp, c_int_p, c_double_p, c_double_p]),
    "gcmb_host_engine_body_info": (ctypes.c_int, [vp, ctypes.c_size_t, c_int_p, c_int_p, c_int_p, c_int_p]),
    "gcmb_host_engine_body_pde": (ctypes.c_int, [vp, ctypes.c_size_t, c_double_p]),
    "gcmb_host_engine_body_handle": (vp, [vp, ctypes.c_size_t]),
    "gcmb_host_engine_context": (vp, [vp]),
    "gcmb_host_engine_body_matrices": (ctypes.c_int, [vp, ctypes.c_size_t, c_int_p, c_double_p, c_double_p, c_double_p]),
    "gcmb_host_engine_seismogram": (ctypes.c_int, [vp, c_double_p, c_float_p, ctypes.c_int]),
    "gcmb_host_matrices": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p,
                                          c_double_p, c_double_p]),
    "gcmb_host_simplex_box_mesh": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p, ctypes.c_double,
                                                  ctypes.c_double, ctypes.c_uint, c_double_p, ctypes.c_int, c_int_p,
                                                  c_double_p, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p]),
    "gcmb_host_simplex_triangulation": (ctypes.c_int, [vp, c_int_p, c_double_p, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p]),
    "gcmb_host_simplex_body_info": (ctypes.c_int, [vp, ctypes.c_size_t, c_int_p, c_double_p, c_double_p]),
    "gcmb_host_simplex_body_pde": (ctypes.c_int, [vp, ctypes.c_size_t, c_double_p, c_double_p, c_double_p, c_double_p]),
    "gcmb_host_simplex_border_nodes": (ctypes.c_int, [vp, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, c_int_p, c_int_p, c_double_p]),
    "gcmb_host_simplex_contact_nodes": (ctypes.c_int, [vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, c_int_p, c_int_p,
                                                       c_int_p, c_double_p]),
    "gcmb_host_simplex_errors": (ctypes.c_int, [vp, c_int_p]),
    "gcmb_host_simplex_save_inm": (ctypes.c_int, [vp, ctypes.c_char_p]),
    "gcmb_host_inm_read": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_double, c_int_p, c_double_p, c_int_p, c_int_p]),
}


class GcmError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("gcm_b200 error %d: %s" % (code, message))
        self.code = code


def dp(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def ip(a):
    return None if a is None else a.ctypes.data_as(c_int_p)


def bp(a):
    return None if a is None else a.ctypes.data_as(c_u8_p)


class Library:
    def __init__(self, cuda_path=None, host_path=None):
        cuda_path = cuda_path or os.path.join(HERE, "libgcm_b200.so")
        host_path = host_path or os.path.join(HERE, "libgcm_b200_host.so")
        for p in (cuda_path, host_path):
            if not os.path.exists(p):
                raise ImportError("%s is missing: build it with `python -m gcm_b200.build` "
                                  "(gcm_b200 has no fallback implementation)" % p)
        self.cuda_path, self.host_path = cuda_path, host_path
        # local + deep binding: each library resolves its references within itself and its own dependencies
        mode = ctypes.RTLD_LOCAL | os.RTLD_DEEPBIND
        self.c = ctypes.CDLL(cuda_path, mode=mode)
        for name, (res, args) in C_ABI.items():
            f = getattr(self.c, name)
            f.restype, f.argtypes = res, args
        self.h = ctypes.CDLL(host_path, mode=mode)
        for name, (res, args) in HOST_ABI.items():
            f = getattr(self.h, name)
            f.restype, f.argtypes = res, args

    def check(self, rc):
        if rc != 0:
            raise GcmError(rc, self.c.gcmb_last_error().decode())

    def hcheck(self, rc):
        if rc != 0:
            raise GcmError(rc, self.h.gcmb_host_last_error().decode())


class Context:
    def __init__(self, lib, device=0, real_bytes=8):
        self.lib = lib
        self.handle = vp()
        self.real = np.float32 if real_bytes == 4 else np.float64
        lib.check(lib.c.gcmb_create(device, real_bytes, ctypes.byref(self.handle)))

    def close(self):
        if self.handle:
            self.lib.c.gcmb_destroy(self.handle)
            self.handle = vp()

    def sync(self):
        self.lib.check(self.lib.c.gcmb_sync(self.handle))

    def timer_start(self):
        self.lib.check(self.lib.c.gcmb_timer_start(self.handle))

    def timer_stop(self):
        ms = ctypes.c_float()
        self.lib.check(self.lib.c.gcmb_timer_stop(self.handle, ctypes.byref(ms)))
        return ms.value

    def profile_enable(self, on=True):
        self.lib.check(self.lib.c.gcmb_profile_enable(self.handle, 1 if on else 0))

    def profile_get(self):
        ms = np.zeros(8)
        n = np.zeros(8, dtype=np.int64)
        self.lib.check(self.lib.c.gcmb_profile_get(self.handle, 8, dp(ms), n.ctypes.data_as(c_ll_p)))
        return ms, n

    def launch_count(self):
        return int(self.lib.c.gcmb_launch_count(self.handle))

    def device_bytes(self):
        return int(self.lib.c.gcmb_device_bytes(self.handle))


AREA_KINDS = {"infinite": 0, "box": 1, "sphere": 2, "cylinder": 3}


def area_args(area):
    """('box', lo, hi) / ('sphere', r, c) / ('cylinder', r, b, e) / ('infinite',) -> (kind, params array)"""
    kind = AREA_KINDS[area[0]]
    flat = []
    for a in area[1:]:
        flat.extend(a if isinstance(a, (list, tuple)) else [a])
    return kind, np.array(flat + [0.0] * (10 - len(flat)), dtype=np.float64)


class CubicBody:
    """Thin wrapper over gcmb_cubic_* for one body."""

    def __init__(self, ctx, D, M, sizes, start, h, border_size):
        self.ctx, self.lib = ctx, ctx.lib
        self.D, self.M, self.bs = D, M, border_size
        self.sizes = np.array(sizes, dtype=np.int32)
        self.start = np.array(start, dtype=np.int32)
        self.h = np.array(h, dtype=np.float64)
        self.handle = vp()
        self.lib.check(self.lib.c.gcmb_cubic_body_create(ctx.handle, D, M, ip(self.sizes), ip(self.start),
                                                         dp(self.h), border_size, ctypes.byref(self.handle)))
        self.border_nq = {}

    def close(self):
        if self.handle:
            self.lib.c.gcmb_cubic_body_destroy(self.handle)
            self.handle = vp()

    def upload(self, array, with_ghosts):
        a = np.ascontiguousarray(array, dtype=self.ctx.real)
        self.lib.check(self.lib.c.gcmb_cubic_upload_state(self.handle, a.ctypes.data_as(vp), 1 if with_ghosts else 0))

    def download(self, with_ghosts=False):
        ext = self.sizes + (2 * self.bs if with_ghosts else 0)
        out = np.empty(tuple(int(e) for e in ext) + (self.M,), dtype=self.ctx.real)
        self.lib.check(self.lib.c.gcmb_cubic_download_state(self.handle, out.ctypes.data_as(vp), 1 if with_ghosts else 0))
        return out

    def set_materials(self, U, U1, L, node_table=None):
        U = np.ascontiguousarray(U, dtype=np.float64)
        U1 = np.ascontiguousarray(U1, dtype=np.float64)
        L = np.ascontiguousarray(L, dtype=np.float64)
        nt = None if node_table is None else np.ascontiguousarray(node_table, dtype=np.uint8)
        self.lib.check(self.lib.c.gcmb_cubic_set_materials(self.handle, U.shape[0], dp(U), dp(U1), dp(L), bp(nt)))

    def assign_table_in_area(self, table, area):
        kind, p = area_args(area)
        self.lib.check(self.lib.c.gcmb_cubic_assign_table_in_area(self.handle, table, kind, dp(p)))

    def add_vector_in_area(self, vec, area):
        kind, p = area_args(area)
        v = np.ascontiguousarray(vec, dtype=np.float64)
        self.lib.check(self.lib.c.gcmb_cubic_add_vector_in_area(self.handle, dp(v), kind, dp(p)))

    def border_set(self, cond, direction, left_mask, right_mask, codes):
        codes = np.array(codes, dtype=np.int32)
        lm = None if left_mask is None else np.ascontiguousarray(left_mask, dtype=np.uint8)
        rm = None if right_mask is None else np.ascontiguousarray(right_mask, dtype=np.uint8)
        self.lib.check(self.lib.c.gcmb_cubic_border_set(self.handle, cond, direction, bp(lm), bp(rm), len(codes), ip(codes)))

    def border_set_area(self, cond, direction, area, codes, sides=3):
        kind, p = area_args(area)
        codes = np.array(codes, dtype=np.int32)
        self.lib.check(self.lib.c.gcmb_cubic_border_set_area(self.handle, cond, direction, sides, kind, dp(p),
                                                             len(codes), ip(codes)))

    def border_apply(self, direction, values):
        v = np.array(values, dtype=np.float64)
        self.lib.check(self.lib.c.gcmb_cubic_border_apply(self.handle, direction, len(v), dp(v)))

    def contact_apply(self, other, box_a, box_b, extent):
        a, b, e = (np.array(x, dtype=np.int32) for x in (box_a, box_b, extent))
        self.lib.check(self.lib.c.gcmb_cubic_contact_apply(self.handle, other.handle, ip(a), ip(b), ip(e)))

    def stage(self, direction, tau):
        self.lib.check(self.lib.c.gcmb_cubic_stage(self.handle, direction, tau))

    def stage_fill_next_border(self, direction, tau, next_direction, values):
        """stage + border fill of next_direction on the new layer; returns True when one kernel did both"""
        v = np.array(values, dtype=np.float64)
        fused = ctypes.c_int()
        self.lib.check(self.lib.c.gcmb_cubic_stage_fill_next_border(self.handle, direction, tau, next_direction, len(v), dp(v),
                                                                     ctypes.byref(fused)))
        return bool(fused.value)

    def stage_with_border(self, direction, tau, values):
        """border fill of `direction` + its stage; returns True when the stage kernel produced the ghost nodes itself"""
        v = np.array(values, dtype=np.float64)
        fused = ctypes.c_int()
        self.lib.check(self.lib.c.gcmb_cubic_stage_with_border(self.handle, direction, tau, len(v), dp(v), ctypes.byref(fused)))
        return bool(fused.value)

    def download_box(self, box_min, extent):
        """asynchronous read-back of a box of nodes (begin + end); [extent..., M]"""
        lo = np.array(box_min, dtype=np.int32)
        ext = np.array(extent, dtype=np.int32)
        out = np.empty(tuple(int(e) for e in ext) + (self.M,), dtype=self.ctx.real)
        self.lib.check(self.lib.c.gcmb_cubic_download_box_begin(self.handle, ip(lo), ip(ext), out.ctypes.data_as(vp)))
        self.lib.check(self.lib.c.gcmb_cubic_download_box_end(self.handle))
        return out

    def ode_maxwell(self, decay):
        d = np.array(decay, dtype=np.float64)
        self.lib.check(self.lib.c.gcmb_cubic_ode_maxwell(self.handle, dp(d)))

    def detector_set_area(self, code, area):
        kind, p = area_args(area)
        self.lib.check(self.lib.c.gcmb_cubic_detector_set_area(self.handle, code, kind, dp(p)))

    def seismo(self, line_comp=None):
        s = ctypes.c_double()
        c = ctypes.c_longlong()
        n = int(self.sizes[self.D - 1])
        line = np.zeros(n) if line_comp is not None else None
        self.lib.check(self.lib.c.gcmb_cubic_seismo(self.handle, ctypes.byref(s), ctypes.byref(c),
                                                    0 if line_comp is None else line_comp, dp(line), n))
        return s.value, c.value, line

    def checksum(self):
        out = ctypes.c_double()
        self.lib.check(self.lib.c.gcmb_cubic_checksum(self.handle, ctypes.byref(out)))
        return out.value

    def halo_exchange(self):
        self.lib.check(self.lib.c.gcmb_cubic_halo_exchange(self.handle))

    def kernel_name(self, direction):
        return self.lib.c.gcmb_cubic_stage_kernel_name(self.handle, direction).decode()

    def halo_get(self, side):
        buf = np.empty(self.lib.c.gcmb_cubic_halo_bytes(self.handle) // np.dtype(self.ctx.real).itemsize, dtype=self.ctx.real)
        self.lib.check(self.lib.c.gcmb_cubic_halo_get(self.handle, side, buf.ctypes.data_as(vp)))
        return buf

    def halo_put(self, side, buf):
        buf = np.ascontiguousarray(buf, dtype=self.ctx.real)
        assert buf.nbytes == self.lib.c.gcmb_cubic_halo_bytes(self.handle)
        self.lib.check(self.lib.c.gcmb_cubic_halo_put(self.handle, side, buf.ctypes.data_as(vp)))


class HostEngine:
    """createEngine(task)->run() of the host layer, driven by a plain-text task (gcm_b200/host/task_file.cpp)."""

    def __init__(self, lib, task_text, device=0, slab_rank=0, slab_count=1, nccl_id=None, real_bytes=8, fma=False):
        self.lib = lib
        self.handle = vp()
        idbuf = None
        if nccl_id is not None:
            idbuf = ctypes.create_string_buffer(bytes(nccl_id), 128)
        self._idbuf = idbuf
        lib.hcheck(lib.h.gcmb_host_engine_create2(task_text.encode(), device, slab_rank, slab_count,
                                                  ctypes.cast(idbuf, vp) if idbuf is not None else None,
                                                  real_bytes, 1 if fma else 0, ctypes.byref(self.handle)))

    def close(self):
        if self.handle:
            self.lib.h.gcmb_host_engine_destroy(self.handle)
            self.handle = vp()

    def run(self):
        self.lib.hcheck(self.lib.h.gcmb_host_engine_run(self.handle))
        return self

    def advance(self, n):
        self.lib.hcheck(self.lib.h.gcmb_host_engine_advance(self.handle, n))
        return self

    def info(self):
        steps = ctypes.c_int()
        t = ctypes.c_double()
        tau = ctypes.c_double()
        self.lib.hcheck(self.lib.h.gcmb_host_engine_info(self.handle, ctypes.byref(steps), ctypes.byref(t), ctypes.byref(tau)))
        return steps.value, t.value, tau.value

    def body_info(self, bid):
        D, M = ctypes.c_int(), ctypes.c_int()
        sizes = np.zeros(3, dtype=np.int32)
        start = np.zeros(3, dtype=np.int32)
        self.lib.hcheck(self.lib.h.gcmb_host_engine_body_info(self.handle, bid, ctypes.byref(D), ctypes.byref(M), ip(sizes), ip(start)))
        return D.value, M.value, sizes[:D.value].copy(), start[:D.value].copy()

    def body_pde(self, bid):
        D, M, sizes, _ = self.body_info(bid)
        out = np.empty((int(np.prod(sizes)), M), dtype=np.float64)
        self.lib.hcheck(self.lib.h.gcmb_host_engine_body_pde(self.handle, bid, dp(out)))
        return out

    def body_matrices(self, bid):
        D, M, _, _ = self.body_info(bid)
        n = ctypes.c_int()
        self.lib.hcheck(self.lib.h.gcmb_host_engine_body_matrices(self.handle, bid, ctypes.byref(n), None, None, None))
        U = np.zeros((n.value, D, M, M))
        U1 = np.zeros((n.value, D, M, M))
        L = np.zeros((n.value, D, M))
        self.lib.hcheck(self.lib.h.gcmb_host_engine_body_matrices(self.handle, bid, ctypes.byref(n), dp(U), dp(U1), dp(L)))
        return U, U1, L

    def seismogram(self):
        n = self.lib.h.gcmb_host_engine_seismogram(self.handle, None, None, 0)
        t = np.zeros(n)
        v = np.zeros(n, dtype=np.float32)
        self.lib.h.gcmb_host_engine_seismogram(self.handle, dp(t), v.ctypes.data_as(c_float_p), n)
        return t, v

    def body_handle(self, bid):
        return self.lib.h.gcmb_host_engine_body_handle(self.handle, bid)

    def kernel_name(self, bid, direction):
        return self.lib.c.gcmb_cubic_stage_kernel_name(self.body_handle(bid), direction).decode()

    def context_handle(self):
        return self.lib.h.gcmb_host_engine_context(self.handle)


class SimplexHostEngine(HostEngine):
    """the simplex engine of the host layer (gcm_b200/host/simplex_engine.cpp): task text with `grid simplex`"""

    def triangulation(self):
        sizes = np.zeros(3, dtype=np.int32)
        self.lib.hcheck(self.lib.h.gcmb_host_simplex_triangulation(self.handle, ip(sizes), None, None, None, None, None, None))
        nV, nC, ninc = (int(x) for x in sizes)
        t = dict(xyz=np.zeros((nV, 3)), cell_v=np.zeros((nC, 4), dtype=np.int32), cell_n=np.zeros((nC, 4), dtype=np.int32),
                 cell_grid=np.zeros(nC, dtype=np.int32), inc_off=np.zeros(nV + 1, dtype=np.int32),
                 inc_cell=np.zeros(ninc, dtype=np.int32))
        self.lib.hcheck(self.lib.h.gcmb_host_simplex_triangulation(self.handle, ip(sizes), dp(t["xyz"]), ip(t["cell_v"]), ip(t["cell_n"]),
                                                                   ip(t["cell_grid"]), ip(t["inc_off"]), ip(t["inc_cell"])))
        return t

    def simplex_body_info(self, bid):
        info = np.zeros(3, dtype=np.int32)
        reals = np.zeros(3)
        basis = np.zeros((3, 3))
        self.lib.hcheck(self.lib.h.gcmb_host_simplex_body_info(self.handle, bid, ip(info), dp(reals), dp(basis)))
        return dict(n_local=int(info[0]), M=int(info[1]), n_conditions=int(info[2]), average_height=reals[0],
                    minimal_height=reals[1], maximal_eigenvalue=reals[2], basis=basis)

    def simplex_pde(self, bid):
        i = self.simplex_body_info(bid)
        out = np.zeros((i["n_local"], i["M"]))
        self.lib.hcheck(self.lib.h.gcmb_host_simplex_body_pde(self.handle, bid, dp(out), None, None, None))
        return out

    def simplex_matrices(self, bid):
        M = self.simplex_body_info(bid)["M"]
        U, U1, L = np.zeros((3, M, M)), np.zeros((3, M, M)), np.zeros((3, M))
        self.lib.hcheck(self.lib.h.gcmb_host_simplex_body_pde(self.handle, bid, None, dp(U), dp(U1), dp(L)))
        return U, U1, L

    def border_nodes(self, bid, condition):
        n = ctypes.c_int()
        self.lib.hcheck(self.lib.h.gcmb_host_simplex_border_nodes(self.handle, bid, condition, 0, ctypes.byref(n), None, None))
        nodes = np.zeros(n.value, dtype=np.int32)
        normals = np.zeros((n.value, 3))
        self.lib.hcheck(self.lib.h.gcmb_host_simplex_border_nodes(self.handle, bid, condition, n.value, ctypes.byref(n), ip(nodes), dp(normals)))
        return nodes, normals

    def contact_nodes(self, a, b):
        n = ctypes.c_int()
        self.lib.hcheck(self.lib.h.gcmb_host_simplex_contact_nodes(self.handle, a, b, 0, ctypes.byref(n), None, None, None))
        first = np.zeros(n.value, dtype=np.int32)
        second = np.zeros(n.value, dtype=np.int32)
        normals = np.zeros((n.value, 3))
        self.lib.hcheck(self.lib.h.gcmb_host_simplex_contact_nodes(self.handle, a, b, n.value, ctypes.byref(n), ip(first), ip(second), dp(normals)))
        return first, second, normals

    def save_inm(self, path):
        self.lib.hcheck(self.lib.h.gcmb_host_simplex_save_inm(self.handle, str(path).encode()))

    def errors(self):
        c = ctypes.c_int()
        self.lib.hcheck(self.lib.h.gcmb_host_simplex_errors(self.handle, ctypes.byref(c)))
        return c.value


def triangle_interpolate(ctx, mode, points, values, gradients, queries):
    """gcmb_triangle_interpolate: (out [n], status [n])"""
    points = np.ascontiguousarray(points, dtype=np.float64)
    values = np.ascontiguousarray(values, dtype=np.float64)
    queries = np.ascontiguousarray(queries, dtype=np.float64)
    grads = None if gradients is None else np.ascontiguousarray(gradients, dtype=np.float64)
    n = len(queries)
    out = np.zeros(n)
    status = np.zeros(n, dtype=np.int32)
    ctx.lib.check(ctx.lib.c.gcmb_triangle_interpolate(ctx.handle, mode, n, dp(points), dp(values), dp(grads), dp(queries), dp(out), ip(status)))
    return out, status


def inm_read(lib, path, scale=1.0):
    """InmMeshLoader::readFromFile: (points [nV][3], cells [nC][4] 0-based, materials [nC])"""
    sizes = np.zeros(2, dtype=np.int32)
    lib.hcheck(lib.h.gcmb_host_inm_read(str(path).encode(), scale, ip(sizes), None, None, None))
    xyz = np.zeros((int(sizes[0]), 3))
    cells = np.zeros((int(sizes[1]), 4), dtype=np.int32)
    mats = np.zeros(int(sizes[1]), dtype=np.int32)
    lib.hcheck(lib.h.gcmb_host_inm_read(str(path).encode(), scale, ip(sizes), dp(xyz), ip(cells), ip(mats)))
    return xyz, cells, mats


def host_matrices(lib, model, D, material):
    """material: ('isotropic', rho, lambda, mu) or ('orthotropic', rho, c[9])"""
    M = D + 1 if model == "acoustic" else D + D * (D + 1) // 2
    U = np.zeros((D, M, M))
    U1 = np.zeros((D, M, M))
    L = np.zeros((D, M))
    if material[0] == "isotropic":
        p = np.array(material[1:4], dtype=np.float64)
        kind = 0
    else:
        p = np.array([material[1]] + list(material[2]), dtype=np.float64)
        kind = 1
    lib.hcheck(lib.h.gcmb_host_matrices(1 if model == "acoustic" else 0, D, kind, dp(p), dp(U), dp(U1), dp(L)))
    return U, U1, L
