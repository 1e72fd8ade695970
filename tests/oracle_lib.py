"""ctypes binding of oracle/liboracle.so (the CPU restatement of the reference).

TEST INFRASTRUCTURE: imported only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product (scone_b200/) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_lib = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def build():
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("capi.cpp", "cedata.hpp", "ceace.hpp", "cereact.hpp", "cephysics.hpp", "mathmode.hpp", "physics.hpp", "geom.hpp", "mgdata.hpp", "rng.hpp")]
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    if (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.exists(s)):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", "all"])
    return so


def dp(a):
    return a.ctypes.data_as(c_dp)


def ip(a):
    return a.ctypes.data_as(c_ip)


def load():
    global _lib
    if _lib is not None:
        return _lib
    so = build()
    L = C.CDLL(so)
    u64, i64, dbl, vp, i32 = C.c_uint64, C.c_int64, C.c_double, C.c_void_p, C.c_int
    sig = {
        "orc_last_error": (C.c_char_p, []),
        "orc_set_math_mode": (None, [i32]),
        "orc_get_math_mode": (i32, []),
        "orc_rng_next": (u64, [u64]),
        "orc_rng_real": (dbl, [u64]),
        "orc_rng_skip": (u64, [u64, i64]),
        "orc_rng_stride": (u64, [u64, C.c_int32]),
        "orc_math_log": (dbl, [dbl]),
        "orc_math_sincos": (None, [dbl, c_dp, c_dp]),
        "orc_rotate_vector": (None, [c_dp, dbl, dbl, c_dp]),
        "orc_grid_search_lin": (i32, [dbl, dbl, i32, dbl]),
        "orc_binary_search": (i32, [c_dp, i32, dbl]),
        "orc_geom_load": (vp, [C.c_char_p, i32]),
        "orc_geom_free": (None, [vp]),
        "orc_geom_info": (i32, [vp] + [c_ip] * 8),
        "orc_geom_graph": (i32, [vp, c_ip, c_ip]),
        "orc_geom_active_mats": (i32, [vp, c_ip, i32]),
        "orc_geom_uni_fill": (i32, [vp, i32, c_ip, i32]),
        "orc_geom_bounds": (i32, [vp, c_dp]),
        "orc_geom_what_is_at": (i32, [vp, c_dp, c_dp, c_ip, c_ip]),
        "orc_geom_what_is_at_n": (i32, [vp, C.c_long, c_dp, c_dp, c_ip, c_ip]),
        "orc_geom_teleport_n": (i32, [vp, C.c_long, c_dp, c_dp, c_dp, c_ip, c_ip]),
        "orc_coords_new": (vp, [vp]),
        "orc_coords_free": (None, [vp]),
        "orc_coords_init": (i32, [vp, c_dp, c_dp]),
        "orc_coords_place": (i32, [vp]),
        "orc_coords_move": (i32, [vp, c_dp, c_ip, i32]),
        "orc_coords_move_global": (i32, [vp, c_dp, c_ip]),
        "orc_coords_teleport": (i32, [vp, dbl]),
        "orc_coords_rotate": (i32, [vp, dbl, dbl]),
        "orc_coords_closest": (i32, [vp, c_dp, c_ip, c_ip]),
        "orc_coords_level_distance": (i32, [vp, i32, c_dp, c_ip]),
        "orc_coords_get": (i32, [vp, c_ip, c_ip, c_ip, c_dp, c_dp, c_ip, c_ip, c_ip, c_ip]),
        "orc_uni_new": (vp, [C.c_char_p, C.c_char_p, C.c_char_p, i32]),
        "orc_uni_free": (None, [vp]),
        "orc_uni_fill": (i32, [vp, c_ip, i32]),
        "orc_uni_enter": (i32, [vp, c_dp, c_dp, c_dp, c_dp, c_ip, c_ip, c_ip]),
        "orc_uni_distance": (i32, [vp, i32, c_dp, c_dp, c_dp, c_ip]),
        "orc_uni_cross": (i32, [vp, i32, c_dp, c_dp, i32, c_ip]),
        "orc_uni_offset": (i32, [vp, i32, c_dp]),
        "orc_surf_new": (vp, [C.c_char_p]),
        "orc_surf_free": (None, [vp]),
        "orc_surf_set_bc": (i32, [vp, c_ip, i32]),
        "orc_surf_query": (i32, [vp, c_dp, c_dp, c_dp, c_dp, c_ip, c_ip]),
        "orc_surf_bc": (i32, [vp, i32, c_dp, c_dp]),
        "orc_mg_load": (vp, [C.c_char_p, C.c_char_p]),
        "orc_mg_free": (None, [vp]),
        "orc_mg_info": (i32, [vp, c_ip, c_ip]),
        "orc_mg_mat_idx": (i32, [vp, C.c_char_p]),
        "orc_mg_macro": (i32, [vp, i32, i32, c_dp]),
        "orc_mg_majorant": (dbl, [vp, i32]),
        "orc_mg_total": (dbl, [vp, i32, i32]),
        "orc_mg_matrices": (i32, [vp, i32, c_dp, c_dp, c_dp, c_dp, c_dp]),
        "orc_mg_sample_scatter": (u64, [vp, i32, i32, u64, c_dp, c_dp, c_ip]),
        "orc_mg_sample_fission": (u64, [vp, i32, u64, c_dp, c_dp, c_ip]),
        "orc_dungeon_norm_size": (i32, [i32, c_ip, c_ip, i32, i32, u64]),
        "orc_dungeon_sort": (i32, [i32, c_ip, c_ip]),
        "orc_eigen_load": (vp, [C.c_char_p, C.c_char_p]),
        "orc_eigen_free": (None, [vp]),
        "orc_eigen_info": (i32, [vp] + [c_ip] * 6),
        "orc_eigen_rng_state": (u64, [vp]),
        "orc_eigen_set_rng_state": (None, [vp, u64]),
        "orc_eigen_keff0": (dbl, [vp]), "orc_eigen_keff": (i32, [vp, i32, c_dp, c_dp]),
        "orc_eigen_init_source": (i32, [vp]),
        "orc_eigen_cycle": (dbl, [vp, i32, dbl]),
        "orc_eigen_run": (i32, [vp]),
        "orc_eigen_bank_size": (i32, [vp]),
        "orc_eigen_bank": (i32, [vp, c_dp, c_dp, c_dp, c_ip, c_ip]),
        "orc_eigen_set_bank": (i32, [vp, i32, c_dp, c_dp, c_dp, c_ip]),
        "orc_eigen_tally_size": (C.c_long, [vp, i32]),
        "orc_eigen_tally": (i32, [vp, i32, c_dp, c_dp, c_ip]),
        "orc_eigen_stats": (i32, [vp, C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_long)]),
    }
    sig.update({
        "orc_ce_nuclide_from_ace": (vp, [C.c_char_p, i32]), "orc_ce_nuclide_from_arrays": (vp, [i32, i32, c_dp, c_dp]),
        "orc_ce_nuclide_free": (None, [vp]), "orc_ce_nuclide_info": (i32, [vp, c_ip, c_ip, c_dp, c_dp]),
        "orc_ce_nuclide_data": (i32, [vp, c_dp, c_dp]), "orc_ce_nuclide_search": (i32, [vp, dbl, c_ip, c_dp]),
        "orc_ce_nuclide_micro": (i32, [vp, dbl, c_dp]), "orc_ce_nuclide_total": (dbl, [vp, dbl]),
        "orc_ce_nuclide_nubar": (i32, [vp, dbl, c_dp, c_dp, c_dp]),
        "orc_ce_db_new": (vp, []), "orc_ce_db_free": (None, [vp]), "orc_ce_db_add_nuclide": (i32, [vp, vp]),
        "orc_ce_db_add_material": (i32, [vp, i32, c_ip, c_dp]), "orc_ce_db_finalise": (i32, [vp]),
        "orc_ce_db_union": (i32, [vp, c_dp, c_dp]), "orc_ce_db_total_n": (i32, [vp, C.c_long, c_dp, c_ip, c_dp]),
        "orc_ce_db_macro_n": (i32, [vp, C.c_long, c_dp, c_ip, c_dp]), "orc_ce_db_majorant_n": (i32, [vp, C.c_long, c_dp, c_dp]),
        "orc_ce_db_index_n": (i32, [vp, i32, C.c_long, c_dp, c_ip]),
        "orc_mem_new": (vp, [C.c_long, i32]), "orc_mem_free": (None, [vp]), "orc_mem_score": (i32, [vp, dbl, C.c_long]),
        "orc_mem_accumulate": (i32, [vp, dbl, C.c_long]), "orc_mem_reduce": (i32, [vp]), "orc_mem_close_bin": (i32, [vp, dbl, C.c_long]),
        "orc_mem_close_cycle": (i32, [vp, dbl]), "orc_mem_last_cycle": (i32, [vp]), "orc_mem_get_score": (dbl, [vp, C.c_long]),
        "orc_mem_result": (i32, [vp, C.c_long, i32, c_dp, c_dp]),
        "orc_keff_implicit_sequence": (i32, [dbl, dbl, dbl, dbl, i32, c_dp, c_dp, c_dp, c_dp, c_dp]),
        "orc_keff_analog_sequence": (i32, [i32, c_dp, c_dp, c_dp, dbl, c_dp, c_dp]),
        "orc_eigen_bank_E": (i32, [vp, c_dp]), "orc_clerk_sequence": (i32, [C.c_char_p, C.c_char_p, dbl, dbl, i32, i32, c_ip, c_dp, c_dp, c_dp, i32]),
        "orc_grid": (i32, [i32, dbl, dbl, i32, c_dp, i32, i32, c_dp, c_ip, c_dp, i32]),
        "orc_ace_card_info": (i32, [C.c_char_p, c_dp, c_ip, C.c_char_p]), "orc_ce_nuclide_elastic_isotropic": (i32, [vp]),
        "orc_heap_queue": (dbl, [i32, i32, c_dp, i32, c_ip]),
        "orc_shannon_sequence": (i32, [C.c_char_p, C.c_char_p, i32, c_ip, c_ip, c_dp, c_dp, c_dp]),
        "orc_response_value": (dbl, [C.c_char_p, c_dp]),
        "orc_map_new": (vp, [C.c_char_p, C.c_char_p]), "orc_map_free": (None, [vp]), "orc_map_bins": (i32, [vp]),
        "orc_map_map": (i32, [vp, c_dp, dbl, i32, i32, i32]), "orc_fixed_cycle": (i32, [vp]), "orc_eigen_is_fixed": (i32, [vp]),
        "orc_tabpdf_sample": (dbl, [i32, c_dp, c_dp, c_dp, i32, dbl]),
        "orc_endftable_at": (dbl, [i32, c_dp, c_dp, i32, c_ip, c_ip, dbl]),
        "orc_ce_nuclide_from_acebin": (vp, [C.c_char_p]), "orc_ce_nuclide_mt_list": (i32, [vp, c_ip, c_ip]),
        "orc_ce_nuclide_sample": (C.c_long, [vp, i32, i32, dbl, u64, c_dp]),
        "orc_ce_nuclide_invert_inelastic": (i32, [vp, dbl, u64]),
        "orc_ce_nuclide_mt_release": (dbl, [vp, i32, dbl]), "orc_ce_nuclide_mt_cm": (i32, [vp, i32]),
    })
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def err(L):
    return L.orc_last_error().decode()


class Geom:
    """Geometry handle + a coordList, mirroring how the reference tests drive geometryStd."""

    def __init__(self, L, text, is_path=False):
        self.L = L
        self.h = L.orc_geom_load(text.encode(), 1 if is_path else 0)
        if not self.h:
            raise RuntimeError(err(L))
        self.c = L.orc_coords_new(self.h)

    def info(self):
        v = [C.c_int() for _ in range(8)]
        self.L.orc_geom_info(self.h, *[C.byref(x) for x in v])
        keys = ["nSurf", "nCell", "nUni", "nGraph", "uniqueCells", "rootIdx", "borderIdx", "nesting"]
        return dict(zip(keys, [x.value for x in v]))

    def graph(self):
        n = self.info()["nGraph"]
        idx = np.zeros(n, np.int32)
        gid = np.zeros(n, np.int32)
        self.L.orc_geom_graph(self.h, ip(idx), ip(gid))
        return idx, gid

    def what_is_at(self, r, u=None):
        r = np.asarray(r, np.float64)
        m, q = C.c_int(), C.c_int()
        up = dp(np.asarray(u, np.float64)) if u is not None else None
        if self.L.orc_geom_what_is_at(self.h, dp(r), up, C.byref(m), C.byref(q)) != 0:
            raise RuntimeError(err(self.L))
        return m.value, q.value

    def bounds(self):
        b = np.zeros(6)
        self.L.orc_geom_bounds(self.h, dp(b))
        return b

    def init(self, r, u):
        self.L.orc_coords_init(self.c, dp(np.asarray(r, np.float64)), dp(np.asarray(u, np.float64)))

    def place(self):
        if self.L.orc_coords_place(self.c) != 0:
            raise RuntimeError(err(self.L))

    def get(self):
        nest, mat, uid = C.c_int(), C.c_int(), C.c_int()
        r = np.zeros((12, 3)); d = np.zeros((12, 3))
        a = [np.zeros(12, np.int32) for _ in range(4)]
        self.L.orc_coords_get(self.c, C.byref(nest), C.byref(mat), C.byref(uid), dp(r), dp(d), *[ip(x) for x in a])
        return dict(nesting=nest.value, mat=mat.value, uid=uid.value, r=r, dir=d, uniIdx=a[0], uniRootID=a[1], localID=a[2], cellIdx=a[3])

    def move(self, max_dist, cache=False):
        md, ev = C.c_double(max_dist), C.c_int()
        if self.L.orc_coords_move(self.c, C.byref(md), C.byref(ev), 1 if cache else 0) != 0:
            raise RuntimeError(err(self.L))
        return md.value, ev.value

    def move_global(self, max_dist):
        md, ev = C.c_double(max_dist), C.c_int()
        if self.L.orc_coords_move_global(self.c, C.byref(md), C.byref(ev)) != 0:
            raise RuntimeError(err(self.L))
        return md.value, ev.value

    def teleport(self, dist):
        if self.L.orc_coords_teleport(self.c, dist) != 0:
            raise RuntimeError(err(self.L))

    def level_distance(self, lvl):
        d, s = C.c_double(), C.c_int()
        if self.L.orc_coords_level_distance(self.c, lvl, C.byref(d), C.byref(s)) != 0:
            raise RuntimeError(err(self.L))
        return d.value, s.value

    def closest(self):
        d, s, l = C.c_double(), C.c_int(), C.c_int()
        if self.L.orc_coords_closest(self.c, C.byref(d), C.byref(s), C.byref(l)) != 0:
            raise RuntimeError(err(self.L))
        return d.value, s.value, l.value

    def slice_plot(self, shape, centre, axis, what, width=None):
        """geometry_inter.f90:322-415 (pixel centres)."""
        ax = "xyz".index(axis)
        plane = [a for a in range(3) if a != ax]
        centre = np.asarray(centre, np.float64)
        low = np.zeros(3); top = np.zeros(3)
        if width is not None:
            for k, p in enumerate(plane):
                low[p] = centre[p] - width[k] * 0.5
                top[p] = centre[p] + width[k] * 0.5
        else:
            b = self.bounds()
            low[:] = b[:3]; top[:] = b[3:]
        low[ax] = centre[ax]; top[ax] = centre[ax]
        step = np.zeros(3)
        for k, p in enumerate(plane):
            step[p] = (top[p] - low[p]) / shape[k]
        corner = low - 0.5 * step
        img = np.zeros(shape, np.int64)
        pt = corner.copy()
        for j in range(1, shape[1] + 1):
            pt[plane[1]] = corner[plane[1]] + step[plane[1]] * j
            for i in range(1, shape[0] + 1):
                pt[plane[0]] = corner[plane[0]] + step[plane[0]] * i
                m, q = self.what_is_at(pt)
                img[i - 1, j - 1] = m if what == "material" else q
        return img


class Uni:
    """One universe built from a dictionary string, as the reference's universe unit tests do."""

    def __init__(self, L, text, mats, idx, env=""):
        self.L = L
        ml = " ".join("%s %d" % (k, v) for k, v in mats.items())
        self.h = L.orc_uni_new(text.encode(), env.encode(), ml.encode(), idx)
        if not self.h:
            raise RuntimeError(err(L))

    def fill(self):
        out = np.zeros(4096, np.int32)
        n = self.L.orc_uni_fill(self.h, ip(out), 4096)
        return out[:n].tolist()

    def enter(self, r, u):
        r = np.asarray(r, np.float64); u = np.asarray(u, np.float64)
        ro = np.zeros(3); uo = np.zeros(3)
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        if self.L.orc_uni_enter(self.h, dp(r), dp(u), dp(ro), dp(uo), C.byref(a), C.byref(b), C.byref(c)) != 0:
            raise RuntimeError(err(self.L))
        return dict(r=ro, dir=uo, uniIdx=a.value, localID=b.value, cellIdx=c.value)

    def distance(self, localID, r, u):
        r = np.asarray(r, np.float64); u = np.asarray(u, np.float64)
        d, s = C.c_double(), C.c_int()
        if self.L.orc_uni_distance(self.h, localID, dp(r), dp(u), C.byref(d), C.byref(s)) != 0:
            raise RuntimeError(err(self.L))
        return d.value, s.value

    def cross(self, localID, r, u, surfIdx):
        r = np.asarray(r, np.float64); u = np.asarray(u, np.float64)
        n = C.c_int()
        if self.L.orc_uni_cross(self.h, localID, dp(r), dp(u), surfIdx, C.byref(n)) != 0:
            raise RuntimeError(err(self.L))
        return n.value

    def offset(self, localID):
        o = np.zeros(3)
        self.L.orc_uni_offset(self.h, localID, dp(o))
        return o


class Surf:
    def __init__(self, L, text):
        self.L = L
        self.h = L.orc_surf_new(text.encode())
        if not self.h:
            raise RuntimeError(err(L))

    def set_bc(self, bc):
        a = np.asarray(bc, np.int32)
        if self.L.orc_surf_set_bc(self.h, ip(a), len(a)) != 0:
            raise RuntimeError(err(self.L))

    def query(self, r, u):
        r = np.asarray(r, np.float64); u = np.asarray(u, np.float64)
        e, d, g, h = C.c_double(), C.c_double(), C.c_int(), C.c_int()
        self.L.orc_surf_query(self.h, dp(r), dp(u), C.byref(e), C.byref(d), C.byref(g), C.byref(h))
        return dict(evaluate=e.value, distance=d.value, going=bool(g.value), halfspace=bool(h.value))

    def bc(self, r, u, transform):
        r = np.array(r, np.float64); u = np.array(u, np.float64)
        self.L.orc_surf_bc(self.h, 1 if transform else 0, dp(r), dp(u))
        return r, u
