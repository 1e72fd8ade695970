"""Continuous-energy cross-section lookup on the engine: the Python face of sb_load_ce_data / sb_ce_lookup.

A nuclide is (eGrid[N], mainData[N][rows]) exactly as aceNeutronNuclide holds them after init
(NuclearData/ceNeutronData/aceDatabase/aceNeutronNuclide_class.f90:737-948; rows = 4 or 8, the rows of one energy
point contiguous = Fortran mainData(rows, N)); a material is a list of (nuclide index 1-based, atomic density).
In SCONE these arrays are built by aceNeutronDatabase%init from the ACE library; here they come from the caller."""
import ctypes as C

import numpy as np

from .lib import EngineError, load_library


class _CeFlat(C.Structure):
    _fields_ = [("n_nuc", C.c_int32), ("grid_size", C.POINTER(C.c_int32)), ("rows", C.POINTER(C.c_int32)),
                ("grid", C.POINTER(C.c_double)), ("data", C.POINTER(C.c_double)),
                ("n_mat", C.c_int32), ("mat_off", C.POINTER(C.c_int32)), ("mat_nuc", C.POINTER(C.c_int32)), ("mat_dens", C.POINTER(C.c_double))]


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class CeDatabase:
    """aceNeutronDatabase on the device: unionised grid + majorant + per-nuclide index table, batched lookups."""

    def __init__(self, nuclides, materials, device=0):
        self.L = load_library()
        self.eng = C.c_void_p()
        if self.L.sb_create(C.byref(self.eng), device) != 0:
            raise EngineError(self.L.sb_last_error(None).decode())
        self.n_nuc, self.n_mat = len(nuclides), len(materials)
        gs = np.array([len(g) for g, _ in nuclides], np.int32)
        rows = np.array([np.asarray(d).shape[1] for _, d in nuclides], np.int32)
        grid = np.ascontiguousarray(np.concatenate([np.asarray(g, np.float64) for g, _ in nuclides]))
        data = np.ascontiguousarray(np.concatenate([np.asarray(d, np.float64).reshape(-1) for _, d in nuclides]))
        off = np.zeros(self.n_mat + 1, np.int32)
        off[1:] = np.cumsum([len(m) for m in materials])
        nuc = np.array([n for m in materials for n, _ in m], np.int32)
        dens = np.array([d for m in materials for _, d in m], np.float64)
        f = _CeFlat(self.n_nuc, _ip(gs), _ip(rows), _dp(grid), _dp(data), self.n_mat, _ip(off), _ip(nuc), _dp(dens))
        if self.L.sb_load_ce_data(self.eng, C.byref(f)) != 0:
            raise EngineError(self._err())
        self.mat_sizes = [len(m) for m in materials]

    def _err(self):
        return self.L.sb_last_error(self.eng).decode()

    def union(self):
        n = self.L.sb_ce_union_size(self.eng)
        g = np.zeros(n); m = np.zeros(n)
        self.L.sb_ce_union(self.eng, _dp(g), _dp(m))
        return g, m

    def lookup(self, E, mat=None, total=False, macro=False, majorant=False):
        """Host arrays in, host arrays out (copies inside the call). Returns a dict of the requested outputs."""
        E = np.ascontiguousarray(E, np.float64)
        n = len(E)
        m = None if mat is None else np.ascontiguousarray(mat, np.int32)
        out = {}
        t = np.zeros(n) if total else None
        x = np.zeros((n, 8)) if macro else None
        j = np.zeros(n) if majorant else None
        rc = self.L.sb_ce_lookup(self.eng, n, _dp(E), None if m is None else _ip(m), None if t is None else _dp(t),
                                 None if x is None else _dp(x), None if j is None else _dp(j))
        if rc != 0:
            raise EngineError(self._err())
        if total:
            out["total"] = t
        if macro:
            out["macro"] = x
        if majorant:
            out["majorant"] = j
        return out

    def lookup_into(self, E, mat, total):
        """Host arrays in, PREALLOCATED host array out (e.g. page-locked numpy views of pinned torch tensors): Sigma_t per particle."""
        if not (E.flags.c_contiguous and mat.flags.c_contiguous and total.flags.c_contiguous) or E.dtype != np.float64 or mat.dtype != np.int32 or total.dtype != np.float64:
            raise ValueError("lookup_into: contiguous float64 / int32 / float64 arrays of equal length are required")
        if self.L.sb_ce_lookup(self.eng, len(E), _dp(E), _ip(mat), _dp(total), None, None) != 0:
            raise EngineError(self._err())
        return total

    def lookup_device(self, dE, dMat, dTotal=0, dMacro=0, dMajorant=0, n=None, sort=False):
        """Device pointers (ints, e.g. torch tensor .data_ptr()); returns the CUDA-event time of the kernel(s) in ms.
        sort=True: the engine bins the lookups by (material, energy) first (sb_ce_lookup_sorted_device)."""
        f = self.L.sb_ce_lookup_sorted_device if sort else self.L.sb_ce_lookup_device
        if f(self.eng, n, dE, dMat, dTotal or None, dMacro or None, dMajorant or None) != 0:
            raise EngineError(self._err())
        ms = C.c_double()
        self.L.sb_ce_last_kernel_ms(self.eng, C.byref(ms))
        return ms.value

    def nuclide_index(self, nuc_idx, E):
        E = np.ascontiguousarray(E, np.float64)
        idx = np.zeros(len(E), np.int32)
        if self.L.sb_ce_nuclide_index(self.eng, nuc_idx, len(E), _dp(E), _ip(idx)) != 0:
            raise EngineError(self._err())
        return idx

    def memory(self):
        """(bytes of the nuclide tables as the reference holds them, bytes of the lookup structures on top, union index table built?)"""
        raw, idx, tab = C.c_int64(), C.c_int64(), C.c_int32()
        if self.L.sb_ce_memory(self.eng, C.byref(raw), C.byref(idx), C.byref(tab)) != 0:
            raise EngineError(self._err())
        return raw.value, idx.value, bool(tab.value)

    def launch_count(self):
        return int(self.L.sb_launch_count(self.eng))

    def close(self):
        if self.eng:
            self.L.sb_destroy(self.eng)
            self.eng = None


def refine_nuclide(grid, data, factor):
    """A nuclide with `factor` times as many grid intervals: factor - 1 points inserted in every interval, log-spaced in energy,
    cross sections interpolated linearly as the lookup does (the shape of the grid - resonance clusters, sparse tails - is kept)."""
    g = np.asarray(grid, np.float64); d = np.asarray(data, np.float64)
    if factor <= 1:
        return g.copy(), d.copy()
    t = np.arange(factor)[None, :] / factor
    lo, hi = g[:-1, None], g[1:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        e = np.where(lo > 0, lo * (hi / np.where(lo > 0, lo, 1.0)) ** t, lo + (hi - lo) * t)
    e[:, 0] = g[:-1]
    f = ((e - lo) / np.where(hi > lo, hi - lo, 1.0))[:, :, None]
    x = d[1:, None, :] * f + (1.0 - f) * d[:-1, None, :]
    gg = np.concatenate([e.reshape(-1), g[-1:]]); dd = np.concatenate([x.reshape(-1, d.shape[1]), d[-1:]])
    # a zero-width interval of the original grid (a discontinuity: the same energy twice) stays one pair of points
    keep = np.ones(len(gg), bool)
    zero = np.repeat(g[1:] == g[:-1], factor)
    inner = np.tile(np.arange(factor) > 0, len(g) - 1)
    keep[:-1] = ~(zero & inner)
    return np.ascontiguousarray(gg[keep]), np.ascontiguousarray(dd[keep])


def synthetic_nuclides(base, n_total, seed=2026):
    """BASELINE configs[4] stress: n_total nuclides made from the `base` list of (grid, data) by seeded energy shifts
    of the interior grid points (end points kept, order kept), cross sections unchanged."""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n_total):
        g, d = base[k % len(base)]
        g = np.asarray(g, np.float64).copy()
        if k >= len(base):
            s = 1.0 + rng.uniform(-0.02, 0.02)
            gi = np.clip(g[1:-1] * s, g[0], g[-1])
            g[1:-1] = np.sort(gi)
        out.append((g, np.asarray(d, np.float64)))
    return out


def large_library(n_nuclides=300, per_material=20, seed=11):
    """A library that does not fit in the GPU's L2: the reference's five bundled nuclides refined to 2e4 - 6e4 grid points
    (refine_nuclide) and cloned with seeded energy shifts to n_nuclides; materials of per_material nuclides that together use
    every nuclide once.  Returns (nuclides, materials) for CeDatabase."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = np.load(os.path.join(root, "tests", "golden", "ce_nuclides.npz"))
    factors = {"1001": 64, "92233": 16, "52126": 16, "91231": 32, "91232": 96}
    base = [refine_nuclide(g["grid_" + n], g["data_" + n], f) for n, f in factors.items()]
    nuclides = synthetic_nuclides(base, n_nuclides, seed=seed)
    rng = np.random.default_rng(seed + 1)
    order = rng.permutation(n_nuclides)
    materials = [[(int(k) + 1, float(rng.uniform(1e-5, 5e-2))) for k in order[i:i + per_material]] for i in range(0, n_nuclides, per_material)]
    return nuclides, materials
