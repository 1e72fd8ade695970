"""Shared helpers of the continuous-energy tests: the golden nuclide fixture and database builders."""
import ctypes as C
import os

import numpy as np

from tests import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["1001", "92233", "52126", "91231", "91232"]


def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "ce_nuclides.npz"))


def base_nuclides():
    g = golden()
    return [(g["grid_" + n], g["data_" + n]) for n in NAMES]


# two materials in the spirit of the reference's test input (water-like, fuel-like) + an all-nuclide mixture
MATERIALS_5 = [[(1, 5.028e-2), (3, 2.505e-2)], [(2, 2.286e-2), (3, 4.572e-2), (4, 1.0e-4), (5, 2.0e-5)], [(1, 1e-2), (2, 2e-2), (3, 3e-3), (4, 4e-4), (5, 5e-5)]]


def oracle_db(orc, nuclides, materials):
    db = orc.orc_ce_db_new()
    for g, d in nuclides:
        g = np.ascontiguousarray(g, np.float64); d = np.ascontiguousarray(d, np.float64)
        h = orc.orc_ce_nuclide_from_arrays(len(g), d.shape[1], ol.dp(g), ol.dp(d))
        orc.orc_ce_db_add_nuclide(db, h)
        orc.orc_ce_nuclide_free(h)
    for m in materials:
        nuc = np.array([n for n, _ in m], np.int32); dens = np.array([x for _, x in m], np.float64)
        orc.orc_ce_db_add_material(db, len(m), ol.ip(nuc), ol.dp(dens))
    nU = orc.orc_ce_db_finalise(db)
    assert nU > 0, ol.err(orc)
    return db, nU


def log_uniform(n, lo, hi, seed):
    rng = np.random.default_rng(seed)
    return np.exp(rng.uniform(np.log(lo), np.log(hi), n))
