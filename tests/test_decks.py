"""The decks this repo authors (decks/gen_decks.py) against the reference's own input files, where /root/reference is present:
both are loaded by the oracle's deck loader and must flatten to the same model - geometry graph (idx, id of every entry),
sizes, bounds, and for every material and group the macroscopic set, the scattering / production matrices, chi and the majorant.
(The GPU box has no /root/reference: the test skips there; the decks themselves travel.)"""
import ctypes as C
import os

import numpy as np
import pytest

from tests import oracle_lib as ol
from tests.gpu_util import DECK

REF = "/root/reference/InputFiles"
PAIRS = [("c5g7", os.path.join(REF, "Benchmarks", "Multigroup", "C5G7")), ("inf", os.path.join(REF, "SCONE_Inf")), ("slab", os.path.join(REF, "SCONE_Slab"))]
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout is not on this machine")


@pytest.fixture(scope="module")
def orc():
    return ol.load()


@pytest.mark.parametrize("name,ref", PAIRS)
def test_geometry_graph_equals_reference_deck(orc, name, ref):
    a, b = ol.Geom(orc, DECK[name], is_path=True), ol.Geom(orc, ref, is_path=True)
    assert a.info() == b.info()
    for x, y in zip(a.graph(), b.graph()):
        assert np.array_equal(x, y)
    assert np.array_equal(a.bounds(), b.bounds())
    # the same material and unique cell at random points (names are compared through the material menu order: equal graphs)
    rng = np.random.default_rng(3)
    lo, hi = a.bounds()[:3], a.bounds()[3:]
    lo = np.maximum(lo, -100.0); hi = np.minimum(hi, 100.0)
    for r in rng.uniform(lo, hi, size=(300, 3)):
        assert a.what_is_at(r) == b.what_is_at(r)


@pytest.mark.parametrize("name,ref", PAIRS)
def test_mg_data_equals_reference_deck(orc, name, ref):
    da, db = orc.orc_mg_load(DECK[name].encode(), b"mg"), orc.orc_mg_load(ref.encode(), b"mg")
    assert da and db, ol.err(orc)
    na, ga, nb, gb = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    orc.orc_mg_info(da, C.byref(na), C.byref(ga)); orc.orc_mg_info(db, C.byref(nb), C.byref(gb))
    assert (na.value, ga.value) == (nb.value, gb.value)
    nG = ga.value
    for g in range(1, nG + 1):
        assert orc.orc_mg_majorant(da, g) == orc.orc_mg_majorant(db, g)
    for m in range(1, na.value + 1):
        for g in range(1, nG + 1):
            xa, xb = np.zeros(8), np.zeros(8)
            assert orc.orc_mg_macro(da, m, g, ol.dp(xa)) == 0 and orc.orc_mg_macro(db, m, g, ol.dp(xb)) == 0
            assert np.array_equal(xa, xb), "macroscopic set of material %d group %d" % (m, g)
        A = [np.zeros(nG * nG) for _ in range(3)] + [np.zeros(nG) for _ in range(2)]
        B = [np.zeros(nG * nG) for _ in range(3)] + [np.zeros(nG) for _ in range(2)]
        assert orc.orc_mg_matrices(da, m, *[ol.dp(x) for x in A]) == orc.orc_mg_matrices(db, m, *[ol.dp(x) for x in B])
        for x, y in zip(A, B):
            assert np.array_equal(x, y)
    orc.orc_mg_free(da); orc.orc_mg_free(db)
