"""CPU: the oracle's continuous-energy restatement against the reference's own regression values and the golden fixture."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import ce_util
from tests import oracle_lib as ol

REF = "/root/reference/IntegrationTestFiles/"
have_ref = os.path.exists(REF + "1001JEF311.ace")


@pytest.mark.skipif(not have_ref, reason="reference ACE files are not on this box")
def test_ace_loader_matches_reference_regression_values(orc):
    # aceNeutronDatabase_iTest.f90:181-195 (TOL = 1e-6 relative)
    h = orc.orc_ce_nuclide_from_ace((REF + "1001JEF311.ace").encode(), 1779)
    assert h, ol.err(orc)
    assert orc.orc_ce_nuclide_total(h, 1.1e-6) == pytest.approx(20.765855864000002, rel=1e-6)
    mic = np.zeros(8)
    assert orc.orc_ce_nuclide_micro(h, 5.6e-3, ol.dp(mic)) == 0
    assert mic[0] == pytest.approx(19.731020820000000, rel=1e-6)
    assert mic[1] == pytest.approx(19.730326000000000, rel=1e-6)
    assert mic[3] == pytest.approx(6.948036800000000e-04, rel=1e-6)
    assert mic[2] == 0.0 and mic[4] == 0.0 and mic[5] == 0.0
    orc.orc_ce_nuclide_free(h)
    # fissionCE_iTest.f90:55-60 (U-233 neutron release; pFUnit assertEqual(a, b, TOL) is an absolute tolerance of 1e-6)
    u = orc.orc_ce_nuclide_from_ace((REF + "92233JEF311.ace").encode(), 1)
    assert u, ol.err(orc)
    t, p, d = C.c_double(), C.c_double(), C.c_double()
    orc.orc_ce_nuclide_nubar(u, 1.6, C.byref(t), C.byref(p), C.byref(d)); assert t.value == pytest.approx(2.65431, abs=1e-6)
    orc.orc_ce_nuclide_nubar(u, 17.0, C.byref(t), C.byref(p), C.byref(d)); assert t.value == pytest.approx(5.147534, abs=1e-6) and d.value == pytest.approx(0.0041725, abs=1e-6)
    orc.orc_ce_nuclide_nubar(u, 0.6e-6, C.byref(t), C.byref(p), C.byref(d)); assert p.value == pytest.approx(2.48098, abs=1e-6) and d.value == pytest.approx(6.73e-3, abs=1e-6)
    # the fixture in tests/golden is what this loader produces
    n, rows, m, kT = C.c_int(), C.c_int(), C.c_double(), C.c_double()
    orc.orc_ce_nuclide_info(u, C.byref(n), C.byref(rows), C.byref(m), C.byref(kT))
    g = np.zeros(n.value); dat = np.zeros(n.value * rows.value)
    orc.orc_ce_nuclide_data(u, ol.dp(g), ol.dp(dat))
    gold = ce_util.golden()
    assert np.array_equal(g, gold["grid_92233"]) and np.array_equal(dat.reshape(n.value, rows.value), gold["data_92233"])
    orc.orc_ce_nuclide_free(u)


def test_nuclide_tables_are_consistent():
    gold = ce_util.golden()
    for name in ce_util.NAMES:
        g, d = gold["grid_" + name], gold["data_" + name]
        assert np.all(np.diff(g) >= 0) and g[0] == 1e-11 and g[-1] >= 20.0      # H-1 runs to 150 MeV
        k = 5 if d.shape[1] == 8 else 4
        assert np.array_equal(d[:, 0], ((d[:, 1] + d[:, 2]) + d[:, 3]) + (d[:, 4] if k == 5 else 0.0))    # total rebuilt from the channels (:922-928)
        assert np.all(d >= 0)


def test_oracle_lookups_reproduce_golden_micro_values(orc):
    gold = ce_util.golden()
    for name in ce_util.NAMES:
        g = np.ascontiguousarray(gold["grid_" + name]); d = np.ascontiguousarray(gold["data_" + name])
        h = orc.orc_ce_nuclide_from_arrays(len(g), d.shape[1], ol.dp(g), ol.dp(d))
        for i, E in enumerate(gold["probeE"]):
            mic = np.zeros(8)
            assert orc.orc_ce_nuclide_micro(h, float(E), ol.dp(mic)) == 0, ol.err(orc)
            assert np.array_equal(mic, gold["micro_" + name][i])
        idx, f = C.c_int(), C.c_double()
        top = float(g[-1])
        assert orc.orc_ce_nuclide_search(h, top, C.byref(idx), C.byref(f)) == 0 and idx.value == len(g) - 1 and f.value == 1.0   # top edge -> N-1
        assert orc.orc_ce_nuclide_search(h, 1e-11, C.byref(idx), C.byref(f)) == 0 and idx.value == 1 and f.value == 0.0
        assert orc.orc_ce_nuclide_search(h, 151.0, C.byref(idx), C.byref(f)) != 0                                               # outside: fatalError
        orc.orc_ce_nuclide_free(h)


def test_majorant_bounds_every_material(orc):
    db, nU = ce_util.oracle_db(orc, ce_util.base_nuclides(), ce_util.MATERIALS_5)
    ug = np.zeros(nU); um = np.zeros(nU)
    orc.orc_ce_db_union(db, ol.dp(ug), ol.dp(um))
    assert np.all(np.diff(ug) > 0) and ug[0] == 1e-11 and ug[-1] == 20.0
    E = ce_util.log_uniform(20000, 1e-11, 20.0, 3)
    maj = np.zeros(len(E)); assert orc.orc_ce_db_majorant_n(db, len(E), ol.dp(E), ol.dp(maj)) == 0
    for m in (1, 2, 3):
        mat = np.full(len(E), m, np.int32); tot = np.zeros(len(E))
        assert orc.orc_ce_db_total_n(db, len(E), ol.dp(E), ol.ip(mat), ol.dp(tot)) == 0
        assert np.all(tot <= maj * (1 + 1e-12))
        mac = np.zeros((len(E), 8)); assert orc.orc_ce_db_macro_n(db, len(E), ol.dp(E), ol.ip(mat), ol.dp(mac)) == 0
        np.testing.assert_allclose(mac[:, 0], tot, rtol=1e-13)
    orc.orc_ce_db_free(db)
