"""Oracle MG data against the reference's integration test.

NuclearData/mgNeutronData/baseMgNeutron/Tests/baseMgNeutronDatabase_iTest.f90:55-315
  (materials IntegrationTestFiles/mgMat1, mgMat2; P0 and P1 variants)
"""
import ctypes as C

import numpy as np
import pytest

from tests import oracle_lib as ol
from tests.fixtures import write_mg_deck

TOL = 1e-6


@pytest.mark.parametrize("pn", ["P0", "P1"])
def test_base_mg_database(orc, tmp_path, pn):
    db = orc.orc_mg_load(write_mg_deck(tmp_path, pn).encode(), b"mg")
    assert db, ol.err(orc)
    nm, ng = C.c_int(), C.c_int()
    orc.orc_mg_info(db, C.byref(nm), C.byref(ng))
    assert (nm.value, ng.value) == (2, 4)
    # only material 1 active in the reference test -> majorant(1) = 2.1; here both are active
    assert orc.orc_mg_total(db, 1, 1) == pytest.approx(2.1, abs=TOL)
    assert orc.orc_mg_total(db, 2, 1) == pytest.approx(3.1, abs=TOL)
    assert orc.orc_mg_total(db, 1, 3) == pytest.approx(6.0, abs=TOL)
    assert orc.orc_mg_majorant(db, 1) == pytest.approx(3.1, abs=TOL)
    x = np.zeros(8)
    orc.orc_mg_macro(db, 2, 1, ol.dp(x))
    np.testing.assert_allclose(x, [3.1, 0.0, 1.1, 1.0, 1.0, 2.3, 202.0, 1.0], atol=TOL)
    orc.orc_mg_macro(db, 1, 4, ol.dp(x))
    np.testing.assert_allclose(x, [7.1, 0.0, 3.1, 4.0, 0.0, 0.0, 0.0, 0.0], atol=TOL)
    P0 = np.zeros(16); prod = np.zeros(16); P1 = np.zeros(16); chi = np.zeros(4); nu = np.zeros(4)
    isP1 = orc.orc_mg_matrices(db, 2, ol.dp(P0), ol.dp(prod), ol.dp(P1), ol.dp(chi), ol.dp(nu))
    assert isP1 == (1 if pn == "P1" else 0)
    # file order is Fortran column-major P0(G_out, G_in): the first file row is "from group 1"
    np.testing.assert_allclose(P0[:4], [0.5, 0.3, 0.2, 0.1])
    np.testing.assert_allclose(chi, [0.8, 0.2, 0.0, 0.0])
    if pn == "P1":
        # P1 <- P1 / P0 * 3 where P0 != 0  (multiScatterP1MG_class.f90:117-124)
        assert P1[0] == pytest.approx(-0.1 / 0.5 * 3.0)
        assert P1[5] == pytest.approx(-0.2 / 1.0 * 3.0)
    orc.orc_mg_free(db)


def test_sampling_statistics(orc, tmp_path):
    """multiScatterMG / fissionMG sampleOut follow the tabulated probabilities."""
    db = orc.orc_mg_load(write_mg_deck(tmp_path).encode(), b"mg")
    s = 12345
    mu, phi, g = C.c_double(), C.c_double(), C.c_int()
    cnt = np.zeros(5)
    N = 20000
    for _ in range(N):
        s = orc.orc_mg_sample_scatter(db, 1, 1, s, C.byref(mu), C.byref(phi), C.byref(g))
        cnt[g.value] += 1
        assert -1.0 <= mu.value <= 1.0 and 0.0 <= phi.value <= 2 * np.pi
    p = np.array([0.5, 0.3, 0.2, 0.1]) / 1.1
    np.testing.assert_allclose(cnt[1:] / N, p, atol=4 * np.sqrt(0.25 / N))
    cnt[:] = 0
    for _ in range(N):
        s = orc.orc_mg_sample_fission(db, 2, s, C.byref(mu), C.byref(phi), C.byref(g))
        cnt[g.value] += 1
    np.testing.assert_allclose(cnt[1:] / N, [0.8, 0.2, 0, 0], atol=4 * np.sqrt(0.25 / N))
    orc.orc_mg_free(db)
