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


def _one_material_deck(tmp_path, xs_text, pn):
    (tmp_path / "mat").write_text(xs_text)
    deck = tmp_path / "deck"
    deck.write_text("nuclearData { handles { mg { type baseMgNeutronDatabase; PN %s; } } materials { m { temp 1; composition { } xsFile ./mat; } } }" % pn)
    return str(deck)


@pytest.mark.parametrize("pn", ["P0", "P1"])
def test_multi_scatter_known_answers(orc, tmp_path, pn):
    # reactionMG/Tests/multiScatterMG_test.f90:13-14,50-66 and multiScatterP1MG_test.f90:13-15,52-74:
    # P0 = [1.3 0.7 0.3 4.0], scatteringMultiplicity = [1.1 1.05 1 1] (column-major (G_out, G_in)), P1 = [0.5 0.1 0 0]
    xs = """numberOfGroups 2; capture (0.0 0.0);
            P0 (1.3 0.7 0.3 4.0); scatteringMultiplicity (1.1 1.05 1.0 1.0); P1 (0.5 0.1 0.0 0.0);"""
    db = orc.orc_mg_load(_one_material_deck(tmp_path, xs, pn).encode(), b"mg")
    assert db, ol.err(orc)
    x = np.zeros(8)
    orc.orc_mg_macro(db, 1, 1, ol.dp(x)); assert x[2] == pytest.approx(2.0, abs=TOL)          # scatterXS(1)
    orc.orc_mg_macro(db, 1, 2, ol.dp(x)); assert x[2] == pytest.approx(4.3, abs=TOL)          # scatterXS(2)
    P0 = np.zeros(4); prod = np.zeros(4); P1 = np.zeros(4); chi = np.zeros(2); nu = np.zeros(2)
    isP1 = orc.orc_mg_matrices(db, 1, ol.dp(P0), ol.dp(prod), ol.dp(P1), ol.dp(chi), ol.dp(nu))
    P0 = P0.reshape(2, 2).T; prod = prod.reshape(2, 2).T; P1 = P1.reshape(2, 2).T          # [G_out - 1, G_in - 1]
    # production(G_in, G_out): (1,1) = 1.1, (2,1) = 1, (1,2) = 1.05
    assert prod[0, 0] == pytest.approx(1.1, abs=TOL) and prod[0, 1] == pytest.approx(1.0, abs=TOL) and prod[1, 0] == pytest.approx(1.05, abs=TOL)
    # releasePrompt(1) = sum(P0 * prod) / scatterXS = 1.0825; release(2) = 1
    assert (P0[:, 0] * prod[:, 0]).sum() / P0[:, 0].sum() == pytest.approx(1.0825, abs=TOL)
    assert (P0[:, 1] * prod[:, 1]).sum() / P0[:, 1].sum() == pytest.approx(1.0, abs=TOL)
    if pn == "P1":
        assert isP1 == 1
        # P1(G_out, G_in): (2,2) = 0, (1,2) = 0, (1,1) = 1.1538461538, (2,1) = 0.4285714287
        assert P1[1, 1] == pytest.approx(0.0, abs=TOL) and P1[0, 1] == pytest.approx(0.0, abs=TOL)
        assert P1[0, 0] == pytest.approx(1.1538461538, abs=TOL) and P1[1, 0] == pytest.approx(0.4285714287, abs=TOL)
    orc.orc_mg_free(db)


@pytest.mark.parametrize("kappa", [None, (200.0, 203.0, 201.0)])
def test_fission_mg_known_answers(orc, tmp_path, kappa):
    # reactionMG/Tests/fissionMG_test.f90:13-15,48-68: nu = [2.3 2.0 1.3], chi = [0.333333 0.333333 0.333334]; kappa defaults to 202.27 MeV
    xs = """numberOfGroups 3; capture (0.0 0.0 0.0); fission (1.0 1.0 1.0); nu (2.3 2.0 1.3); chi (0.333333 0.333333 0.333334);
            P0 (1 0 0 0 1 0 0 0 1); scatteringMultiplicity (1 1 1 1 1 1 1 1 1);"""
    if kappa:
        xs += " kappa (%s);" % " ".join(map(str, kappa))
    db = orc.orc_mg_load(_one_material_deck(tmp_path, xs, "P0").encode(), b"mg")
    assert db, ol.err(orc)
    x = np.zeros(8)
    for g, nu in ((2, 2.0), (3, 1.3)):                        # release(2), releasePrompt(3) through nuFission = nu * fission
        orc.orc_mg_macro(db, 1, g, ol.dp(x))
        assert x[5] == pytest.approx(nu, abs=TOL) and x[7] == 1.0
    orc.orc_mg_macro(db, 1, 2, ol.dp(x))
    assert x[6] == pytest.approx(203.0 if kappa else 202.27, abs=1e-5)       # getKappa
    P0 = np.zeros(9); prod = np.zeros(9); P1 = np.zeros(9); chi = np.zeros(3); nu = np.zeros(3)
    orc.orc_mg_matrices(db, 1, ol.dp(P0), ol.dp(prod), ol.dp(P1), ol.dp(chi), ol.dp(nu))
    np.testing.assert_allclose(chi, [0.333333, 0.333333, 0.333334]); np.testing.assert_allclose(nu, [2.3, 2.0, 1.3])
    orc.orc_mg_free(db)
