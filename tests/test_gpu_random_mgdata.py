"""Randomly generated multigroup data (seeded): 1 - 9 groups, dense or sparse scattering matrices with up- and down-scattering,
(n,xn) multiplicities, P1 moments of either sign, materials without fission or without scattering in some groups, kappa given or not.
Two materials in the void-gap geometry; whole cycles under DT / ST / HT, banks bit-identical to the oracle, tallies to rounding."""
import ctypes as C

import numpy as np
import pytest

import scone_b200
from tests import oracle_lib as ol
from tests.gpu_util import DECK
from tests.test_gpu_eigen import oracle_bank
from tests.test_gpu_variants import GEOM_VOID

pytestmark = pytest.mark.gpu


def fmt(a):
    return " ".join("%.9E" % v for v in np.asarray(a).ravel())


def gen_material(rng, nG, fissile, p1):
    P0 = rng.uniform(0.0, 0.4, size=(nG, nG))                     # rows = from-group (file order)
    if rng.random() < 0.5:
        P0 *= rng.random(size=(nG, nG)) < 0.5                     # sparse
    for g in range(nG):
        if P0[g].sum() == 0.0 and rng.random() < 0.7:
            P0[g, g] = rng.uniform(0.05, 0.4)                     # most groups scatter somewhere; a few do not at all
    mult = np.where(rng.random(size=(nG, nG)) < 0.15, rng.uniform(1.0, 2.0, size=(nG, nG)), 1.0)
    cap = rng.uniform(0.01, 0.2, size=nG)
    text = "numberOfGroups %d; capture (%s); scatteringMultiplicity (%s); P0 (%s);" % (nG, fmt(cap), fmt(mult), fmt(P0))
    if p1:
        P1 = P0 * rng.uniform(-0.3, 0.3, size=(nG, nG))           # |P1| < P0 / 3 keeps the linear pdf positive
        text += " P1 (%s);" % fmt(P1)
    if fissile:
        fis = rng.uniform(0.0, 0.12, size=nG) * (rng.random(size=nG) < 0.8)
        nu = rng.uniform(2.0, 3.0, size=nG)
        chi = rng.random(size=nG) * (rng.random(size=nG) < 0.7); chi[0] += 0.1; chi /= chi.sum()
        text += " fission (%s); nu (%s); chi (%s);" % (fmt(fis), fmt(nu), fmt(chi))
        if rng.random() < 0.5:
            text += " kappa (%s);" % fmt(rng.uniform(190.0, 210.0, size=nG))
    return text


@pytest.mark.parametrize("seed", list(range(16)))
def test_random_mg_data(orc, tmp_path, seed):
    rng = np.random.default_rng(1000 + seed)
    nG = int(rng.integers(1, 10))
    p1 = bool(rng.random() < 0.5)
    (tmp_path / "fuel.xs").write_text(gen_material(rng, nG, True, p1))
    (tmp_path / "mod.xs").write_text(gen_material(rng, nG, bool(rng.random() < 0.3), p1))
    nd = ("nuclearData { handles { mg { type baseMgNeutronDatabase; PN %s; %s} } materials { "
          "UO2 { temp 300; xsFile %s; composition { } } water { temp 300; xsFile %s; composition { } } } }" % (
              "P1" if p1 else "P0", "avgDist 4.0; " if rng.random() < 0.4 else "", tmp_path / "fuel.xs", tmp_path / "mod.xs"))
    tally = ("activeTally { f { type collisionClerk; map { type spaceMap; axis z; grid lin; min -4.0; max 4.0; N 4; } response (fl fi sc); "
             "fl { type fluxResponse; } fi { type macroResponse; MT -7; } sc { type macroResponse; MT -4; } } k { type keffImplicitClerk; } }")
    for tracking in ("transportOperatorDT", "transportOperatorST", "transportOperatorHT"):
        ov = "pop 3000; inactive 1; active 2; seed %d; inactiveTally { } transportOperator { type %s; } %s %s %s" % (
            seed + 50, tracking, GEOM_VOID % ("UO2", "water"), nd, tally)
        orc.orc_set_math_mode(1)
        try:
            e = orc.orc_eigen_load(DECK["c5g7"].encode(), ov.encode())
            assert e, ol.err(orc)
            assert orc.orc_eigen_init_source(e) == 0, ol.err(orc)
            pp = scone_b200.EigenPhysicsPackage(DECK["c5g7"], ov, device=0)
            pp.generateInitialState()
            for a, b in zip(pp.bank(), oracle_bank(orc, e)):
                assert np.array_equal(a, b)
            k_o = orc.orc_eigen_keff0(e)
            ok = True
            for cyc in range(3):
                try:
                    pp.cycle(cyc >= 1); gpu_err = None
                except scone_b200.EngineError as ex:
                    gpu_err = str(ex)
                k_o = orc.orc_eigen_cycle(e, 1 if cyc >= 1 else 0, k_o)
                if np.isnan(k_o) or gpu_err:                  # bank died out or overflowed: both sides must stop
                    assert np.isnan(k_o) and gpu_err, "only one side failed: oracle %r, device %r" % (ol.err(orc), gpu_err)
                    ok = False
                    break
                for a, b in zip(pp.bank(), oracle_bank(orc, e)):
                    assert np.array_equal(a, b), "bank differs after cycle %d (%s, %d groups, P1 %s)" % (cyc, tracking, nG, p1)
                assert pp.k == pytest.approx(k_o, rel=1e-11)
            if ok:
                n = orc.orc_eigen_tally_size(e, 1)
                cs, cs2, nb = pp.tally(True)
                ocs = np.zeros(n); ocs2 = np.zeros(n); b = C.c_int()
                orc.orc_eigen_tally(e, 1, ol.dp(ocs), ol.dp(ocs2), C.byref(b))
                np.testing.assert_allclose(cs, ocs, rtol=1e-10, atol=1e-300)
            pp.close(); orc.orc_eigen_free(e)
        finally:
            orc.orc_set_math_mode(0)
