"""Randomly generated continuous-energy compositions (seeded) from the five bundled ACE nuclides: 1 - 5 nuclides per material at random
densities, fuel with at least one fissile nuclide, with and without a minimum collision distance, random energy thresholds of the
free-gas treatment.  Void-gap geometry; DT / ST / HT; banks bit-identical to the oracle, tallies to rounding."""
import ctypes as C

import numpy as np
import pytest

import scone_b200
from tests import oracle_lib as ol
from tests.gpu_util import DECK
from tests.test_gpu_ce_transport import oracle_bank
from tests.test_gpu_variants import GEOM_VOID

pytestmark = pytest.mark.gpu
NUC = ["1001.03", "92233.03", "52126.03", "91231.03", "91232.03"]
FISSILE = ["92233.03", "91231.03", "91232.03"]


def composition(rng, need_fissile):
    k = int(rng.integers(1, 6))
    pick = list(rng.choice(NUC, size=k, replace=False))
    if need_fissile and not any(n in FISSILE for n in pick):
        pick.append(FISSILE[int(rng.integers(0, 3))])
    return " ".join("%s %.6E;" % (n, 10.0 ** rng.uniform(-5.0, -1.3)) for n in pick)


@pytest.mark.parametrize("seed", list(range(12)))
def test_random_ce_compositions(orc, seed):
    rng = np.random.default_rng(2000 + seed)
    nd = ("nuclearData { handles { ce { type aceNeutronDatabase; aceLibrary ../../data/ace/aceLib; ures 0; majorant 1; %s} } materials { "
          "fuel { temp %d; composition { %s } } water { temp %d; composition { %s } } } }" % (
              "avgDist 5.0; " if rng.random() < 0.4 else "", int(rng.integers(250, 900)), composition(rng, True), int(rng.integers(250, 900)), composition(rng, False)))
    co = "collisionOperator { neutronCE { type neutronCEstd; energyThreshold %.1f; massThreshold %.2f; minEnergy %.3E; maxEnergy %.1f; } }" % (
        rng.choice([0.0, 50.0, 400.0, 1.0e6]), rng.choice([0.0, 1.0, 150.0, 300.0]), 10.0 ** rng.uniform(-11, -7), rng.choice([20.0, 10.0, 5.0]))
    tally = ("activeTally { f { type collisionClerk; map { type energyMap; grid log; min 1.0E-9; max 20.0; N 12; } response (fl ab); "
             "fl { type fluxResponse; } ab { type macroResponse; MT -21; } } k { type keffImplicitClerk; } }")
    for tracking in ("transportOperatorDT", "transportOperatorST", "transportOperatorHT"):
        ov = "pop 1500; inactive 1; active 2; seed %d; inactiveTally { } transportOperator { type %s; } %s %s %s %s" % (
            seed + 70, tracking, GEOM_VOID % ("fuel", "water"), nd, co, tally)
        orc.orc_set_math_mode(1)
        try:
            e = orc.orc_eigen_load(DECK["ce_pin"].encode(), ov.encode())
            assert e, ol.err(orc)
            assert orc.orc_eigen_init_source(e) == 0, ol.err(orc)
            pp = scone_b200.EigenPhysicsPackage(DECK["ce_pin"], ov, device=0)
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
                if np.isnan(k_o) or gpu_err:
                    assert np.isnan(k_o) and gpu_err, "only one side failed: oracle %r, device %r" % (ol.err(orc), gpu_err)
                    ok = False
                    break
                for a, b, what in zip(pp.bank(), oracle_bank(orc, e), ("r", "dir", "w", "E")):
                    assert np.array_equal(a, b), "bank (%s) differs after cycle %d (%s)" % (what, cyc, tracking)
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
