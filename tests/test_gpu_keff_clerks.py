"""k-eff clerks named by the deck itself (keffAnalogClerk / keffImplicitClerk inside inactiveTally / activeTally / tally, as most of
the reference's input files have them) next to collision clerks: memory layout, normalisation and accumulated k against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import scone_b200
from tests import oracle_lib as ol
from tests.gpu_util import DECK

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TALLY = ("%s { norm fiss; normVal 100; k_ana { type keffAnalogClerk; } "
         "fiss { type collisionClerk; response (fiss); fiss { type macroResponse; MT -6; } } k_imp { type keffImplicitClerk; } }")


@pytest.mark.parametrize("deck,pop,extra", [
    (DECK["c5g7"], 6000, ""), (DECK["ce_pin"], 2500, ""),
    (DECK["slab"], 4000, " transportOperator { type transportOperatorST; }")])
def test_user_keff_clerks_eigen(orc, deck, pop, extra):
    ov = "pop %d; inactive 2; active 3; seed 21; %s %s%s" % (pop, TALLY % "inactiveTally", TALLY % "activeTally", extra)
    orc.orc_set_math_mode(1)
    try:
        e = orc.orc_eigen_load(deck.encode(), ov.encode())
        assert e, ol.err(orc)
        pp = scone_b200.EigenPhysicsPackage(deck, ov, device=0)
        orc.orc_eigen_init_source(e); pp.generateInitialState()
        k_o = orc.orc_eigen_keff0(e)
        for cyc in range(5):
            pp.cycle(cyc >= 2)
            k_o = orc.orc_eigen_cycle(e, 1 if cyc >= 2 else 0, k_o)
            assert pp.k == pytest.approx(k_o, rel=1e-11)
        for phase, nb_expected in ((0, 2), (1, 3)):
            n = orc.orc_eigen_tally_size(e, phase)
            assert n == 3 + 1 + 5
            cs, cs2, nb = pp.tally(bool(phase))
            ocs = np.zeros(n); ocs2 = np.zeros(n); b = C.c_int()
            orc.orc_eigen_tally(e, phase, ol.dp(ocs), ol.dp(ocs2), C.byref(b))
            assert nb == b.value == nb_expected
            np.testing.assert_allclose(cs, ocs, rtol=1e-10, atol=1e-300)
            np.testing.assert_allclose(cs2, ocs2, rtol=1e-10, atol=1e-300)
            assert cs[2] > 0 and cs[8] > 0 and cs[3] == pytest.approx(100.0 * nb)      # k bins filled, norm applied to the fission bin
        pp.close(); orc.orc_eigen_free(e)
    finally:
        orc.orc_set_math_mode(0)


def test_user_keff_implicit_clerk_fixed_source(orc):
    """InputFiles/sphere_with_DT has a keffImplicitClerk in its fixed-source tally: leakage of every secondary counts."""
    deck = os.path.join(ROOT, "decks", "fixed", "ce_sphere")
    ov = "pop 5000; cycles 3; seed 4; tally { k_eff { type keffImplicitClerk; } fiss { type collisionClerk; response (fiss); fiss { type macroResponse; MT -6; } } }"
    orc.orc_set_math_mode(1)
    try:
        e = orc.orc_eigen_load(deck.encode(), ov.encode())
        assert e, ol.err(orc)
        pp = scone_b200.FixedSourcePhysicsPackage(deck, ov, device=0)
        for _ in range(3):
            assert orc.orc_fixed_cycle(e) == 0, ol.err(orc)
            pp.fixed_cycle()
        n = orc.orc_eigen_tally_size(e, 1)
        assert n == 6
        cs, cs2, nb = pp.tally(True)
        ocs = np.zeros(n); ocs2 = np.zeros(n); b = C.c_int()
        orc.orc_eigen_tally(e, 1, ol.dp(ocs), ol.dp(ocs2), C.byref(b))
        np.testing.assert_allclose(cs, ocs, rtol=1e-10, atol=1e-300)
        np.testing.assert_allclose(cs2, ocs2, rtol=1e-10, atol=1e-300)
        assert 0.0 < cs[4] / nb < 1.0 and cs[3] > 0          # subcritical k estimate, leakage scored
        pp.close(); orc.orc_eigen_free(e)
    finally:
        orc.orc_set_math_mode(0)


TRACK = ("activeTally { flxT { type trackClerk; map { type spaceMap; axis x; grid lin; min %g; max %g; N 12; } response (flux fis); flux { type fluxResponse; } fis { type macroResponse; MT -6; } } "
         "flxC { type collisionClerk; map { type spaceMap; axis x; grid lin; min %g; max %g; N 12; } response (flux); flux { type fluxResponse; } } }")


@pytest.mark.parametrize("deck,pop,lo,hi,tracking", [
    (DECK["c5g7"], 5000, -32.13, 32.13, "transportOperator { type transportOperatorST; }"),
    (DECK["c5g7"], 5000, -32.13, 32.13, "transportOperator { type transportOperatorHT; cutoff 0.6; }"),
    (DECK["ce_pin"], 2500, -0.63, 0.63, "transportOperator { type transportOperatorST; cache 0; }"),
    (DECK["slab"], 4000, -9.4959, 9.4959, "transportOperator { type transportOperatorDT; }")])
def test_track_clerk_path_length_estimator(orc, deck, pop, lo, hi, tracking):
    """trackClerk (path-length flux) next to a collisionClerk on the same mesh: every bin against the oracle; the two flux estimators
    agree statistically; under delta tracking the track clerk stays empty, as in the reference (no path reports)."""
    ov = "pop %d; inactive 1; active 3; seed 8; %s %s" % (pop, TRACK % (lo, hi, lo, hi), tracking)
    orc.orc_set_math_mode(1)
    try:
        e = orc.orc_eigen_load(deck.encode(), ov.encode())
        assert e, ol.err(orc)
        pp = scone_b200.EigenPhysicsPackage(deck, ov, device=0)
        orc.orc_eigen_init_source(e); pp.generateInitialState()
        k_o = orc.orc_eigen_keff0(e)
        for cyc in range(4):
            pp.cycle(cyc >= 1)
            k_o = orc.orc_eigen_cycle(e, 1 if cyc >= 1 else 0, k_o)
        n = orc.orc_eigen_tally_size(e, 1)
        assert n == 24 + 12
        cs, cs2, nb = pp.tally(True)
        ocs = np.zeros(n); ocs2 = np.zeros(n); b = C.c_int()
        orc.orc_eigen_tally(e, 1, ol.dp(ocs), ol.dp(ocs2), C.byref(b))
        np.testing.assert_allclose(cs, ocs, rtol=1e-10, atol=1e-300)
        np.testing.assert_allclose(cs2, ocs2, rtol=1e-10, atol=1e-300)
        fluxT, fluxC = cs[0:24:2].sum(), cs[24:].sum()
        if "DT" in tracking:
            assert fluxT == 0.0 and fluxC > 0.0
        else:
            assert fluxT > 0.0 and (abs(fluxT / fluxC - 1.0) < 0.1 or "HT" in tracking)
        pp.close(); orc.orc_eigen_free(e)
    finally:
        orc.orc_set_math_mode(0)


ENTROPY = ("%s { entropy { type shannonEntropyClerk; cycles %d; map { type multiMap; maps (mx my); "
           "mx { type spaceMap; axis x; grid lin; min %g; max %g; N 8; } my { type spaceMap; axis y; grid lin; min %g; max %g; N 6; } } } "
           "fiss { type collisionClerk; response (fiss); fiss { type macroResponse; MT -6; } } }")


@pytest.mark.parametrize("deck,pop,lo,hi", [(DECK["c5g7"], 8000, -32.13, 32.13), (DECK["ce_pin"], 3000, -0.63, 0.63)])
def test_shannon_entropy_clerk_eigen(orc, deck, pop, lo, hi):
    """shannonEntropyClerk (7 of the reference's input files): the weights of every cycle's fission bank, taken before normSize_Repr,
    binned over a 2-D space map; entropy per cycle in its own bin; cycles beyond `cycles` are not scored (3 active cycles, 2 scored)."""
    ov = "pop %d; inactive 3; active 3; seed 77; %s %s" % (pop, ENTROPY % ("inactiveTally", 3, lo, hi, lo, hi), ENTROPY % ("activeTally", 2, lo, hi, lo, hi))
    orc.orc_set_math_mode(1)
    try:
        e = orc.orc_eigen_load(deck.encode(), ov.encode())
        assert e, ol.err(orc)
        pp = scone_b200.EigenPhysicsPackage(deck, ov, device=0)
        orc.orc_eigen_init_source(e); pp.generateInitialState()
        k_o = orc.orc_eigen_keff0(e)
        for cyc in range(6):
            pp.cycle(cyc >= 3)
            k_o = orc.orc_eigen_cycle(e, 1 if cyc >= 3 else 0, k_o)
        for phase, cycles in ((0, 3), (1, 2)):
            n = orc.orc_eigen_tally_size(e, phase)
            assert n == 48 + 1 + cycles + 1
            cs, cs2, nb = pp.tally(bool(phase))
            ocs = np.zeros(n); ocs2 = np.zeros(n); b = C.c_int()
            orc.orc_eigen_tally(e, phase, ol.dp(ocs), ol.dp(ocs2), C.byref(b))
            assert len(cs) == n
            np.testing.assert_allclose(cs, ocs, rtol=1e-10, atol=1e-14)
            np.testing.assert_allclose(cs2, ocs2, rtol=1e-10, atol=1e-14)
            ent = cs[49:49 + cycles]
            assert (cs[:49] == 0).all()                              # the weight bins are reset every cycle
            assert (ent > 0.5).all() and (ent < np.log2(48) + 1e-9).all()
            assert cs[-1] > 0                                        # the collision clerk behind the entropy clerk keeps its address
        pp.close(); orc.orc_eigen_free(e)
    finally:
        orc.orc_set_math_mode(0)


ALL_RESP = ("%s { a { type collisionClerk; response (r1 r2 r22 r4 r6 r7 r80 r21); %s } "
            "b { type collisionClerk; map { type materialMap; materials (%s); undefBin yes; } response (fl r3 r20 r8 r9 m18 m101); fl { type fluxResponse; } %s } }")


@pytest.mark.parametrize("deck,pop,mats,extra", [
    (DECK["c5g7"], 5000, "UO2 water", ""), (DECK["c5g7"], 4000, "UO2 water", " transportOperator { type transportOperatorST; }"),
    (DECK["ce_pin"], 2500, "fuel water", ""), (DECK["ce_pin"], 2500, "fuel water", " transportOperator { type transportOperatorDT; }")])
def test_every_macro_response_against_oracle(orc, deck, pop, mats, extra):
    """macroResponse over the whole table of neutronMacroXSs%get (neutronXsPackages_class.f90:143-190): total, capture, elastic, non-elastic,
    inelastic, all scattering, fission, nu-fission, kappa-fission, prompt / delayed nu-fission, absorption, and ENDF MT numbers (18, 101)."""
    def defs(mts, pre="r"):
        return " ".join("%s%d { type macroResponse; MT %d; }" % (pre, abs(m), m) for m in mts)
    t1 = defs([-1, -2, -22, -4, -6, -7, -80, -21])
    t2 = defs([-3, -20, -8, -9]) + " " + defs([18, 101], "m")
    ov = "pop %d; inactive 1; active 3; seed 5; inactiveTally { } %s%s" % (pop, ALL_RESP % ("activeTally", t1, mats, t2), extra)
    orc.orc_set_math_mode(1)
    try:
        e = orc.orc_eigen_load(deck.encode(), ov.encode())
        assert e, ol.err(orc)
        pp = scone_b200.EigenPhysicsPackage(deck, ov, device=0)
        orc.orc_eigen_init_source(e); pp.generateInitialState()
        k_o = orc.orc_eigen_keff0(e)
        for cyc in range(4):
            pp.cycle(cyc >= 1)
            k_o = orc.orc_eigen_cycle(e, 1 if cyc >= 1 else 0, k_o)
        n = orc.orc_eigen_tally_size(e, 1)
        assert n == 8 + 3 * 7
        cs, cs2, nb = pp.tally(True)
        ocs = np.zeros(n); ocs2 = np.zeros(n); b = C.c_int()
        orc.orc_eigen_tally(e, 1, ol.dp(ocs), ol.dp(ocs2), C.byref(b))
        np.testing.assert_allclose(cs, ocs, rtol=1e-10, atol=1e-300)
        np.testing.assert_allclose(cs2, ocs2, rtol=1e-10, atol=1e-300)
        a = cs[:8]
        assert (a[[0, 1, 2, 4, 5, 6, 7]] > 0).all()
        assert a[7] == pytest.approx(a[1] + a[4], rel=1e-9)            # absorption = capture + fission
        assert sum(cs[8 + 7 * b + 5] for b in range(3)) == pytest.approx(a[4], rel=1e-9)      # MT 18 over the material bins = fission
        assert sum(cs[8 + 7 * b + 6] for b in range(3)) == pytest.approx(a[1], rel=1e-9)      # MT 101 = capture
        pp.close(); orc.orc_eigen_free(e)
    finally:
        orc.orc_set_math_mode(0)
