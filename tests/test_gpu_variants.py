"""Deck features that the benchmark decks do not touch, driven through whole cycles on the GPU against the oracle (bit-identical banks,
tallies to rounding): rotated and translated universes, a cell universe with the overlap check inside a lattice with an offset map,
unstructured space / energy grids, a linear energy grid, material maps with the undefined bin."""
import ctypes as C

import numpy as np
import pytest

import scone_b200
from tests import oracle_lib as ol
from tests.gpu_util import DECK
from tests.test_gpu_ce_transport import oracle_bank as oracle_bank_ce
from tests.test_gpu_eigen import oracle_bank as oracle_bank_mg

pytestmark = pytest.mark.gpu

GEOM = """geometry {
  type geometryStd;
  boundary (1 1 1 1 1 1);
  graph { type %s; }
  surfaces {
    bound { id 1; type box; origin (0.0 0.0 0.0); halfwidth (4.0 4.0 5.0); }
    ball  { id 2; type sphere; origin (0.3 -0.2 0.4); radius 1.1; }
    slab  { id 3; type plane; coeffs (1.0 1.0 0.5 0.7); }
  }
  cells {
    in   { id 1; type simpleCell; surfaces (-2);   filltype mat; material %s; }
    low  { id 2; type simpleCell; surfaces (2 -3); filltype mat; material %s; }
    high { id 3; type simpleCell; surfaces (2 3);  filltype uni; universe 31; }
  }
  universes {
    root  { id 1; type rootUniverse; border 1; fill u<10>; }
    lat   { id 10; type latUniverse; origin (0.0 0.0 0.0); shape (2 2 2); pitch (4.0 4.0 5.0); padMat %s;
            map (31 32 32 31 32 31 31 32); offsetMap (1 0 1 1 1 1 0 1); }
    tilt  { id 31; type pinUniverse; origin (0.4 0.0 0.0); rotation (20.0 30.0 10.0); radii (0.9 1.4 0.0); fills (%s %s %s); }
    cells { id 32; type cellUniverse; origin (0.1 0.2 -0.3); rotation (0.0 45.0 0.0); checkOverlap 1; cells (1 2 3); }
  }
}"""
SPACE = "mz { type spaceMap; axis z; grid unstruct; bins (-5.0 -2.0 -0.5 0.1 0.9 3.3 5.0); }"


def run(orc, deck, ov, ncyc, bank_fn, n_bins):
    orc.orc_set_math_mode(1)
    try:
        e = orc.orc_eigen_load(deck.encode(), ov.encode())
        assert e, ol.err(orc)
        pp = scone_b200.EigenPhysicsPackage(deck, ov, device=0)
        assert orc.orc_eigen_init_source(e) == 0, ol.err(orc)
        pp.generateInitialState()
        k_o = orc.orc_eigen_keff0(e)
        for cyc in range(ncyc):
            pp.cycle(cyc >= 1)
            k_o = orc.orc_eigen_cycle(e, 1 if cyc >= 1 else 0, k_o)
            assert not np.isnan(k_o), ol.err(orc)
            for a, b in zip(pp.bank(), bank_fn(orc, e)):
                assert np.array_equal(a, b), "bank differs after cycle %d" % cyc
        n = orc.orc_eigen_tally_size(e, 1)
        assert n == n_bins
        cs, cs2, nb = pp.tally(True)
        ocs = np.zeros(n); ocs2 = np.zeros(n); b = C.c_int()
        orc.orc_eigen_tally(e, 1, ol.dp(ocs), ol.dp(ocs2), C.byref(b))
        np.testing.assert_allclose(cs, ocs, rtol=1e-10, atol=1e-300)
        np.testing.assert_allclose(cs2, ocs2, rtol=1e-10, atol=1e-300)
        assert np.count_nonzero(cs) > n_bins // 3
        pp.close(); orc.orc_eigen_free(e)
    finally:
        orc.orc_set_math_mode(0)


@pytest.mark.parametrize("graph,tracking", [("shrunk", "transportOperatorDT"), ("extended", "transportOperatorST"), ("shrunk", "transportOperatorHT")])
def test_mg_rotated_universes_and_unstructured_maps(orc, graph, tracking):
    geom = GEOM % (graph, "UO2", "water", "water", "mox87", "GT", "water")
    tally = ("activeTally { f { type collisionClerk; map { type multiMap; maps (mz mm); %s "
             "mm { type materialMap; materials (UO2 mox87); undefBin yes; } } response (fl fi); fl { type fluxResponse; } fi { type macroResponse; MT -6; } } }" % SPACE)
    ov = "pop 5000; inactive 1; active 2; seed 9; inactiveTally { } transportOperator { type %s; } %s %s" % (tracking, geom, tally)
    run(orc, DECK["c5g7"], ov, 3, oracle_bank_mg, 6 * 3 * 2)


@pytest.mark.parametrize("tracking", ["transportOperatorDT", "transportOperatorST"])
def test_ce_rotated_universes_and_energy_grids(orc, tracking):
    geom = GEOM % ("shrunk", "fuel", "water", "water", "fuel", "water", "water")
    tally = ("activeTally { f { type collisionClerk; map { type multiMap; maps (eu mz); "
             "eu { type energyMap; grid unstruct; bins (1.0E-11 1.0E-7 6.25E-7 1.0E-4 0.1 1.0 20.0); } %s } response (fl); fl { type fluxResponse; } } "
             "g { type collisionClerk; map { type energyMap; grid lin; min 0.0; max 10.0; N 25; } response (fl ab); fl { type fluxResponse; } ab { type macroResponse; MT -21; } } }" % SPACE)
    ov = "pop 2500; inactive 1; active 2; seed 10; inactiveTally { } transportOperator { type %s; } %s %s" % (tracking, geom, tally)
    run(orc, DECK["ce_pin"], ov, 3, oracle_bank_ce, 6 * 6 + 25 * 2)


GEOM_PERIODIC = """geometry {
  type geometryStd;
  boundary (%s);
  graph { type shrunk; }
  surfaces {
    bound { id 1; type %s; origin (0.0 0.0 0.0); halfwidth (%s); }
    cx { id 2; type xCylinder; origin (0.0 0.5 -0.5); radius 0.8; }
    cy { id 3; type yCylinder; origin (-1.0 0.0 1.0); radius 0.6; }
    px { id 4; type xPlane; x0 0.9; }
    sq { id 5; type xSquareCylinder; origin (0.0 -1.2 -1.2); halfwidth (0.0 0.5 0.4); }
  }
  cells {
    a { id 1; type simpleCell; surfaces (-2 -4);    filltype mat; material UO2; }
    b { id 2; type simpleCell; surfaces (-3 2);     filltype mat; material mox43; }
    c { id 3; type simpleCell; surfaces (-5 2 3);   filltype mat; material GT; }
    d { id 4; type simpleCell; surfaces (2 3 5);    filltype mat; material water; }
    e { id 5; type simpleCell; surfaces (-2 4 3);   filltype mat; material water; }
  }
  universes {
    root { id 1; type rootUniverse; border 1; fill u<2>; }
    main { id 2; type cellUniverse; cells (1 2 3 4 5); }
  }
}"""


@pytest.mark.parametrize("border,hw,bc,tracking", [
    ("box", "2.0 2.5 2.2", "2 2 1 1 2 2", "transportOperatorDT"),              # periodic in x and z, reflective in y
    ("box", "2.0 2.5 2.2", "2 2 0 1 2 2", "transportOperatorST"),              # one vacuum face
    ("ySquareCylinder", "2.0 0.0 2.2", "2 2 0 0 1 1", "transportOperatorST"),  # infinite in y: periodic x, reflective z
    ("zSquareCylinder", "2.0 2.5 0.0", "1 1 2 2 0 0", "transportOperatorHT")])
def test_mg_axis_cylinders_planes_and_periodic_boundaries(orc, border, hw, bc, tracking):
    geom = GEOM_PERIODIC % (bc, border, hw)
    tally = "activeTally { f { type collisionClerk; map { type spaceMap; axis x; grid lin; min -2.0; max 2.0; N 8; } response (fl); fl { type fluxResponse; } } }"
    ov = "pop 4000; inactive 1; active 2; seed 4; inactiveTally { } transportOperator { type %s; cache %d; } %s %s" % (
        tracking, 0 if tracking == "transportOperatorHT" else 1, geom, tally)
    if tracking == "transportOperatorDT":
        ov = ov.replace("cache 1; ", "")
    run(orc, DECK["c5g7"], ov, 3, oracle_bank_mg, 8)


GEOM_VOID = """geometry {
  type geometryStd;
  boundary (1 1 1 1 0 0);
  graph { type shrunk; }
  surfaces {
    bound { id 1; type box; origin (0.0 0.0 0.0); halfwidth (3.0 3.0 4.0); }
    gapIn  { id 2; type zCylinder; origin (0.0 0.0 0.0); radius 1.0; }
    gapOut { id 3; type zCylinder; origin (0.0 0.0 0.0); radius 1.4; }
  }
  cells {
    fuel { id 1; type simpleCell; surfaces (-2);   filltype mat; material %s; }
    gap  { id 2; type simpleCell; surfaces (2 -3); filltype mat; material void; }
    mod  { id 3; type simpleCell; surfaces (3);    filltype mat; material %s; }
  }
  universes {
    root { id 1; type rootUniverse; border 1; fill u<2>; }
    main { id 2; type cellUniverse; cells (1 2 3); }
  }
}"""
ND_MG = """nuclearData { handles { mg { type baseMgNeutronDatabase; PN P0; avgDist 2.5; } }
  materials { UO2 { temp 300; xsFile ./xs/UO2.xs; composition { } } water { temp 300; xsFile ./xs/moder.xs; composition { } } } }"""
ND_CE = """nuclearData { handles { ce { type aceNeutronDatabase; aceLibrary ../../data/ace/aceLib; ures 0; majorant 1; avgDist 3.0; } }
  materials { fuel  { temp 293; composition { 92233.03 1.5E-4; 52126.03 2.2E-2; 91231.03 5.0E-5; 91232.03 2.0E-6; } }
              water { temp 293; composition { 1001.03 6.67E-2; 52126.03 1.0E-3; } } } }"""
FLUX = "activeTally { f { type collisionClerk; map { type spaceMap; axis z; grid lin; min -4.0; max 4.0; N 8; } response (fl); fl { type fluxResponse; } } }"


@pytest.mark.parametrize("tracking", ["transportOperatorDT", "transportOperatorST", "transportOperatorHT"])
def test_mg_void_gap_and_minimum_collision_distance(orc, tracking):
    """A void gap (no collisions, flux still scored at virtual collisions) and `avgDist` of the database (collisionXS = 1 / avgDist as the
    floor of the tracking cross section, baseMgNeutronDatabase_class.f90:95-133), vacuum top and bottom."""
    ov = "pop 4000; inactive 1; active 2; seed 6; inactiveTally { } transportOperator { type %s; } %s %s %s" % (
        tracking, GEOM_VOID % ("UO2", "water"), ND_MG, FLUX)
    run(orc, DECK["c5g7"], ov, 3, oracle_bank_mg, 8)


@pytest.mark.parametrize("tracking", ["transportOperatorDT", "transportOperatorST", "transportOperatorHT"])
def test_ce_void_gap_and_minimum_collision_distance(orc, tracking):
    ov = "pop 2500; inactive 1; active 2; seed 7; inactiveTally { } transportOperator { type %s; } %s %s %s" % (
        tracking, GEOM_VOID % ("fuel", "water"), ND_CE, FLUX)
    run(orc, DECK["ce_pin"], ov, 3, oracle_bank_ce, 8)


@pytest.mark.parametrize("deck,pop,tracking,bank_fn", [
    ("c5g7", 5000, "transportOperatorDT", oracle_bank_mg), ("c5g7", 4000, "transportOperatorHT", oracle_bank_mg), ("c5g7", 4000, "transportOperatorST", oracle_bank_mg),
    ("ce_pin", 2500, "transportOperatorDT", oracle_bank_ce), ("ce_pin", 2500, "transportOperatorHT", oracle_bank_ce)])
def test_collision_clerk_without_virtual_collisions(orc, deck, pop, tracking, bank_fn):
    """handleVirtual 0 (collisionClerk_class.f90:204-231): virtual collisions are skipped and the flux is w / Sigma_t of the material, next to a
    clerk with the default handling in the same tally; track clerk alongside (scores only on surface-tracking segments)."""
    tally = ("activeTally { real { type collisionClerk; handleVirtual 0; response (fl ab); fl { type fluxResponse; } ab { type macroResponse; MT -21; } } "
             "all { type collisionClerk; response (fl ab); fl { type fluxResponse; } ab { type macroResponse; MT -21; } } "
             "path { type trackClerk; response (fl ab); fl { type fluxResponse; } ab { type macroResponse; MT -21; } } }")
    ov = "pop %d; inactive 1; active 3; seed 8; inactiveTally { } transportOperator { type %s; } %s" % (pop, tracking, tally)
    run(orc, DECK[deck], ov, 4, bank_fn, 6)


@pytest.mark.parametrize("deck,nd,mats,src,tracking", [
    ("c5g7", ND_MG, ("UO2", "water"), "source { type pointSource; r (0.2 0.1 0.0); G 1; }", "transportOperatorST"),
    ("c5g7", ND_MG, ("UO2", "water"), "source { type materialSource; mat water; data mg; G 2; boundingBox (1.5 1.5 -3.9 2.9 2.9 3.9); }", "transportOperatorHT"),
    ("c5g7", ND_MG, ("UO2", "water"), "source { type pointSource; r (1.2 0.0 0.0); G 7; }", "transportOperatorDT"),      # born in the void gap
    ("ce_pin", ND_CE, ("fuel", "water"), "source { type pointSource; r (0.0 0.0 1.0); E 14.1; }", "transportOperatorST"),
    ("ce_pin", ND_CE, ("fuel", "water"), "source { type materialSource; mat fuel; E 2.0; boundingBox (-0.7 -0.7 -3.9 0.7 0.7 3.9); }", "transportOperatorDT")])
def test_fixed_source_with_void_gap(orc, deck, nd, mats, src, tracking):
    """Fixed-source batches (secondaries from the private buffer) in the geometry with a void gap, vacuum ends and a minimum collision
    distance: segment and collision counts equal to the oracle's, tallies to rounding."""
    ov = ("type fixedSourcePhysicsPackage; pop 3000; cycles 2; seed 13; buffer 60; transportOperator { type %s; } %s %s %s "
          "tally { f { type collisionClerk; map { type spaceMap; axis z; grid lin; min -4.0; max 4.0; N 8; } response (fl); fl { type fluxResponse; } } "
          "p { type trackClerk; response (fl); fl { type fluxResponse; } } }" % (tracking, GEOM_VOID % mats, nd, src))
    orc.orc_set_math_mode(1)
    try:
        e = orc.orc_eigen_load(DECK[deck].encode(), ov.encode())
        assert e, ol.err(orc)
        pp = scone_b200.FixedSourcePhysicsPackage(DECK[deck], ov, device=0)
        segs = colls = 0
        for _ in range(2):
            assert orc.orc_fixed_cycle(e) == 0, ol.err(orc)
            res = pp.fixed_cycle()
            segs += res.n_segments; colls += res.n_collisions
            assert pp.rng_state == orc.orc_eigen_rng_state(e)
        seg, coll, hist = C.c_long(), C.c_long(), C.c_long()
        orc.orc_eigen_stats(e, C.byref(seg), C.byref(coll), C.byref(hist))
        assert segs == seg.value and colls == coll.value
        n = orc.orc_eigen_tally_size(e, 1)
        cs, cs2, nb = pp.tally(True)
        ocs = np.zeros(n); ocs2 = np.zeros(n); b = C.c_int()
        orc.orc_eigen_tally(e, 1, ol.dp(ocs), ol.dp(ocs2), C.byref(b))
        np.testing.assert_allclose(cs, ocs, rtol=1e-10, atol=1e-300)
        np.testing.assert_allclose(cs2, ocs2, rtol=1e-10, atol=1e-300)
        assert cs[:8].sum() > 0
        pp.close(); orc.orc_eigen_free(e)
    finally:
        orc.orc_set_math_mode(0)


def test_material_source_point_in_void_is_the_reference_error(orc):
    """materialSource asks the database for the material at every sampled point; a void region has none (materialSource_class.f90:176-177)."""
    ov = ("type fixedSourcePhysicsPackage; pop 500; cycles 1; seed 13; transportOperator { type transportOperatorDT; } %s %s "
          "source { type materialSource; mat water; data mg; G 2; } tally { }" % (GEOM_VOID % ("UO2", "water"), ND_MG))
    e = orc.orc_eigen_load(DECK["c5g7"].encode(), ov.encode())
    assert e, ol.err(orc)
    assert orc.orc_fixed_cycle(e) != 0 and "did not return neutron material" in ol.err(orc)
    orc.orc_eigen_free(e)
    pp = scone_b200.FixedSourcePhysicsPackage(DECK["c5g7"], ov, device=0)
    with pytest.raises(scone_b200.EngineError, match="did not return neutron material"):
        pp.fixed_cycle()
    pp.close()


@pytest.mark.parametrize("src,tracking", [
    ("r (0.5 0.0 -1.0); dir (1.0 0.0 0.0); G 1;", "transportOperatorST"),        # along the axis of the x truncated cylinder
    ("r (3.5 0.0 -1.0); dir (0.0 0.0 1.0); G 3;", "transportOperatorST"),        # born ON the rod surface, moving parallel to it
    ("r (3.5 0.0 -1.0); dir (-1.0 0.0 0.0); G 3;", "transportOperatorHT"),
    ("r (0.5 0.0 3.0); dir (0.0 1.0 0.0); G 2;", "transportOperatorST"),         # born ON the top face of the rod
    ("r (0.5 0.0 -5.999999999999); G 7;", "transportOperatorDT")])               # a hair above the reflective bottom
def test_fixed_source_on_surfaces_and_along_axes(orc, src, tracking):
    """Source particles that start exactly on surfaces or fly along coordinate axes (zero direction components, parallel-to-surface
    branches of `going`, the tolerance branches of `distance`): decks/mg/can as a fixed-source problem."""
    ov = ("type fixedSourcePhysicsPackage; pop 3000; cycles 2; seed 21; transportOperator { type %s; } source { type pointSource; %s } "
          "tally { f { type collisionClerk; map { type spaceMap; axis z; grid lin; min -6.0; max 6.0; N 6; } response (fl); fl { type fluxResponse; } } }" % (tracking, src))
    orc.orc_set_math_mode(1)
    try:
        e = orc.orc_eigen_load(DECK["can"].encode(), ov.encode())
        assert e, ol.err(orc)
        pp = scone_b200.FixedSourcePhysicsPackage(DECK["can"], ov, device=0)
        segs = colls = 0
        for _ in range(2):
            assert orc.orc_fixed_cycle(e) == 0, ol.err(orc)
            res = pp.fixed_cycle()
            segs += res.n_segments; colls += res.n_collisions
        seg, coll, hist = C.c_long(), C.c_long(), C.c_long()
        orc.orc_eigen_stats(e, C.byref(seg), C.byref(coll), C.byref(hist))
        assert segs == seg.value and colls == coll.value
        n = orc.orc_eigen_tally_size(e, 1)
        cs, cs2, nb = pp.tally(True)
        ocs = np.zeros(n); ocs2 = np.zeros(n); b = C.c_int()
        orc.orc_eigen_tally(e, 1, ol.dp(ocs), ol.dp(ocs2), C.byref(b))
        np.testing.assert_allclose(cs, ocs, rtol=1e-10, atol=1e-300)
        assert cs.sum() > 0
        pp.close(); orc.orc_eigen_free(e)
    finally:
        orc.orc_set_math_mode(0)


def test_ce_predefined_energy_grids(orc):
    """energyMap with `grid predef` (energyMap_class.f90:137-177): the named group structures, thermal to fast, in a multi-map with space."""
    tally = ("activeTally { w { type collisionClerk; map { type energyMap; grid predef; name wims69; } response (fl); fl { type fluxResponse; } } "
             "v { type collisionClerk; map { type multiMap; maps (e z); e { type energyMap; grid predef; name casmo7; } %s } response (fl ab); "
             "fl { type fluxResponse; } ab { type macroResponse; MT -21; } } }" % SPACE.replace("mz {", "z {"))
    ov = "pop 2500; inactive 1; active 2; seed 14; inactiveTally { } %s" % tally
    run(orc, DECK["ce_pin"], ov, 3, oracle_bank_ce, 69 + 7 * 6 * 2)
    with pytest.raises(scone_b200.EngineError, match="is undefined"):
        scone_b200.EigenPhysicsPackage(DECK["ce_pin"], "seed 1; activeTally { w { type collisionClerk; map { type energyMap; grid predef; name nogrid; } response (fl); fl { type fluxResponse; } } }", device=-1)
