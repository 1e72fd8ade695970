"""Oracle geometry against the reference's own unit / integration test vectors.

Geometry/Tests/geometryStd_iTest.f90:22-296        (test_lat, test_cyl decks)
Geometry/Universes/Tests/latUniverse_test.f90:15-440
Geometry/Universes/Tests/pinUniverse_test.f90:18-285
Geometry/Tests/geomGraph_test.f90 is covered through the graph arrays of test_lat.
"""
import numpy as np
import pytest

from tests import oracle_lib as ol
from tests.fixtures import TEST_CYL, TEST_LAT

INF = 2.0 ** 63
SURF_TOL = 1.0e-12
COLL_EV, BOUNDARY_EV, CROSS_EV = 1, 2, 3
SQRT2 = np.sqrt(2.0)


def unit(v):
    v = np.asarray(v, float)
    return v / np.sqrt((v * v).sum())


# --------------------------------------------------------------------------- geometryStd_iTest
def test_lattice_geom(orc):
    g = ol.Geom(orc, TEST_LAT)
    mats = dict(water=1, mox43=2, uox=3)
    assert g.what_is_at([0, 0, 0])[0] == mats["water"]
    assert g.what_is_at([0.63, -0.09, 0.0], [0, -1, 0])[0] == mats["mox43"]

    r = np.array([0.1, 0.1, 0.0]); u = np.array([0.0, 0.0, 1.0])
    g.init(r, u); g.place()
    c = g.get()
    np.testing.assert_allclose(c["r"][0], r, atol=1e-7)
    np.testing.assert_allclose(c["r"][1], r, atol=1e-7)
    np.testing.assert_allclose(c["r"][2], r - [0.63, 0.63, 0.0], atol=1e-7)
    for l in range(3):
        np.testing.assert_allclose(c["dir"][l], u, atol=1e-7)

    img = g.slice_plot((10, 10), [0, 0, 0], "z", "material")
    assert img[0, 0] == mats["water"] and img[1, 5] == mats["water"]
    assert img[2, 6] == mats["mox43"]
    assert img[2, 2] == mats["uox"]
    img = g.slice_plot((10, 10), [-0.63, -0.63, 0.0], "z", "uniqueID", width=[1.26, 1.26])
    assert img[4, 4] == 2 and img[0, 0] == 3

    np.testing.assert_allclose(g.bounds(), [-1.26, -1.26, 0.0, 1.26, 1.26, 0.0], atol=1e-7)

    # teleport with reflective (x) and periodic (y) BCs
    g.init([0, 0, 0], unit([-1, -2, 0]))
    g.teleport(3.0)
    c = g.get()
    np.testing.assert_allclose(c["r"][0], [-1.1783592, -0.1632816, 0.0], atol=1e-7)
    np.testing.assert_allclose(c["dir"][0], unit([1, -2, 0]), atol=1e-7)
    assert c["mat"] == mats["water"]

    # global movement
    g.init([0, 0, 0], [0, -1, 0])
    md, ev = g.move_global(1.0)
    c = g.get()
    np.testing.assert_allclose(c["r"][0], [0, -1.0, 0], atol=1e-7)
    assert ev == COLL_EV and c["mat"] == mats["water"] and md == pytest.approx(1.0, abs=1e-7)
    md, ev = g.move_global(1.0)
    c = g.get()
    np.testing.assert_allclose(c["r"][0], [0, 1.26, 0], atol=1e-7)
    np.testing.assert_allclose(c["dir"][0], [0, -1, 0], atol=1e-7)
    assert ev == BOUNDARY_EV and c["mat"] == mats["water"] and md == pytest.approx(0.26, abs=1e-7)

    # normal movement
    for cache in (False, True):
        g.init([-0.63, -0.63, 0.0], [0, -1, 0]); g.place()
        md, ev = g.move(1.0, cache)
        c = g.get()
        np.testing.assert_allclose(c["r"][0], [-0.63, -1.13, 0], atol=1e-7)
        assert ev == CROSS_EV and c["mat"] == mats["water"] and md == pytest.approx(0.5, abs=1e-7)
        md, ev = g.move(1.0, cache)
        c = g.get()
        np.testing.assert_allclose(c["r"][0], [-0.63, 1.26, 0], atol=1e-7)
        assert ev == BOUNDARY_EV and c["mat"] == mats["water"] and md == pytest.approx(0.13, abs=1e-7)
        md, ev = g.move(0.08, cache)
        c = g.get()
        np.testing.assert_allclose(c["r"][0], [-0.63, 1.18, 0], atol=1e-7)
        assert ev == COLL_EV and c["mat"] == mats["water"] and md == pytest.approx(0.08, abs=1e-7)


def test_tilted_cylinder(orc):
    g = ol.Geom(orc, TEST_CYL)
    W, F = 1, 2
    img = g.slice_plot((20, 20), [1.0, 0.0, 0.0], "x", "material")
    assert img[7, 10] == W and img[16, 2] == W and img[9, 9] == F and img[17, 0] == F
    img = g.slice_plot((20, 20), [0.0, 3.0, 0.0], "y", "material")
    assert img[14, 0] == W and img[12, 3] == W and img[12, 2] == F and img[13, 1] == F
    # small box entirely inside fuel (voxelPlot with width 0.5 around (1,0,0))
    for dx in (-0.2, 0.0, 0.2):
        for dy in (-0.2, 0.2):
            for dz in (-0.2, 0.2):
                assert g.what_is_at([1.0 + dx, dy, dz])[0] == F
    info = g.info()
    assert info["nUni"] == 3 and info["nesting"] == 2


def test_shrunk_graph_layout(orc):
    g = ol.Geom(orc, TEST_LAT)
    idx, gid = g.graph()
    # root(2) + lat10(5) + pin31(2) + sqPin(4): universes laid out once each (shrunk)
    assert len(idx) == 2 + 5 + 2 + 4
    assert idx[0] < 0 and gid[0] == 3          # root inside -> lattice at location 3
    assert idx[1] == 0                          # outside
    assert g.info()["uniqueCells"] == 1 + 2 + 4   # padMat + pin cells + cellUniverse cells (incl. undef/overlap)


# --------------------------------------------------------------------------- latUniverse_test
UNI1_DEF = """id 1; type latUniverse; origin (0.0 0.0 0.0); rotation (0.0 0.0 0.0);
pitch (1.0 2.0 3.0); shape (3 2 2); padMat void;
 map ( 3 4 5
      7 4 8

      1 2 3
      4 5 6);
 offsetMap ( 1 1 1
            0 1 1
            1 1 1
            1 1 1 ); """
UNI2_DEF = "id 2; type latUniverse; pitch (1.0 2.0 0.0); shape (2 1 0); padMat u<1>; map (1 2); "


@pytest.fixture()
def lat(orc):
    u1 = ol.Uni(orc, UNI1_DEF, {"void": 3}, 8)
    u2 = ol.Uni(orc, UNI2_DEF, {"void": 3}, 3)
    return u1, u2


def test_lat_fill(lat):
    u1, u2 = lat
    assert u1.fill() == [-4, -5, -6, -1, -2, -3, -7, -4, -8, -3, -4, -5, 3]
    assert u2.fill() == [-1, -2, -1]


def test_lat_enter(lat):
    u1, u2 = lat
    e = u1.enter([1.0, 1.0, 0.5], [0, 0, 1]); assert (e["uniIdx"], e["localID"], e["cellIdx"]) == (8, 12, 0)
    e = u1.enter([1.6, 0.5, 0.5], [0, 0, 1]); assert e["localID"] == 13
    e = u1.enter([-0.5, 0.0, 0.0], unit([-1, 1, -1])); assert e["localID"] == 4
    e = u2.enter([0.5, 0.5, 13.5], [0, 0, 1]); assert (e["uniIdx"], e["localID"]) == (3, 2)
    e = u2.enter([1.6, 0.5, 0.5], [0, 0, 1]); assert e["localID"] == 3
    e = u2.enter([0.0, 0.0, 0.0], unit([-1, 1, -1])); assert e["localID"] == 1


def test_lat_distance(lat):
    u1, u2 = lat
    d, s = u1.distance(11, [0.0, 0.1, 0.5], [0.0, -0.0, 1.0]); assert d == pytest.approx(2.5, rel=1e-7) and s == -6
    d, s = u1.distance(13, [-4.0, 0.1, 0.5], unit([1, 1, 0])); assert d == INF and s == -7
    eps = 0.5 * SURF_TOL
    d, s = u1.distance(4, [-1.0, 0.0 - eps, -0.5], unit([1, 1, 0])); assert d == pytest.approx(SQRT2 * 0.5, rel=1e-7) and s == -2
    d, s = u1.distance(4, [-0.5 + eps, 0.0 + eps, -0.5], unit([1, 1, 0])); assert d == pytest.approx(0.0, abs=1e-7) and s == -2
    d, s = u2.distance(2, [0.5, 0.6, 0.5], [0, 0, 1]); assert d == INF
    d, s = u2.distance(2, [0.5, 0.6, 0.5], unit([0, 0.01, 1])); assert d == pytest.approx(np.sqrt(40.0 ** 2 + 0.4 ** 2), rel=1e-7) and s == -4
    d, s = u2.distance(3, [-1.5, 0.6, 0.5], unit([1, 0, 1])); assert d == pytest.approx(0.5 * SQRT2, rel=1e-7) and s == -7


def test_lat_cross(lat):
    u1, u2 = lat
    assert u1.cross(1, [-1.0, 0.0, -0.5], unit([-1, 1, -1]), -4) == 4
    assert u1.cross(13, [1.0, 2.0, -0.5], unit([1, -1, -1]), -7) == 6
    assert u1.cross(6, [1.5, 1.0, -1.0], [1, 0, 0], -2) == 13
    assert u2.cross(1, [0.0, 0.0, 16.5], unit([1, 1, -1]), -2) == 2
    assert u2.cross(3, [-1.0, -0.5, -78.5], unit([1, 1, 0]), -7) == 1


def test_lat_offset(lat):
    u1, u2 = lat
    np.testing.assert_allclose(u1.offset(11), [0.0, 1.0, 1.5], atol=1e-7)
    np.testing.assert_allclose(u1.offset(7), [0, 0, 0], atol=1e-7)      # offsetMap entry 0
    np.testing.assert_allclose(u1.offset(13), [0, 0, 0], atol=1e-7)
    np.testing.assert_allclose(u2.offset(2), [0.5, 0.0, 0.0], atol=1e-7)
    np.testing.assert_allclose(u2.offset(3), [0, 0, 0], atol=1e-7)


# --------------------------------------------------------------------------- pinUniverse_test
PIN_DEF = "id 7; type pinUniverse; origin (0.0 0.0 0.0); rotation (0.0 0.0 0.0); radii (2.5 1.5 0.0); fills (u<7> u<14> void);"
MOVING_IN, MOVING_OUT = -1, -2


@pytest.fixture()
def pin(orc):
    return ol.Uni(orc, PIN_DEF, {"void": 13}, 3)


def test_pin_fill_and_enter(pin):
    assert pin.fill() == [-14, -7, 13]
    e = pin.enter([0.0, 1.0, 0.0], [0, 0, 1]); assert (e["uniIdx"], e["localID"], e["cellIdx"]) == (3, 1, 0)
    e = pin.enter([2.3, 0.0, -980.0], [0, 0, 1]); assert e["localID"] == 2
    e = pin.enter([2.6, 0.0, -980.0], [0, 0, 1]); assert e["localID"] == 3


def test_pin_distance_cross(pin):
    d, s = pin.distance(1, [1.0, 0.0, 0.0], [1, 0, 0]); assert d == pytest.approx(0.5, rel=1e-7) and s == MOVING_OUT
    d, s = pin.distance(3, [2.0, 1.6, 0.0], [1, 0, 0]); assert d == INF
    d, s = pin.distance(2, [0.0, 1.6, 0.0], [0, -1, 0]); assert d == pytest.approx(0.1, rel=1e-7) and s == MOVING_IN
    eps = 0.5 * SURF_TOL
    assert pin.cross(1, [0.0, 1.5 - eps, 0.0], [0, 1, 0], MOVING_OUT) == 2
    assert pin.cross(2, [0.0, 1.5 + eps, 0.0], [0, -1, 0], MOVING_IN) == 1
    np.testing.assert_array_equal(pin.offset(1), [0, 0, 0])


def test_pin_edge_cases(pin):
    eps = 0.5 * SURF_TOL
    r = [0.0, 1.5 - eps, 0.0]; u = unit([1, -0.00001, 0])
    e = pin.enter(r, u)
    assert e["localID"] == 1
    d, s = pin.distance(1, r, u)
    assert d == pytest.approx(0.0, abs=1e-3) and s == MOVING_OUT


# --------------------------------------------------------------------------- cellUniverse_test
CELL_ENV = """surfaces { surf1 { id 1; type sphere; origin (0.0 0.0 0.0); radius 2;} surf2 { id 2; type sphere; origin (4.0 0.0 0.0); radius 1;} }
cells { cell1 {id 1; type simpleCell; surfaces (-1); filltype uni; universe 3;}
        cell2 {id 2; type simpleCell; surfaces (1 2); filltype uni; universe 4;} }"""
CELL_UNI = "id 1; type cellUniverse; origin (2.0 0.0 0.0); rotation (90.0 90.0 90.0); cells (1 2);"
CELL2_ENV = """surfaces { surf1 { id 1; type sphere; origin (21.0 0.0 0.0); radius 4.5;} surf2 { id 2; type sphere; origin (0.0 0.0 0.0); radius 24.1;} }
cells { cell1 {id 1; type simpleCell; surfaces (-1); filltype uni; universe 3;}
        cell2 {id 2; type simpleCell; surfaces (1 -2); filltype uni; universe 4;}
        cell3 {id 3; type simpleCell; surfaces (2); filltype uni; universe 5;} }"""
CELL2_UNI = "id 2; type cellUniverse; origin (0.0 0.0 0.0); checkOverlap 1; cells (1 2 3);"
VOID_MAT = 2147483647                   # universalVariables.f90: huge(shortInt)
UNDEF_MAT, OVERLAP_MAT = VOID_MAT - 1, VOID_MAT - 2


def test_cell_universe(orc):
    # cellUniverse_test.f90:17-36 definitions (rotation maps x -> z, y -> -y, z -> x), :98-99 fill, test_enter :142-199,
    # test_distance :204-246, test_cross :251-268, test_cellOffset :273-296
    u = ol.Uni(orc, CELL_UNI, {}, 8, env=CELL_ENV)
    assert u.fill() == [-3, -4, UNDEF_MAT, OVERLAP_MAT]
    e = u.enter([0.0, 0.0, 3.0], [0, 0, 1])
    np.testing.assert_allclose(e["r"], [1.0, 0.0, 0.0], atol=1e-7); np.testing.assert_allclose(e["dir"], [1, 0, 0], atol=1e-7)
    assert (e["uniIdx"], e["localID"], e["cellIdx"]) == (8, 1, 1)
    e = u.enter([2.0, 0.0, 1.0], [0, 1, 0])
    np.testing.assert_allclose(e["r"], [-1.0, 0.0, 2.0], atol=1e-7); np.testing.assert_allclose(e["dir"], [0, -1, 0], atol=1e-7)
    assert (e["uniIdx"], e["localID"], e["cellIdx"]) == (8, 2, 2)
    e = u.enter([0.0, 0.0, 6.5], [1, 0, 0])                       # the undefined region
    np.testing.assert_allclose(e["r"], [4.5, 0.0, 0.0], atol=1e-7); np.testing.assert_allclose(e["dir"], [0, 0, 1], atol=1e-7)
    assert (e["uniIdx"], e["localID"], e["cellIdx"]) == (8, 3, 0)
    d, s = u.distance(1, [-1.0, 0.0, 0.0], [1, 0, 0]); assert d == pytest.approx(3.0, rel=1e-7) and s == 1
    d, s = u.distance(2, [7.0, 0.0, 0.0], [-1, 0, 0]); assert d == pytest.approx(2.0, rel=1e-7) and s == 2
    d, s = u.distance(2, [7.0, 0.0, 0.0], [1, 0, 0]); assert d == INF and s == 0
    assert u.cross(1, [0.0, 2.0, 0.0], [0, 1, 0], 1) == 2
    np.testing.assert_array_equal(u.offset(1), [0, 0, 0]); np.testing.assert_array_equal(u.offset(2), [0, 0, 0])


def test_cell_universe_overlap_check(orc):
    # cellUniverse_test.f90:38-48 (checkOverlap 1), :99 fill, test_overlap :322-373
    u = ol.Uni(orc, CELL2_UNI, {}, 9, env=CELL2_ENV)
    assert u.fill() == [-3, -4, -5, UNDEF_MAT, OVERLAP_MAT]
    e = u.enter([22.0, 0.0, 0.0], [1, 0, 0])
    np.testing.assert_allclose(e["r"], [22.0, 0.0, 0.0], atol=1e-7)
    assert (e["uniIdx"], e["localID"], e["cellIdx"]) == (9, 1, 1)
    e = u.enter([0.0, 25.0, 0.0], [1, 0, 0]); assert (e["localID"], e["cellIdx"]) == (3, 3)
    e = u.enter([24.2, 0.0, 0.0], [1, 0, 0]); assert (e["uniIdx"], e["localID"], e["cellIdx"]) == (9, 5, 0)      # inside cells 1 and 3


# --------------------------------------------------------------------------- rootUniverse_test
def test_root_universe(orc):
    # rootUniverse_test.f90:16-21 (border = surface id 1, the second one defined), :52 fill, test_enter :83-110, test_distance :115-133,
    # test_cross :138-152, test_cellOffset :157-176
    env = "surfaces { surf1 { id 4; type sphere; origin (0.0 5.0 0.0); radius 0.5;} surf2 { id 1; type sphere; origin (0.0 0.0 0.0); radius 2;} }"
    u = ol.Uni(orc, "id 1; type rootUniverse; border 1; fill u<17>;", {}, 8, env=env)
    OUTSIDE_MAT = 0
    assert u.fill() == [-17, OUTSIDE_MAT]
    e = u.enter([1.0, -1.0, 1.0], [1, 0, 0])
    np.testing.assert_allclose(e["r"], [1.0, -1.0, 1.0], atol=1e-7); np.testing.assert_allclose(e["dir"], [1, 0, 0], atol=1e-7)
    assert (e["uniIdx"], e["cellIdx"], e["localID"]) == (8, 0, 1)
    e = u.enter([2.0, -2.0, 1.0], [1, 0, 0])
    assert (e["uniIdx"], e["cellIdx"], e["localID"]) == (8, 0, 2)
    d, s = u.distance(1, [1.0, 0.0, 0.0], [1, 0, 0])
    assert d == pytest.approx(1.0, rel=1e-7) and s == 2                  # index of surface id 1 on the shelf
    assert u.cross(1, [2.0, 0.0, 0.0], [1, 0, 0], 2) == 2
    np.testing.assert_array_equal(u.offset(1), [0, 0, 0]); np.testing.assert_array_equal(u.offset(2), [0, 0, 0])
