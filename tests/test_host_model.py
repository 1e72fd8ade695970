"""Host-side flattening (scone_b200/csrc/host/model.hpp) against the oracle's object model: the flat
geometry graph, universe fills, active materials and XS tables the engine is given must be the ones
SCONE's csg/geomGraph/baseMgNeutronDatabase would build. CPU only (no engine is created)."""
import os

import numpy as np
import pytest

import scone_b200
from tests import oracle_lib as ol
from tests.fixtures import TEST_CYL, TEST_LAT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECKS = [os.path.join(ROOT, "decks", "c5g7", "c5g7_2d"), os.path.join(ROOT, "decks", "c5g7", "c5g7_3d_rodded"),
         os.path.join(ROOT, "decks", "urr", "inf"), os.path.join(ROOT, "decks", "urr", "slab"), os.path.join(ROOT, "decks", "mg", "can")]


@pytest.mark.parametrize("src", [TEST_LAT, TEST_CYL] + DECKS)
def test_flat_geometry_equals_oracle(orc, src):
    is_path = os.path.exists(src)
    text = open(src).read() if is_path else src
    g = scone_b200.GeometryHandle(text, device=-1)
    o = ol.Geom(orc, text)
    gi, oi = g.info(), o.info()
    assert gi == oi
    idx, gid = g.graph()
    oidx, ogid = o.graph()
    np.testing.assert_array_equal(idx, oidx)
    np.testing.assert_array_equal(gid, ogid)
    for u in range(1, gi["nUni"] + 1):
        out = np.zeros(1 << 16, np.int32)
        n = orc.orc_geom_uni_fill(o.h, u, ol.ip(out), len(out))
        assert g.uni_fill(u) == out[:n].tolist()
    out = np.zeros(4096, np.int32)
    n = orc.orc_geom_active_mats(o.h, ol.ip(out), 4096)
    assert g.active_mats() == out[:n].tolist()


def test_c5g7_graph_size():
    g = scone_b200.GeometryHandle(open(DECKS[0]).read(), device=-1)
    info = g.info()
    # root 2 + core lattice 10 + 4 assemblies x (290 + 289 x 2) + 5 reflector pins x 1   
    assert info["nGraph"] == 2 + 10 + 4 * (290 + 289 * 2) + 5
    assert info["nesting"] == 4 and info["nUni"] == 11


@pytest.mark.parametrize("deck", DECKS)
def test_flat_xs_equals_oracle(orc, deck):
    pp = scone_b200.EigenPhysicsPackage(deck, device=-1)
    data, maj = pp.model_xs()
    e = orc.orc_eigen_load(deck.encode(), b"")
    assert e, ol.err(orc)
    # same arithmetic, same order => bit-identical tables
    import ctypes as C
    # the oracle database of the eigen handle is reached through a fresh MG load with all materials active;
    # compare row by row on the materials, and the majorant through the eigen driver's own active set
    db = orc.orc_mg_load(deck.encode(), b"mg")
    x = np.zeros(8)
    for m in range(pp.n_mat):
        for g in range(pp.n_groups):
            orc.orc_mg_macro(db, m + 1, g + 1, ol.dp(x))
            row = data[m, g]
            assert row[0] == x[0] and row[1] == x[2] and row[2] == x[3] and row[3] == x[4] and row[4] == x[5] and row[5] == x[6]
    orc.orc_mg_free(db)
    orc.orc_eigen_free(e)
    assert (maj > 0).all()


# --------------------------------------------------------------------------- geomGraph_test / uniFills_test
# A geometry whose universe fill vectors are the ones Geometry/Tests/geomGraph_test.f90:16-46 builds by hand:
#   idx 1 (id 3)    [-7, OUTSIDE]         idx 2 (id 7)    [-1001, -1001, -1003]     idx 3 (id 1001) [1, 4]
#   idx 4 (id 1002) [1, 3]                idx 5 (id 1003) [1, 2]                     idx 6, 7 (ids 200, 201) [-1001, -200] (unused)
GRAPH_GEOM = """
boundary (0 0 0 0 0 0);
graph { type %s; }
surfaces { bound { id 1; type sphere; origin (0.0 0.0 0.0); radius 10.0; } }
cells { }
universes {
  root { id 3; type rootUniverse; border 1; fill u<7>; }
  u7    { id 7;    type pinUniverse; radii (1.0 2.0 0.0); fills (u<1001> u<1001> u<1003>); }
  u1001 { id 1001; type pinUniverse; radii (0.5 0.0); fills (m1 m4); }
  u1002 { id 1002; type pinUniverse; radii (0.5 0.0); fills (m1 m3); }
  u1003 { id 1003; type pinUniverse; radii (0.5 0.0); fills (m1 m2); }
  u200  { id 200;  type pinUniverse; radii (0.5 0.0); fills (u<1001> u<200>); }
  u201  { id 201;  type pinUniverse; radii (0.5 0.0); fills (u<1001> u<200>); }
}
nuclearData { materials { m1 { temp 1; composition { } } m2 { temp 1; composition { } } m3 { temp 1; composition { } } m4 { temp 1; composition { } } } }
"""


@pytest.mark.parametrize("kind,idx_ref,id_ref,unique", [
    ("shrunk", [-2, 0, -3, -3, -5, 1, 4, 1, 2], [3, 0, 6, 6, 8, 1, 2, 3, 4], 4),                       # geomGraph_test.f90 test_shrunk :60-66
    ("extended", [-2, 0, -3, -3, -5, 1, 4, 1, 4, 1, 2], [3, 0, 6, 8, 10, 1, 2, 3, 4, 5, 6], 6)])      # test_extended :96-102
def test_geom_graph_known_answers(orc, kind, idx_ref, id_ref, unique):
    text = GRAPH_GEOM % kind
    for g in (scone_b200.GeometryHandle(text, device=-1), ol.Geom(orc, text)):      # the product's host builder and the oracle's
        idx, gid = g.graph()
        assert idx.tolist() == idx_ref and gid.tolist() == id_ref
        info = g.info()
        assert info["uniqueCells"] == unique and info["nUni"] == 7 and info["rootIdx"] == 1
        assert info["nesting"] == 3                                                  # uniFills_test.f90 test_nesting_count :118
    g = scone_b200.GeometryHandle(text, device=-1)
    assert g.uni_fill(1) == [-2, 0] and g.uni_fill(2) == [-3, -3, -5] and g.uni_fill(3) == [1, 4] and g.uni_fill(6) == [-3, -6]
    assert g.active_mats() == [1, 2, 4]                                              # graph % usedMats, :65 / :101


def test_geometry_structure_errors(orc):
    # uniFills_test.f90 test_cycles :74-111 (a universe below itself), test_outside_search :125-131 (outside below the root)
    cyc = GRAPH_GEOM.replace("fills (m1 m4)", "fills (m1 u<7>)") % "shrunk"
    out = GRAPH_GEOM.replace("fills (m1 m2)", "fills (m1 outside)") % "shrunk"
    for text, msg in ((cyc, "recursion"), (out, "outside fill")):
        with pytest.raises(scone_b200.EngineError, match=msg):
            scone_b200.GeometryHandle(text, device=-1)
        with pytest.raises(RuntimeError, match=msg):
            ol.Geom(orc, text)


@pytest.mark.parametrize("ov,msg", [
    ("reproducible 0;", "reproducible 0"),
    ("uniformFissionSites { type uniFissSitesField; }", "uniformFissionSites"),
    ("temperature { type cartesianField; }", "temperature"),
    ("activeTally { batchSize 5; fiss { type collisionClerk; response (f); f { type fluxResponse; } } }", "batchSize"),
    ("activeTally { c { type collisionClerk; response (f); f { type fluxResponse; } filter { type energyFilter; Emin 0.0; Emax 1.0; } } }", "filters"),
    ("activeTally { c { type mgXsClerk; } }", "not supported"),
    ("printSource 5;", "printSource must be")])
def test_unsupported_options_are_refused_not_ignored(ov, msg):
    """Options of the reference that would change the results and have no device implementation stop the run with a message."""
    with pytest.raises(scone_b200.EngineError, match=msg):
        scone_b200.EigenPhysicsPackage(DECKS[0], "seed 1; " + ov, device=-1)
