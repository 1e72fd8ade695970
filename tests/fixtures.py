"""Small input fixtures transcribed from the reference's IntegrationTestFiles (input DATA, not code).

TEST_LAT / TEST_CYL: IntegrationTestFiles/Geometry/test_lat, test_cyl
MG_MAT1 / MG_MAT2 : IntegrationTestFiles/mgMat1, mgMat2
"""
TEST_LAT = """
boundary ( 1 1 2 2 0 0);
graph { type shrunk; }
surfaces {
  squareBound { id 1; type zSquareCylinder; origin (0.0 0.0 0.0); halfwidth (1.26 1.26 0.0);}
  sqPinBound { id 2; type zSquareCylinder; origin (0.0 0.0 0.0); halfwidth  (0.5 0.5 0.0);}
}
cells {
  sqPinIn {id 1; type simpleCell; surfaces (-2); filltype mat; material uox; }
  sqPinOut {id 2; type simpleCell; surfaces (2); filltype mat; material water;}
}
universes {
  root { id 1; type rootUniverse; border 1; fill u<10>;}
  pin31 { id 31; type pinUniverse; radii (0.5400 0.0 ); fills (mox43 water);}
  sqPin { id 32; type cellUniverse; cells (1 2);}
  lat10 {id 10; type latUniverse; shape (2 2 0); pitch (1.26 1.26 0.0); padMat water; map (31 32 32 31);}
}
nuclearData { materials {
    water { temp 75675; composition {  } }
    mox43 { temp 87476; composition { } }
    uox { temp 6786; composition { } }
} }
"""

TEST_CYL = """
boundary ( 1 1 2 2 0 0);
graph { type shrunk; }
surfaces {
  squareBound { id 1; type box; origin (0.0 0.0 0.0); halfwidth (5.0 5.0 5.0);}
}
cells { }
universes {
  root { id 1; type rootUniverse; border 1; fill u<10>;}
  pin31 { id 10; type pinUniverse; origin (1.0 0.0 0.0); rotation (0.0 30.0 0.0); radii (0.900 0.0 ); fills (mox43 water);}
  pin32 {id 32; type pinUniverse; radii (0.0); fills (water);}
}
nuclearData { materials {
    water { temp 75675; composition {  } }
    mox43 { temp 87476; composition { } }
    uox { temp 6786; composition { } }
} }
"""

MG_MAT1 = """
numberOfGroups 4;
capture (1.0 2.0 3.0 4.0);
scatteringMultiplicity (
  1.0 1.0 1.0 1.0
  1.0 1.0 1.0 1.0
  1.0 1.0 1.0 1.0
  1.0 1.0 1.0 1.0);
P0 (
  0.5 0.3 0.2 0.1
  0.0 1.0 0.3 0.1
  0.0 0.0 2.0 1.0
  0.0 0.0 0.1 3.0 );
P1 (
 -0.1  0.0 0.0 0.0
  0.0 -0.2 0.0 0.0
  0.0  0.0 0.0 0.0
  0.0  0.0 0.0 0.0 );
"""

MG_MAT2 = MG_MAT1 + """
fission (1.0 0.0 0.0 0.0);
chi (0.8 0.2 0.0 0.0);
nu (2.3 0.0 0.0 0.0);
kappa (202 193 180 200);
"""


def write_mg_deck(tmp_path, pn="P0"):
    """materials mat1 (mgMat1) and mat2 (mgMat2) as in baseMgNeutronDatabase_iTest.f90:22-50."""
    (tmp_path / "mgMat1").write_text(MG_MAT1)
    (tmp_path / "mgMat2").write_text(MG_MAT2)
    deck = tmp_path / "deck"
    deck.write_text("""
nuclearData {
  handles { mg { type baseMgNeutronDatabase; PN %s; } }
  materials {
    mat1 { temp 273; composition { } xsFile ./mgMat1; }
    mat2 { temp 1; composition { } xsFile ./mgMat2; }
  }
}
""" % pn)
    return str(deck)
