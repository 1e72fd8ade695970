"""Randomly generated geometries (seeded): nested lattices (2-D and 3-D, with and without offset maps, universe and material padding),
pin universes with several radii, rotated / translated cell universes with spheres, cylinders, planes, boxes and truncated cylinders,
every border type.  For each: material and unique-cell ID at random points and after teleports with the boundary transformations
(bit-identical to the oracle), then whole cycles under surface and delta tracking (banks bit-identical)."""
import numpy as np
import pytest

import scone_b200
from tests import oracle_lib as ol
from tests.gpu_util import DECK, random_points
from tests.test_gpu_eigen import oracle_bank

pytestmark = pytest.mark.gpu
MATS = ["UO2", "mox43", "mox7", "mox87", "GT", "FC", "water"]


def gen_geometry(seed):
    rng = np.random.default_rng(seed)

    def euler(lo, hi):
        """ZXZ angles for which the reference's rotationMatrix (genericProcedures.f90:1128-1138) is orthonormal: its element (1,3) is
        sin(psi) sin(phi) where the rotation has sin(psi) sin(theta), so psi = 0 or theta = phi (the oracle and the engine keep the
        element as the reference has it; with other angles directions stop being unit vectors inside the rotated universe)."""
        a, b, c = rng.uniform(lo, hi, 3)
        return (a, b, 0.0) if rng.random() < 0.5 else (a, a, c)

    surfaces, cells, unis = [], [], []
    sid = [10]; cid = [10]; uid = [100]

    def mat():
        return MATS[int(rng.integers(0, len(MATS)))]

    def pin(depth):
        n = int(rng.integers(1, 4))
        radii = np.sort(rng.uniform(0.15, 0.6, size=n))
        u = uid[0]; uid[0] += 1
        fills = " ".join(mat() for _ in range(n + 1))
        rot = "rotation (%.1f %.1f %.1f); " % euler(0, 90) if rng.random() < 0.3 else ""
        org = "origin (%.3f %.3f 0.0); " % tuple(rng.uniform(-0.05, 0.05, 2)) if rng.random() < 0.3 else ""
        unis.append("p%d { id %d; type pinUniverse; %s%sradii (%s 0.0); fills (%s); }" % (u, u, org, rot, " ".join("%.4f" % r for r in radii), fills))
        return u

    def cellu(depth, half):
        """cell universe: a sphere / z-cylinder / truncated cylinder / box inside, the rest outside"""
        u = uid[0]; uid[0] += 1
        kind = int(rng.integers(0, 4))
        s = sid[0]; sid[0] += 1
        if kind == 0:
            surfaces.append("s%d { id %d; type sphere; origin (%.3f %.3f %.3f); radius %.3f; }" % (s, s, *rng.uniform(-0.1, 0.1, 3), half * rng.uniform(0.4, 0.8)))
        elif kind == 1:
            surfaces.append("s%d { id %d; type %sCylinder; origin (%.3f %.3f %.3f); radius %.3f; }" % (s, s, "xyz"[int(rng.integers(0, 3))], *rng.uniform(-0.1, 0.1, 3), half * rng.uniform(0.3, 0.7)))
        elif kind == 2:
            surfaces.append("s%d { id %d; type %sTruncCylinder; origin (%.3f %.3f %.3f); halfwidth %.3f; radius %.3f; }" % (
                s, s, "xyz"[int(rng.integers(0, 3))], *rng.uniform(-0.1, 0.1, 3), half * rng.uniform(0.3, 0.8), half * rng.uniform(0.3, 0.7)))
        else:
            surfaces.append("s%d { id %d; type box; origin (%.3f %.3f %.3f); halfwidth (%.3f %.3f %.3f); }" % (s, s, *rng.uniform(-0.1, 0.1, 3), *(half * rng.uniform(0.3, 0.8, 3))))
        p = sid[0]; sid[0] += 1
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        surfaces.append("s%d { id %d; type plane; coeffs (%.4f %.4f %.4f %.4f); }" % (p, p, *n, rng.uniform(-0.1, 0.1)))
        c1, c2, c3 = cid[0], cid[0] + 1, cid[0] + 2; cid[0] += 3
        inner = ("filltype uni; universe %d;" % pin(depth + 1)) if (depth < 2 and rng.random() < 0.4) else ("filltype mat; material %s;" % mat())
        cells.append("c%d { id %d; type simpleCell; surfaces (-%d); %s }" % (c1, c1, s, inner))
        cells.append("c%d { id %d; type simpleCell; surfaces (%d -%d); filltype mat; material %s; }" % (c2, c2, s, p, mat()))
        cells.append("c%d { id %d; type simpleCell; surfaces (%d %d); filltype mat; material %s; }" % (c3, c3, s, p, mat()))
        rot = "rotation (%.1f %.1f %.1f); " % euler(0, 180) if rng.random() < 0.5 else ""
        org = "origin (%.3f %.3f %.3f); " % tuple(rng.uniform(-0.1, 0.1, 3)) if rng.random() < 0.5 else ""
        unis.append("u%d { id %d; type cellUniverse; %s%scells (%d %d %d); }" % (u, u, org, rot, c1, c2, c3))
        return u

    def lattice(depth, half):
        """lattice filling a cube of half-width `half` (or a column if 2-D)"""
        u = uid[0]; uid[0] += 1
        three_d = rng.random() < 0.4
        nx, ny = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        nz = int(rng.integers(1, 3)) if three_d else 0
        px, py = 2 * half / nx, 2 * half / ny
        pz = 2 * half / nz if three_d else 0.0
        sub_half = 0.5 * min(px, py, pz if three_d else px)
        n = nx * ny * max(1, nz)
        kids = []
        for _ in range(int(rng.integers(1, 4))):
            r = rng.random()
            kids.append(lattice(depth + 1, sub_half) if (depth < 1 and r < 0.25 and sub_half > 0.5) else (cellu(depth + 1, sub_half) if r < 0.6 else pin(depth + 1)))
        m = " ".join(str(kids[int(rng.integers(0, len(kids)))]) for _ in range(n))
        off = ("offsetMap (%s); " % " ".join(str(int(rng.integers(0, 2))) for _ in range(n))) if rng.random() < 0.4 else ""
        pad = mat() if rng.random() < 0.7 else "u<%d>" % pin(depth + 1)
        unis.append("l%d { id %d; type latUniverse; origin (0.0 0.0 0.0); shape (%d %d %d); pitch (%.6f %.6f %.6f); padMat %s; %smap (%s); }" % (
            u, u, nx, ny, nz, px, py, pz, pad, off, m))
        return u

    half = float(rng.uniform(1.5, 3.0))
    top = lattice(0, half) if rng.random() < 0.7 else cellu(0, half)
    b = int(rng.integers(0, 4))
    pick = lambda: int(rng.integers(0, 3))          # noqa: E731
    if b == 0:
        bc = [pick() for _ in range(6)]
        for a in range(3):
            if (bc[2 * a] == 2) != (bc[2 * a + 1] == 2):
                bc[2 * a] = bc[2 * a + 1] = 2
        border = "bound { id 1; type box; origin (0.0 0.0 0.0); halfwidth (%.4f %.4f %.4f); }" % (half, half, half)
    elif b == 1:
        bc = [1, 1, 1, 1, 0, 0]
        if rng.random() < 0.5:
            bc[0] = bc[1] = 2
        border = "bound { id 1; type zSquareCylinder; origin (0.0 0.0 0.0); halfwidth (%.4f %.4f 0.0); }" % (half, half)
        # infinite in z needs something that ends histories: keep z finite through the materials (absorbers) - fine for few cycles
    elif b == 2:
        bc = [int(rng.integers(0, 2)), int(rng.integers(0, 2)), 0, 0, 0, 0]
        border = "bound { id 1; type zTruncCylinder; origin (0.0 0.0 0.0); halfwidth %.4f; radius %.4f; }" % (half, half)
    else:
        bc = [0, 0, 0, 0, 0, 0]
        border = "bound { id 1; type sphere; origin (0.0 0.0 0.0); radius %.4f; }" % half
    text = "geometry { type geometryStd; boundary (%s); graph { type %s; } surfaces { %s %s } cells { %s } universes { root { id 1; type rootUniverse; border 1; fill u<%d>; } %s } }" % (
        " ".join(map(str, bc)), "extended" if rng.random() < 0.3 else "shrunk", border, " ".join(surfaces), " ".join(cells), top, " ".join(unis))
    return text, half


@pytest.mark.parametrize("seed", list(range(32)))
def test_random_geometry(orc, seed):
    geom, half = gen_geometry(seed)
    deck_text = open(DECK["c5g7"]).read()
    nd = deck_text[deck_text.index("nuclearData"):]
    nd = nd.replace("./xs/", DECK["c5g7"].rsplit("/", 1)[0] + "/xs/")
    gtext = geom[geom.index("{") + 1: geom.rindex("}")] + " " + nd                      # geometry-level dictionary with its materials
    g = scone_b200.GeometryHandle(gtext, device=0)
    o = ol.Geom(orc, gtext)
    assert g.info() == o.info()
    n = 60000
    r, u = random_points(n, -1.1 * half, 1.1 * half, seed + 1000)
    mat, uid, _, _ = g.geom_query(r, u)
    om = np.zeros(n, np.int32); oq = np.zeros(n, np.int32)
    rr = np.ascontiguousarray(r); uu = np.ascontiguousarray(u)
    assert orc.orc_geom_what_is_at_n(o.h, n, ol.dp(rr), ol.dp(uu), ol.ip(om), ol.ip(oq)) == 0
    np.testing.assert_array_equal(mat, om); np.testing.assert_array_equal(uid, oq)
    inside = om != 0
    r2 = np.ascontiguousarray(r[inside]); u2 = np.ascontiguousarray(u[inside])
    dist = np.random.default_rng(seed + 7).exponential(1.5 * half, len(r2))
    mat, uid, rg, ug = g.geom_query(r2, u2, dist)
    ro = r2.copy(); uo = u2.copy()
    om = np.zeros(len(r2), np.int32); oq = np.zeros(len(r2), np.int32)
    assert orc.orc_geom_teleport_n(o.h, len(r2), ol.dp(ro), ol.dp(uo), ol.dp(dist), ol.ip(om), ol.ip(oq)) == 0
    np.testing.assert_array_equal(mat, om); np.testing.assert_array_equal(uid, oq)
    assert np.array_equal(rg, ro) and np.array_equal(ug, uo)
    g.close()
    # whole cycles
    for tracking in ("transportOperatorST", "transportOperatorDT"):
        ov = "pop 1500; inactive 1; active 1; seed %d; inactiveTally { } activeTally { } transportOperator { type %s; } %s" % (seed + 3, tracking, geom)
        orc.orc_set_math_mode(1)
        try:
            e = orc.orc_eigen_load(DECK["c5g7"].encode(), ov.encode())
            assert e, ol.err(orc)
            if orc.orc_eigen_init_source(e) != 0:
                assert "fissile" in ol.err(orc)                      # a geometry without fissile material: both sides must refuse
                with pytest.raises(scone_b200.EngineError, match="fissile"):
                    pp = scone_b200.EigenPhysicsPackage(DECK["c5g7"], ov, device=0); pp.generateInitialState()
                orc.orc_eigen_free(e)
                continue
            pp = scone_b200.EigenPhysicsPackage(DECK["c5g7"], ov, device=0)
            pp.generateInitialState()
            for a, b in zip(pp.bank(), oracle_bank(orc, e)):
                assert np.array_equal(a, b)
            k_o = orc.orc_eigen_keff0(e)
            for cyc in range(2):
                try:
                    pp.cycle(cyc >= 1)
                    gpu_err = None
                except scone_b200.EngineError as ex:
                    gpu_err = str(ex)
                k_o = orc.orc_eigen_cycle(e, 1 if cyc >= 1 else 0, k_o)
                if np.isnan(k_o) or gpu_err:                         # e.g. the fission bank died out: both sides must stop
                    assert np.isnan(k_o) and gpu_err, "only one side failed: oracle %r, device %r" % (ol.err(orc), gpu_err)
                    break
                for a, b in zip(pp.bank(), oracle_bank(orc, e)):
                    assert np.array_equal(a, b), "bank differs after cycle %d (%s)" % (cyc, tracking)
            pp.close(); orc.orc_eigen_free(e)
        finally:
            orc.orc_set_math_mode(0)
