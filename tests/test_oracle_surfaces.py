"""CPU: the oracle's surfaces (oracle/geom.hpp) against the known answers of the reference's own surface tests
(Geometry/Surfaces/Tests/{sphere,box,cylinder,aPlane,plane}_test.f90): halfspace with the surface tolerance, ray distance,
explicit and transform boundary conditions."""
import ctypes as C

import numpy as np
import pytest

from tests import oracle_lib as ol

SURF_TOL = 1.0e-12
INF = 9223372036854775808.0          # universalVariables.f90 INF = 2^63
SQRT2, SQRT3 = np.sqrt(2.0), np.sqrt(3.0)


class Surf:
    def __init__(self, orc, text):
        self.orc = orc
        self.h = orc.orc_surf_new(text.encode())
        assert self.h, ol.err(orc)

    def q(self, r, u):
        r = np.ascontiguousarray(r, np.float64); u = np.ascontiguousarray(u, np.float64)
        ev, d, g, hs = C.c_double(), C.c_double(), C.c_int(), C.c_int()
        assert self.orc.orc_surf_query(self.h, ol.dp(r), ol.dp(u), C.byref(ev), C.byref(d), C.byref(g), C.byref(hs)) == 0
        return ev.value, d.value, bool(g.value), bool(hs.value)

    def halfspace(self, r, u):
        return self.q(r, u)[3]

    def distance(self, r, u):
        return self.q(r, u)[1]

    def bc(self, transform, r, u, bcs):
        b = np.ascontiguousarray(bcs, np.int32)
        assert self.orc.orc_surf_set_bc(self.h, ol.ip(b), len(b)) == 0, ol.err(self.orc)
        r = np.array(r, np.float64); u = np.array(u, np.float64)
        assert self.orc.orc_surf_bc(self.h, transform, ol.dp(r), ol.dp(u)) == 0
        return r, u

    def close(self):
        self.orc.orc_surf_free(self.h)


def unit(v):
    v = np.array(v, float)
    return v / np.sqrt((v * v).sum())


def test_sphere(orc):
    # sphere_test.f90:12 SPH_DEF, testHalfspace :95-128, testDistance :131-183, testBC :68-91
    s = Surf(orc, "type sphere; id 7; origin (1.0 2.0 1.0); radius 2.0;")
    r = np.array([-1.0, 2.0, 1.0]); u = np.array([1.0, 0.0, 0.0])
    assert not s.halfspace(r, u)
    assert not s.halfspace(r - 0.5 * SURF_TOL * u, u)
    assert s.halfspace(r - 1.00001 * SURF_TOL * u, u)
    assert s.halfspace(r - u, u) and not s.halfspace(r + u, u)
    assert s.halfspace(r, [0.0, 1.0, 0.0])
    r = np.array([-2.0, 2.0, 1.0])
    assert s.distance(r, u) == pytest.approx(1.0, rel=1e-7)
    ref = 1.5
    ux = (ref ** 2 + 3.0 ** 2 - 2.0 ** 2) / (2.0 * ref * 3.0)
    assert s.distance(r, unit([ux, np.sqrt(1.0 - ux ** 2), 0.0])) == pytest.approx(ref, rel=1e-7)
    r = np.array([-1.0, 2.0, 1.0])
    assert s.distance(r, [0.0, 1.0, 0.0]) == INF
    assert s.distance(r, u) == 4.0
    assert s.distance(r - np.array([1.0, 0, 0]) * 0.5 * SURF_TOL, u) == 4.0 + 0.5 * SURF_TOL
    r = np.array([-0.5, 2.0, 1.0])
    assert s.distance(r, u) == 3.5 and s.distance(r, -u) == 0.5
    for tr in (0, 1):                                   # a sphere has vacuum BC only: nothing changes
        r2, u2 = s.bc(tr, [1.0, 0.0, 1.0], [0.0, 1.0, 0.0], [0])
        assert np.array_equal(r2, [1.0, 0.0, 1.0]) and np.array_equal(u2, [0.0, 1.0, 0.0])
    s.close()


def test_box(orc):
    # box_test.f90:12 BOX_DEF, testBC :72-133, testHalfspace :136-170, testDistance :173-230
    s = Surf(orc, "type box; id 7; origin (1.0 2.0 1.0); halfwidth (1.0 2.0 3.0);")
    bcs = [0, 1, 2, 2, 1, 1]                               # VACUUM, REFLECTIVE, PERIODIC, PERIODIC, REFLECTIVE, REFLECTIVE
    r, u = s.bc(0, [0.0, 1.0, -1.0], [1.0, 0.0, 0.0], bcs)
    np.testing.assert_allclose(r, [0.0, 1.0, -1.0], atol=1e-6); np.testing.assert_allclose(u, [1.0, 0.0, 0.0], atol=1e-6)
    r, u = s.bc(0, [1.0, 1.0, -2.0], [0.0, 0.0, -1.0], bcs)
    np.testing.assert_allclose(r, [1.0, 1.0, -2.0], atol=1e-6); np.testing.assert_allclose(u, [0.0, 0.0, 1.0], atol=1e-6)
    r, u = s.bc(0, [1.0, 0.0, -1.0], [0.0, -1.0, 0.0], bcs)
    np.testing.assert_allclose(r, [1.0, 4.0, -1.0], atol=1e-6); np.testing.assert_allclose(u, [0.0, -1.0, 0.0], atol=1e-6)
    r, u = s.bc(0, [2.0, 0.0, 4.0], unit([1.0, -1.0, 1.0]), bcs)
    np.testing.assert_allclose(r, [2.0, 4.0, 4.0], atol=1e-6); np.testing.assert_allclose(u, unit([-1.0, -1.0, -1.0]), atol=1e-6)
    r, u = s.bc(1, [4.5, -10.0, 4.0], unit([1.0, 1.0, 1.0]), bcs)
    np.testing.assert_allclose(r, [-0.5, 2.0, 4.0], atol=1e-6); np.testing.assert_allclose(u, unit([-1.0, 1.0, -1.0]), atol=1e-6)
    r, u = s.bc(1, [2.0, 0.0, 4.0], unit([1.0, -1.0, 1.0]), bcs)
    np.testing.assert_allclose(r, [2.0, 4.0, 4.0], atol=1e-6); np.testing.assert_allclose(u, unit([-1.0, -1.0, -1.0]), atol=1e-6)
    z = [0.0, 0.0, 1.0]
    assert not s.halfspace([0.5, 1.0, 3.6], z) and not s.halfspace([1.5, 1.0, 0.5], z)
    assert s.halfspace([-0.5, 0.0, 3.6], z) and s.halfspace([2.0, 5.0, 0.5], z)
    r = np.array([1.5, 4.0, 0.5]); u = np.array([0.0, 1.0, 0.0])
    assert s.halfspace(r, u) and s.halfspace(r - 0.5 * SURF_TOL * u, u) and not s.halfspace(r - 1.00001 * SURF_TOL * u, u)
    u2 = [1.0, 0.0, 0.0]
    assert s.halfspace(r + 0.5 * SURF_TOL * u, u2) and not s.halfspace(r - 0.5 * SURF_TOL * u, u2)
    r = np.array([-2.0, 0.001, 1.0]); u = np.array([1.0, 0.0, 0.0])
    assert s.distance(r, u) == pytest.approx(2.0, rel=1e-7) and s.distance(r, -u) == INF
    assert s.distance(r, unit([1.0, -0.002, 1.0])) == INF
    assert s.distance(r, unit([1.0, 0.0, 1.0])) == pytest.approx(2.0 * SQRT2, rel=1e-7)
    assert s.distance([-2.7, -0.3, 1.0 / 3.0], unit([2.7, 0.3, 3.0 + 2.0 / 3.0])) == INF
    assert s.distance([-2.0, 0.0, 1.0], u) == INF
    r = np.array([1.0, 2.0, -2.0])
    assert s.distance(r, z) == pytest.approx(6.0, rel=1e-7) and s.distance(r, [0.0, 0.0, -1.0]) == INF
    assert s.distance([1.0, 2.0, -2.0 - 0.5 * SURF_TOL], z) == pytest.approx(6.0 + 0.5 * SURF_TOL, rel=1e-7)
    r = np.array([1.0, 2.0, -1.0]); u = unit([0.0, 1.0, 1.0])
    assert s.distance(r, u) == pytest.approx(2.0 * SQRT2, rel=1e-7) and s.distance(r, -u) == pytest.approx(SQRT2, rel=1e-7)
    s.close()


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_cylinder(orc, axis):
    # cylinder_test.f90:42-70 (origin (2 2 2), radius 2), testHalfspace, testDistance
    s = Surf(orc, "type %sCylinder; id 75; origin (2.0 2.0 2.0); radius 2.0;" % "xyz"[axis])
    a = axis; p1, p2 = [i for i in range(3) if i != axis]
    r = np.full(3, 2.0); r[p1] = 0.0
    u = np.zeros(3); u[p1] = 1.0; u[a] = 1.0; u = unit(u)
    assert not s.halfspace(r, u) and not s.halfspace(r - SURF_TOL * u, u)
    assert s.halfspace(r - SQRT2 * 1.00001 * SURF_TOL * u, u) and s.halfspace(r - 2.0 * u, u) and not s.halfspace(r + 2.0 * u, u)
    v = np.zeros(3); v[p2] = 1.0; v[a] = 1.0
    assert s.halfspace(r, unit(v))
    r = np.full(3, 2.0); r[p1] = -1.0
    assert s.distance(r, u) == pytest.approx(SQRT2, rel=1e-7)
    w = np.zeros(3); w[a] = 1.0; w[p1] = SQRT3 / 2.0; w[p2] = 0.5
    assert s.distance(r, unit(w)) == pytest.approx(1.275200556 * SQRT2, rel=1e-7)
    w = np.zeros(3); w[a] = 1.0e20; w[p1] = 1.0
    assert s.distance(r, unit(w)) == INF
    r = np.full(3, 2.0); r[p1] = 0.0
    assert s.distance(r, u) == pytest.approx(4.0 * SQRT2, rel=1e-7) and s.distance(r, -u) == INF
    assert s.distance(r, -unit(v)) == INF
    r[p1] = 0.0 - 0.5 * SURF_TOL
    assert s.distance(r, u) == pytest.approx((4.0 + 0.5 * SURF_TOL) * SQRT2, rel=1e-7)
    r = np.full(3, 2.0); r[p1] = 1.0
    assert s.distance(r, u) == pytest.approx(3.0 * SQRT2, rel=1e-7) and s.distance(r, -u) == pytest.approx(SQRT2, rel=1e-7)
    s.close()


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_axis_plane(orc, axis):
    # aPlane_test.f90:42-66 (x0 = 4.3), testHalfspace, testDistance
    s = Surf(orc, "type %sPlane; id 75; %s0 4.3;" % ("xyz"[axis], "xyz"[axis]))
    r = np.zeros(3); u = np.zeros(3); r[axis] = 4.3; u[axis] = 1.0
    assert s.halfspace(r, u) and s.halfspace(r - 0.5 * SURF_TOL * u, u) and not s.halfspace(r - 1.0001 * SURF_TOL * u, u)
    assert s.halfspace(r + 2.0 * u, u) and not s.halfspace(r - 2.0 * u, u)
    u2 = np.roll(u, -1)                                  # cshift(u, 1)
    assert s.halfspace(r + 0.5 * SURF_TOL * u, u2) and not s.halfspace(r - 0.5 * SURF_TOL * u, u2)
    r = np.full(3, 2.0); r[axis] = 0.0; d = unit([1.0, 1.0, 1.0])
    assert s.distance(r, d) == pytest.approx(4.3 * SQRT3, rel=1e-7)
    r[axis] = 4.3
    assert s.distance(r, d) == INF
    r[axis] = 4.3 - 0.5 * SURF_TOL
    assert s.distance(r, d) == INF
    r[axis] = 5.0
    assert s.distance(r, d) == INF and s.distance(r, -d) == pytest.approx(0.7 * SQRT3, rel=1e-7)
    r = np.full(3, 2.0); r[axis] = 0.0
    assert s.distance(r, np.roll(u, -1)) == INF
    s.close()


def test_general_plane(orc):
    # plane_test.f90:12 PLANE_DEF "coeffs (1 1 1 3)", testHalfspace, testDistance
    s = Surf(orc, "type plane; id 7; coeffs (1.0 1.0 1.0 3.0);")
    r = np.array([1.0, 1.0, 1.0]); u = unit([1.0, 1.0, 1.0])
    assert s.halfspace(r, u) and s.halfspace(r - 0.5 * SURF_TOL * SQRT3 * u, u) and not s.halfspace(r - 1.00001 * SURF_TOL * SQRT3 * u, u)
    assert not s.halfspace(r - u, u) and s.halfspace(r + u, u)
    u2 = unit([-1.0, 0.0, 1.0])
    assert s.halfspace(r + 0.5 * SURF_TOL * u, u2) and not s.halfspace(r - 0.5 * SURF_TOL * u, u2)
    x = np.array([1.0, 0.0, 0.0])
    assert s.distance([-1.0, 0.0, 0.0], x) == pytest.approx(4.0, rel=1e-7)
    assert s.distance([3.0, 0.0, 0.0], x) == INF and s.distance([3.0 - SURF_TOL, 0.0, 0.0], x) == INF
    ref = SQRT3 * 1.1 * SURF_TOL
    assert s.distance([3.0 - ref, 0.0, 0.0], x) == pytest.approx(ref, rel=1e-2)
    assert s.distance([4.0, 0.0, 0.0], x) == INF and s.distance([4.0, 0.0, 0.0], -x) == pytest.approx(1.0, rel=1e-7)
    assert s.distance([-1.0, 0.0, 0.0], u2) == INF
    s.close()


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_square_cylinder(orc, axis):
    # squareCylinder_test.f90:72-118 (origin: axis 2, p1 1, p2 2; halfwidths p1 2, p2 3), testBC :186-271, testHalfspace :277-328,
    # testDistance :334-410, testEdgeCases :463-520
    ax = axis
    p1, p2 = [(1, 2), (0, 2), (0, 1)][axis]
    origin = [2.0, 2.0, 2.0]; origin[p1] = 1.0; origin[p2] = 2.0
    hw = [0.0, 0.0, 0.0]; hw[p1] = 2.0; hw[p2] = 3.0
    s = Surf(orc, "type %sSquareCylinder; id 75; origin (%s); halfwidth (%s);" % ("xyz"[axis], " ".join(map(str, origin)), " ".join(map(str, hw))))

    def v(a, b, c):                 # vector given as (axis, p1, p2) components
        out = np.zeros(3); out[ax] = a; out[p1] = b; out[p2] = c
        return out

    bcs = [0] * 6
    bcs[p1 * 2 + 1] = 1; bcs[p1 * 2] = 0; bcs[p2 * 2 + 1] = 2; bcs[p2 * 2] = 2          # BC(p*2) is the +ve face (1-based), BC(p*2-1) the -ve
    r, u = s.bc(0, v(0, -1, 3), v(0, -1, 0), bcs)                                           # vacuum face
    np.testing.assert_allclose(r, v(0, -1, 3), atol=1e-6); np.testing.assert_allclose(u, v(0, -1, 0), atol=1e-6)
    r, u = s.bc(0, v(0, 3, 3), v(0, 1, 0), bcs)                                             # reflection
    np.testing.assert_allclose(r, v(0, 3, 3), atol=1e-6); np.testing.assert_allclose(u, v(0, -1, 0), atol=1e-6)
    r, u = s.bc(0, v(0, 2, 5), v(0, 0, 1), bcs)                                             # periodic
    np.testing.assert_allclose(r, v(0, 2, -1), atol=1e-6); np.testing.assert_allclose(u, v(0, 0, 1), atol=1e-6)
    r, u = s.bc(0, v(0, 3, -1), unit(v(0, 1, -1)), bcs)                                     # corner
    np.testing.assert_allclose(r, v(0, 3, 5), atol=1e-6); np.testing.assert_allclose(u, unit(v(0, -1, -1)), atol=1e-6)
    r, u = s.bc(1, v(0, 20, -13), unit(v(0, 1, -1)), bcs)                                   # transformBC
    np.testing.assert_allclose(r, v(0, -14, 5), atol=1e-6); np.testing.assert_allclose(u, unit(v(0, -1, -1)), atol=1e-6)
    r, u = s.bc(1, v(0, 3, -1), unit(v(0, 1, -1)), bcs)
    np.testing.assert_allclose(r, v(0, 3, 5), atol=1e-6); np.testing.assert_allclose(u, unit(v(0, -1, -1)), atol=1e-6)

    up2 = v(0, 0, 1)
    assert not s.halfspace(v(0, 2, 0), up2) and not s.halfspace(v(0, -0.5, 2), up2)
    assert s.halfspace(v(0, -1.5, 2), up2) and s.halfspace(v(0, 0.5, 5.2), up2)
    r = v(0, -1, 3); u = v(0, -1, 0)
    assert s.halfspace(r, u) and not s.halfspace(r, -u)
    assert s.halfspace(r - 0.5 * SURF_TOL * u, u) and not s.halfspace(r - 1.0001 * SURF_TOL * u, u)
    assert s.halfspace(r + 0.5 * SURF_TOL * u, up2) and not s.halfspace(r - 0.5 * SURF_TOL * u, up2)

    r = v(0, -2, 0)
    u = unit(v(1, 1, 0))
    assert s.distance(r, u) == pytest.approx(SQRT2, rel=1e-7) and s.distance(r, -u) == INF
    assert s.distance(r, unit(v(1, 1, 1))) == pytest.approx(SQRT3, rel=1e-7)
    assert s.distance(r, unit(v(1, 1, -2))) == INF
    assert s.distance(v(0, -1.3, 0.3), unit(v(0, 0.3, -1.3))) == INF                       # corner skim
    assert s.distance(v(0, -1.3, 0.3), v(0, 0, 1)) == INF                                   # parallel
    assert s.distance(v(0, -1, 0.5), v(0, 1, 0)) == pytest.approx(4.0, rel=1e-7)            # at the surface
    assert s.distance(v(0, -1 - 0.5 * SURF_TOL, 0.5), v(0, 1, 0)) == pytest.approx(4.0 + 0.5 * SURF_TOL, rel=1e-7)
    assert s.distance(v(0, -1 + 0.5 * SURF_TOL, 0.5), v(0, 1, 0)) == pytest.approx(4.0 - 0.5 * SURF_TOL, rel=1e-7)
    assert s.distance(v(9, 1.15, 1), v(0, 0, 1)) == pytest.approx(4.0, rel=1e-7)
    assert s.distance(v(9, 1.15, 1), v(0, 0, -1)) == pytest.approx(2.0, rel=1e-7)

    eps = 5.0 * np.finfo(float).eps                        # a particle almost at a corner is outside or escapes with a short move
    for r, u in ((v(0, -1 + eps, -1 + eps), unit(v(0, 2, -1))), (v(0, -1 + eps, -1 + eps), unit(v(0, -1, 2))),
                 (v(0, -1 + eps, -1 + 2 * eps), unit(v(0, 2, -1))), (v(0, -1 + 2 * eps, -1 + eps), unit(v(0, -1, 2)))):
        if not s.halfspace(r, u):
            d = s.distance(r, u)
            assert abs(d) < 1e-6 and s.halfspace(r + d * u, u)
    s.close()


def test_square_cylinder_regression_case(orc):
    # squareCylinder_test.f90:529-546 test_problems: on the +y face moving inwards -> inside
    s = Surf(orc, "type zSquareCylinder; id 7; origin (0.0 0.0 0.0); halfwidth (8.0 1.26 0.0);")
    assert not s.halfspace([-7.63, 1.26, 0.0], [0.0, -1.0, 0.0])
    s.close()


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_trunc_cylinder(orc, axis):
    # truncCylinder_test.f90:66-78 (origin: axis 2, p1 1, p2 2; halfwidth 1.5; radius 2), testBC, testHalfspace, testDistance
    ax = axis
    p1, p2 = [(1, 2), (0, 2), (0, 1)][axis]
    origin = [2.0, 2.0, 2.0]; origin[p1] = 1.0; origin[p2] = 2.0
    s = Surf(orc, "type %sTruncCylinder; id 75; origin (%s); halfwidth 1.5; radius 2.0;" % ("xyz"[axis], " ".join(map(str, origin))))

    def v(a, b, c):
        out = np.zeros(3); out[ax] = a; out[p1] = b; out[p2] = c
        return out

    # BCs: { a_min, a_max }
    r, u = s.bc(0, v(0.5, 0, 2.3), v(-1, 0, 0), [0, 1])                     # vacuum face
    np.testing.assert_allclose(r, v(0.5, 0, 2.3), atol=1e-7); np.testing.assert_allclose(u, v(-1, 0, 0), atol=1e-7)
    r, u = s.bc(0, v(3.5, 0, 2.3), v(1, 0, 0), [0, 1])                      # reflective face
    np.testing.assert_allclose(r, v(3.5, 0, 2.3), atol=1e-7); np.testing.assert_allclose(u, v(-1, 0, 0), atol=1e-7)
    r, u = s.bc(1, v(30.0, 0, 2.3), v(1, 0, 0), [0, 1])                     # transformBC
    np.testing.assert_allclose(r, v(-23.0, 0, 2.3), atol=1e-7); np.testing.assert_allclose(u, v(-1, 0, 0), atol=1e-7)
    r, u = s.bc(0, v(3.5, 0, 2.3), v(1, 0, 0), [2, 2])                      # periodic
    np.testing.assert_allclose(r, v(0.5, 0, 2.3), atol=1e-7); np.testing.assert_allclose(u, v(1, 0, 0), atol=1e-7)
    r, u = s.bc(1, v(12.0, 0, 2.3), v(1, 0, 0), [2, 2])
    np.testing.assert_allclose(r, v(3.0, 0, 2.3), atol=1e-7); np.testing.assert_allclose(u, v(1, 0, 0), atol=1e-7)

    u = v(0, 1, 0)
    assert s.halfspace(v(5.0, 3.0, 4.0), u) and s.halfspace(v(1.3, 1.0, 5.0), u) and s.halfspace(v(1.3, -2.0, 3.3), u)
    assert not s.halfspace(v(3.3, 1.1, 2.2), u) and not s.halfspace(v(2.1, -0.7, 2.1), u)
    u = unit(v(1, 1, 0))
    assert not s.halfspace(v(3.3, -1.0, 2.0), u) and s.halfspace(v(3.3, -1.0, 2.0), -u)               # on the cylinder face
    assert not s.halfspace(v(3.3, -1.0 - 0.5 * SURF_TOL, 2.0), u)
    assert s.halfspace(v(3.3, -1.0 - 1.001 * SURF_TOL, 2.0), u)
    u = v(1, 0, 0)                                                                                     # parallel to the face
    assert s.halfspace(v(3.3, -1.0 - 0.5 * SURF_TOL, 2.0), u) and not s.halfspace(v(3.3, -1.0 + 0.5 * SURF_TOL, 2.0), u)
    u = unit(v(-1, 0, 1))
    assert not s.halfspace(v(3.5, 1.3, 1.8), u) and s.halfspace(v(3.5, 1.3, 1.8), -u)                 # on the axial face
    u = unit(v(-1, 1, 0))
    assert not s.halfspace(v(3.5 + 0.5 * SURF_TOL, 1.3, 1.8), u) and s.halfspace(v(3.5 + 1.001 * SURF_TOL, 1.3, 1.8), u)
    u = unit(v(0, -1, 1))
    assert s.halfspace(v(3.5 + 0.5 * SURF_TOL, 1.3, 1.8), u) and not s.halfspace(v(3.5 - 0.5 * SURF_TOL, 1.3, 1.8), u)

    r = v(5.0, 1.0, 2.0)                                                                               # outside, beyond the axial face
    assert s.distance(r, v(-1, 0, 0)) == pytest.approx(1.5, rel=1e-7)
    assert s.distance(r, unit(v(-1, -1, 0))) == pytest.approx(1.5 * SQRT2, rel=1e-7)
    assert s.distance(r, unit(v(-1, -2, -2))) == INF and s.distance(r, unit(v(0, -2, -2))) == INF and s.distance(r, unit(v(1, -1, -1))) == INF
    r = v(1.5, -2.0, 2.0)                                                                              # outside, beside the cylinder face
    assert s.distance(r, v(0, 1, 0)) == pytest.approx(1.0, rel=1e-7)
    assert s.distance(r, unit(v(1, 1, 0))) == pytest.approx(SQRT2, rel=1e-7)
    uu = v(0, 1.5, np.sqrt(7.0) * 0.5); ref = np.sqrt((uu * uu).sum())
    assert s.distance(r, uu / ref) == pytest.approx(ref, rel=1e-7)
    assert s.distance(r, unit(v(1, 1, 2))) == INF and s.distance(r, v(1, 0, 0)) == INF and s.distance(r, unit(v(1, -1, 0))) == INF
    eps = 0.5 * SURF_TOL                                                                               # in the surface tolerance
    assert s.distance(v(3.5 + eps, 1.3, 1.8), v(-1, 0, 0)) == pytest.approx(3.0, rel=1e-7)
    assert s.distance(v(3.5 - eps, 1.3, 1.8), v(1, 0, 0)) == INF
    assert s.distance(v(3.3, -1.0 - eps, 2.0), v(0, 1, 0)) == pytest.approx(4.0, rel=1e-7)
    assert s.distance(v(3.3, -1.0 + eps, 2.0), v(0, -1, 0)) == INF
    r = v(3.0, 0.0, 2.0); u = unit(v(-1, -1, 0))                                                       # inside
    assert s.distance(r, u) == pytest.approx(SQRT2, rel=1e-7) and s.distance(r, -u) == pytest.approx(0.5 * SQRT2, rel=1e-7)
    assert s.distance(r, v(0, -1, 0)) == pytest.approx(1.0, rel=1e-7) and s.distance(r, v(-1, 0, 0)) == pytest.approx(2.5, rel=1e-7)
    s.close()
