"""GPU parity, component level: the device functions the cycle kernel is built from, called through the
C ABI batch queries, against the CPU oracle on the same inputs.  Bar: bit-exact (integers AND doubles)."""
import ctypes as C

import numpy as np
import pytest

import scone_b200
from tests import oracle_lib as ol
from tests.fixtures import TEST_CYL, TEST_LAT
from tests.gpu_util import DECK, random_points

pytestmark = pytest.mark.gpu


def test_rng_skip_and_get(orc):
    L = scone_b200.load_library()
    rng = np.random.default_rng(0)
    n = 20000
    st = rng.integers(0, 2**63 - 1, n, dtype=np.uint64)
    sk = rng.integers(-2**40, 2**40, n, dtype=np.int64)
    sk[:100] = np.arange(100)
    out = np.zeros(n, np.uint64); real = np.zeros(n)
    assert L.sb_rng_query(n, st.ctypes.data_as(C.POINTER(C.c_uint64)), sk.ctypes.data_as(C.POINTER(C.c_int64)),
                          out.ctypes.data_as(C.POINTER(C.c_uint64)), real.ctypes.data_as(C.POINTER(C.c_double))) == 0
    for i in list(range(200)) + list(range(n - 200, n)):
        s = orc.orc_rng_skip(int(st[i]), int(sk[i]))
        s = orc.orc_rng_next(s)
        assert int(out[i]) == s
        assert real[i] == orc.orc_rng_real(s)


def test_deterministic_math_bit_exact(orc):
    L = scone_b200.load_library()
    orc.orc_set_math_mode(1)
    rng = np.random.default_rng(1)
    n = 200000
    x = rng.uniform(0.0, 1.0, n)
    x[:1000] = rng.uniform(0, 2 * np.pi, 1000)
    x[1000:1100] = 2.0 ** -rng.integers(1, 63, 100)
    lg = np.zeros(n); sn = np.zeros(n); cs = np.zeros(n)
    assert L.sb_math_query(n, ol.dp(x), ol.dp(lg), ol.dp(sn), ol.dp(cs)) == 0
    s, c = C.c_double(), C.c_double()
    for i in range(0, n, 37):
        assert lg[i] == orc.orc_math_log(x[i])
        orc.orc_math_sincos(x[i], C.byref(s), C.byref(c))
        assert sn[i] == s.value and cs[i] == c.value
    orc.orc_set_math_mode(0)
    # and within 1 ulp of libm
    np.testing.assert_allclose(lg[x > 0], np.log(x[x > 0]), rtol=3e-16)


def test_fast_div_sqrt_exact():
    """rotateVector's divisions and square roots are nvcc's own expansions of `/` and sqrt with ONE range test for
    the group (sb_device.cuh: rcpRefined / divBy / sqrtFast): every result must be the IEEE one, bit for bit."""
    L = scone_b200.load_library()
    for seed, span in ((1, 3), (2, 60), (3, 500)):
        bad = C.c_int64(-1)
        assert L.sb_fastmath_check(200_000_000, seed, span, C.byref(bad)) == 0
        assert bad.value == 0


@pytest.mark.parametrize("name,src,lo,hi", [
    ("test_lat", TEST_LAT, -1.5, 1.5), ("test_cyl", TEST_CYL, -5.5, 5.5),
    ("c5g7", DECK["c5g7"], -33.0, 33.0), ("c5g7_3d", DECK["c5g7_3d"], -33.0, 70.0), ("can", DECK["can"], -7.5, 7.5)])
def test_place_and_teleport_bit_exact(orc, name, src, lo, hi):
    import os
    text = open(src).read() if os.path.exists(src) else src
    g = scone_b200.GeometryHandle(text, device=0)
    o = ol.Geom(orc, text)
    n = 400000
    r, u = random_points(n, lo, hi, 7)
    # points exactly on lattice faces / pin radii exercise the tolerance branches
    r[:2000, 0] = np.round(r[:2000, 0] / 1.26) * 1.26
    r[2000:4000, 1] = np.round(r[2000:4000, 1] / 0.63) * 0.63
    mat, uid, _, _ = g.geom_query(r, u)
    om = np.zeros(n, np.int32); oq = np.zeros(n, np.int32)
    rr = np.ascontiguousarray(r); uu = np.ascontiguousarray(u)
    assert orc.orc_geom_what_is_at_n(o.h, n, ol.dp(rr), ol.dp(uu), ol.ip(om), ol.ip(oq)) == 0
    np.testing.assert_array_equal(mat, om)
    np.testing.assert_array_equal(uid, oq)
    # teleport with boundary transformations; start inside the domain
    inside = om != 0
    r2 = np.ascontiguousarray(r[inside]); u2 = np.ascontiguousarray(u[inside])
    dist = np.random.default_rng(3).exponential(3.0, len(r2))
    mat, uid, rg, ug = g.geom_query(r2, u2, dist)
    ro = r2.copy(); uo = u2.copy()
    om = np.zeros(len(r2), np.int32); oq = np.zeros(len(r2), np.int32)
    assert orc.orc_geom_teleport_n(o.h, len(r2), ol.dp(ro), ol.dp(uo), ol.dp(dist), ol.ip(om), ol.ip(oq)) == 0
    np.testing.assert_array_equal(mat, om)
    np.testing.assert_array_equal(uid, oq)
    assert np.array_equal(rg, ro) and np.array_equal(ug, uo)      # positions and directions bit-identical
    g.close()


@pytest.mark.parametrize("deck", ["c5g7", "inf", "slab"])
def test_mg_xs_bit_exact(orc, deck):
    pp = scone_b200.EigenPhysicsPackage(DECK[deck], "pop 1000;", device=0)
    db = orc.orc_mg_load(DECK[deck].encode(), b"mg")
    mats = np.repeat(np.arange(1, pp.n_mat + 1), pp.n_groups).astype(np.int32)
    G = np.tile(np.arange(1, pp.n_groups + 1), pp.n_mat).astype(np.int32)
    tot, maj = pp.mg_query(mats, G)
    for i in range(len(mats)):
        assert tot[i] == orc.orc_mg_total(db, int(mats[i]), int(G[i]))
        assert maj[i] == orc.orc_mg_majorant(db, int(G[i]))
    orc.orc_mg_free(db)
    pp.close()
