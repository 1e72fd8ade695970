"""CPU: the oracle's scoreMemory and k-eff clerks against the known answers of the reference's own tests
(Tallies/Tests/scoreMemory_test.f90, Tallies/TallyClerks/Tests/keffImplicitClerk_test.f90, keffAnalogClerk_test.f90)."""
import ctypes as C

import numpy as np
import pytest

from tests import oracle_lib as ol


def _scores():
    # scoreMemory_test.f90:36-52: LC PRNG A = 2469, M = 65521, seed 9294; scores = 2 + sin(pi*random - pi/2)
    seed, rnd = 9294, []
    for _ in range(200):
        seed = (2469 * seed) % 65521
        rnd.append(seed / 65521.0)
    rnd = np.array(rnd)
    return 2.0 + np.sin(np.pi * rnd - np.pi / 2), (rnd * 100).astype(np.int32)


def _result(orc, m, idx, samples=-1):
    mean, std = C.c_double(), C.c_double()
    orc.orc_mem_result(m, idx, samples, C.byref(mean), C.byref(std))
    return mean.value, std.value


def test_score_memory_known_answers(orc):
    # scoreMemory_test.f90:55-130 testScoring (TOL 1e-9)
    scores, ints = _scores()
    m = orc.orc_mem_new(7, 1)
    for i in range(10):
        for j in range(20 * i, 20 * (i + 1)):
            orc.orc_mem_score(m, float(scores[j]), 1); orc.orc_mem_score(m, float(ints[j]), 2); orc.orc_mem_score(m, float(ints[j]), 3)
            orc.orc_mem_accumulate(m, float(scores[j]), 4); orc.orc_mem_accumulate(m, float(ints[j]), 5); orc.orc_mem_accumulate(m, float(ints[j]), 6)
        orc.orc_mem_reduce(m)
        assert orc.orc_mem_close_bin(m, 1.2, 3) == 0
        orc.orc_mem_close_cycle(m, 0.7)
    for idx, samples, mean, std in ((1, -1, 26.401471259728442, 0.645969443981583), (2, -1, 623.0, 27.982494527829360), (3, -1, 1068.0, 47.969990619136050),
                                    (4, 200, 1.885819375694888, 0.049102082638055), (5, 200, 44.5, 2.015580019267494), (6, 200, 44.5, 2.015580019267494),
                                    (7, -1, 0.0, 0.0), (-7, -1, 0.0, 0.0), (8, -1, 0.0, 0.0)):
        r = _result(orc, m, idx, samples)
        assert r[0] == pytest.approx(mean, abs=1e-9) and r[1] == pytest.approx(std, abs=1e-9), idx
    orc.orc_mem_free(m)


def test_score_memory_batches_and_get_score(orc):
    # testLastCycle :132-146 (batchSize 8), testGetScore :148-160
    m = orc.orc_mem_new(1, 8)
    for i in range(1, 17):
        assert bool(orc.orc_mem_last_cycle(m)) == (i in (8, 16))
        orc.orc_mem_close_cycle(m, 1.0)
    orc.orc_mem_free(m)
    m = orc.orc_mem_new(1, 1)
    for _ in range(3):
        orc.orc_mem_score(m, 1.0, 1)
    orc.orc_mem_reduce(m)
    assert orc.orc_mem_get_score(m, 1) == 3.0 and orc.orc_mem_get_score(m, 0) == 0.0 and orc.orc_mem_get_score(m, 2) == 0.0
    orc.orc_mem_free(m)


def test_keff_implicit_clerk_known_answer(orc):
    # keffImplicitClerk_test.f90:33-34 (total 1, capture 2, fission 1, nuFission 3), test1CycleBatch :45-88
    a = lambda v: np.ascontiguousarray(v, np.float64)
    wc, wp, wl = a([0.7, 0.6]), a([0.1, 0.1]), a([0.3, 0.3])
    k, s = C.c_double(), C.c_double()
    assert orc.orc_keff_implicit_sequence(1.0, 2.0, 1.0, 3.0, 2, ol.dp(wc), ol.dp(wp), ol.dp(wl), C.byref(k), C.byref(s)) == 0, ol.err(orc)
    assert k.value == pytest.approx(0.906521739130435, abs=1e-9) and s.value == pytest.approx(0.006521739130435, abs=1e-9)


def test_keff_analog_clerk_known_answer(orc):
    # keffAnalogClerk_test.f90 test1CycleBatch: (1000 -> 1200, k 1.0), (1000 -> 900, k 1.2): k = 1.14 +- 0.06
    a = lambda v: np.ascontiguousarray(v, np.float64)
    ws, we, kn = a([1000.0, 1000.0]), a([1200.0, 900.0]), a([1.0, 1.2])
    k, s = C.c_double(), C.c_double()
    assert orc.orc_keff_analog_sequence(2, ol.dp(ws), ol.dp(we), ol.dp(kn), 0.8, C.byref(k), C.byref(s)) == 0, ol.err(orc)
    assert k.value == pytest.approx(1.14, abs=1e-9) and s.value == pytest.approx(0.06, abs=1e-9)


# --------------------------------------------------------------------------- tally maps
class Map:
    def __init__(self, orc, text, mats=None):
        self.orc = orc
        ml = " ".join("%s %d" % kv for kv in (mats or {}).items())
        self.h = orc.orc_map_new(text.encode(), ml.encode())
        assert self.h, ol.err(orc)

    def bins(self):
        return self.orc.orc_map_bins(self.h)

    def map(self, r=(0.0, 0.0, 0.0), E=1.0, mg=False, G=0, mat=0):
        r = np.ascontiguousarray(r, np.float64)
        return self.orc.orc_map_map(self.h, ol.dp(r), float(E), 1 if mg else 0, G, mat)


def test_energy_map_known_answers(orc):
    # Tallies/TallyMaps/Maps1D/Tests/energyMap_test.f90: testLinearGrid, testLogGrid, testUnstructGrid, testMGParticle, testBinNumber
    lin = Map(orc, "type energyMap; grid lin; min 0.01; max 10.0; N 20;")
    log = Map(orc, "type energyMap; grid log; min 1.0E-7; max 10.0; N 20;")
    grid = [0.00000000001, 0.00000003, 0.000000058, 0.00000014, 0.00000028, 0.00000035, 0.000000625, 0.000000972, 0.00000102,
            0.000001097, 0.00000115, 0.000001855, 0.000004, 0.000009877, 0.000015968, 0.000148728, 0.00553, 0.009118, 0.111, 0.5,
            0.821, 1.353, 2.231, 3.679, 6.0655, 10.0]
    uns = Map(orc, "type energyMap; grid unstruct; bins (%s);" % " ".join("%.9E" % x for x in grid))
    assert [lin.map(E=e) for e in (7.5774, 9.3652, 3.9223, 6.5548, 1.7119, 20.0)] == [16, 19, 8, 14, 4, 0]
    assert [log.map(E=e) for e in (0.0445008907555061, 1.79747463687278e-07, 1.64204055725811e-05, 2.34083673923110e-07,
                                   5.98486350302033e-07, 20.0)] == [15, 1, 6, 1, 2, 0]
    assert [uns.map(E=e) for e in (0.0761191517392624, 0.00217742635754091, 6.38548311340975e-08, 2.52734532533842,
                                   2.59031729968032e-11, 20.0)] == [18, 16, 3, 23, 1, 0]
    pre = Map(orc, "type energyMap; grid predef; name casmo23;")           # testPredefGrid
    assert [pre.map(E=e) for e in (0.0445008907555061, 1.79747463687278e-07, 1.64204055725811e-05, 2.34083673923110e-07,
                                   5.98486350302033e-07, 20.0)] == [16, 4, 13, 4, 6, 0]
    assert lin.map(mg=True, G=1) == 0 and log.map(mg=True, G=1) == 0 and uns.map(mg=True, G=1) == 0 and pre.map(mg=True, G=1) == 0
    assert (lin.bins(), log.bins(), uns.bins(), pre.bins()) == (20, 20, 25, 23)
    for name, n in (("wims69", 69), ("wims172", 172), ("casmo40", 40), ("casmo12", 12), ("casmo7", 7), ("ecco33", 33), ("vitaminj", 175)):
        assert Map(orc, "type energyMap; grid predef; name %s;" % name).bins() == n


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_space_map_known_answers(orc, axis):
    # spaceMap_test.f90: testStructuredGrid, testUnstructuredGrid, testBins
    s = Map(orc, "type spaceMap; axis %s; grid lin; min -10.0; max 10.0; N 20;" % "xyz"[axis])
    u = Map(orc, "type spaceMap; axis %s; grid unstruct; bins (-10.0 -8.0 -6.0 -4.0 -2.0 0.0 2.0 4.0 6.0 8.0 10.0);" % "xyz"[axis])

    def at(x):
        r = [0.0, 0.0, 0.0]; r[axis] = x
        return r
    assert [s.map(r=at(x)) for x in (0.5, -10.1)] == [11, 0]
    assert [u.map(r=at(x)) for x in (0.5, -10.1)] == [6, 0]
    assert (s.bins(), u.bins()) == (20, 10)


def test_material_map_known_answers(orc):
    # materialMap_test.f90: materials mat1..mat5, map over (mat2 mat3 mat5), with and without the undefined bin
    mats = {"mat%d" % i: i for i in range(1, 6)}
    a = Map(orc, "type materialMap; materials (mat2 mat3 mat5);", mats)
    b = Map(orc, "type materialMap; materials (mat2 mat3 mat5); undefBin true;", mats)
    assert [a.map(mat=i) for i in range(1, 6)] == [0, 1, 2, 0, 3]
    assert [b.map(mat=i) for i in range(1, 6)] == [4, 1, 2, 4, 3]
    assert (a.bins(), b.bins()) == (3, 4)


def test_multi_map_known_answers(orc):
    # Tallies/TallyMaps/Tests/multiMap_test.f90: testBinAndDimension, testMapping
    m = Map(orc, """type multiMap; maps (map1 map2 map3);
                    map1 {type spaceMap; axis x; grid unstruct; bins (0.0 1.0 2.0); }
                    map2 {type spaceMap; axis y; grid unstruct; bins (0.0 2.0 4.0 6.0); }
                    map3 {type spaceMap; axis z; grid unstruct; bins (0.0 3.0 6.0); }""")
    assert m.bins() == 12
    assert m.map(r=[0.1, 0.1, 0.1]) == 1 and m.map(r=[1.1, 5.1, 5.1]) == 12 and m.map(r=[1.1, 3.1, 2.1]) == 4
    assert m.map(r=[-1.1, 5.1, 5.1]) == 0 and m.map(r=[1.1, 50.1, 5.1]) == 0 and m.map(r=[1.1, 5.1, -5.1]) == 0


# --------------------------------------------------------------------------- collisionClerk / trackClerk
MATS7 = {"m%d" % i: i for i in range(1, 8)}
MAP7 = "map { type materialMap; materials (m1 m2 m3 m4 m5 m6 m7); }"          # stands in for testMap (bin = matIdx, maxIdx 7)


def clerk_sequence(orc, text, kind, events, xs=0.3, tracking=0.3):
    mat = np.array([e[0] for e in events], np.int32); w = np.array([e[1] for e in events], float); aux = np.array([e[2] for e in events], float)
    out = np.zeros(64)
    ml = " ".join("%s %d" % kv for kv in MATS7.items())
    n = orc.orc_clerk_sequence(text.encode(), ml.encode(), xs, tracking, kind, len(events), ol.ip(mat), ol.dp(w), ol.dp(aux), ol.dp(out), 64)
    assert n >= 0, ol.err(orc)
    return out[:n]


@pytest.mark.parametrize("virtual_handling", [False, True])
def test_collision_clerk_known_answers(orc, virtual_handling):
    # collisionClerk_test.f90 testScoring (:92-157, handleVirtual 0: the virtual collision of weight 1000.3 is ignored) and
    # testScoringVirtual (:160-245, default handling: a virtual collision scores w / tracking XS): cases 1 and 3 (no filter);
    # constant cross sections 0.3, weights 0.7 (material 1) and 1.3 (material 6)
    s1, s2 = 0.7 / 0.3, 1.3 / 0.3
    if virtual_handling:
        head, ev = "type collisionClerk;", [(1, 0.7, 1), (6, 1.3, 0)]
    else:
        head, ev = "type collisionClerk; handleVirtual 0;", [(1, 0.7, 0), (6, 1.3, 0), (6, 1000.3, 1)]
    flux = " response (flux); flux { type fluxResponse; }"
    r = clerk_sequence(orc, head + flux, 0, ev)
    assert len(r) == 1 and r[0] == pytest.approx(s1 + s2, abs=1e-9)                       # case 1: single bin
    r = clerk_sequence(orc, head + flux + MAP7, 0, ev)
    ref = np.zeros(7); ref[0] = s1; ref[5] = s2
    np.testing.assert_allclose(r, ref, atol=1e-9)                                          # case 3: map, no filter
    # cases 5 / 7 with the second response: the reference's testResponse (constant 1.3) is replaced by the total macroscopic
    # cross section of the constant database (0.3); bins are response-fastest as in the reference (results(1), (2), (11), (12))
    two = " response (flux tot); flux { type fluxResponse; } tot { type macroResponse; MT -1; }"
    r = clerk_sequence(orc, head + two + MAP7, 0, ev)
    ref = np.zeros(14); ref[0] = s1; ref[1] = s1 * 0.3; ref[10] = s2; ref[11] = s2 * 0.3
    np.testing.assert_allclose(r, ref, atol=1e-9)


def test_track_clerk_known_answers(orc):
    # trackClerk_test.f90 testScoring: path length 0.3 at weights 0.7 (material 1) and 1.3 (material 6): score = w * L, cases 1 and 3
    s1, s2 = 0.7 * 0.3, 1.3 * 0.3
    ev = [(1, 0.7, 0.3), (6, 1.3, 0.3)]
    flux = "type trackClerk; response (flux); flux { type fluxResponse; }"
    r = clerk_sequence(orc, flux, 1, ev)
    assert len(r) == 1 and r[0] == pytest.approx(s1 + s2, abs=1e-9)
    r = clerk_sequence(orc, flux + MAP7, 1, ev)
    ref = np.zeros(7); ref[0] = s1; ref[5] = s2
    np.testing.assert_allclose(r, ref, atol=1e-9)


# --------------------------------------------------------------------------- grid_class
def grid(orc, kind, mini=0.0, maxi=0.0, N=0, bins=(), keys=()):
    b = np.ascontiguousarray(bins, float) if len(bins) else np.zeros(1)
    k = np.ascontiguousarray(keys, float) if len(keys) else np.zeros(1)
    idx = np.zeros(max(1, len(keys)), np.int32); bounds = np.zeros(256)
    n = orc.orc_grid(kind, mini, maxi, N, ol.dp(b), len(bins), len(keys), ol.dp(k), ol.ip(idx), ol.dp(bounds), 256)
    assert n > 0, ol.err(orc)
    return idx[:len(keys)].tolist(), bounds[:n]


def test_grid_known_answers(orc):
    # SharedModules/Tests/grid_test.f90: lin(-10.71, 10.71, 17), log(1e-11, 20, 70), the 9-point unstructured grid; valueOutsideArray = -1
    FP = 50 * np.finfo(float).eps
    OUT = -1
    idx, b = grid(orc, 0, -10.71, 10.71, 17, keys=[-10.71, 3.13245, -8.96, -20.0, 10.72])
    assert idx == [1, 11, 2, OUT, OUT] and len(b) == 18
    for i, v in ((1, -10.71), (4, -6.93), (7, -3.15), (14, 5.67), (18, 10.71)):
        assert b[i - 1] == pytest.approx(v, rel=FP)
    idx, b = grid(orc, 1, 1.0e-11, 20.0, 70, keys=[1.0e-11, 6.7e-4, 1.0, 9.0e-12, 21.0])
    assert idx == [1, 45, 63, OUT, OUT] and len(b) == 71
    for i, v in ((1, 1.0e-11), (35, 9.435957972757912e-6), (70, 13.344459739056777), (71, 20.0)):
        assert b[i - 1] == pytest.approx(v, rel=FP)
    idx, b = grid(orc, 2, bins=[1.00e-11, 5.80e-08, 1.40e-07, 2.80e-07, 6.25e-07, 4.00e-06, 0.005530, 0.821000, 10.0],
                  keys=[1.0e-11, 6.7e-4, 1.0, 9.0e-12, 21.0])
    assert idx == [1, 6, 8, OUT, OUT] and len(b) == 9


def test_shannon_entropy_clerk_known_answers(orc):
    # shannonEntropyClerk_test.f90 testSimpleUseCase: 2 bins, 2 cycles; sites of equal weight in both bins -> 1 bit, both in one bin -> 0;
    # memory = N + 1 + cycles bins (testMap is stood in by a materialMap over two materials)
    text = "type shannonEntropyClerk; cycles 2; map { type materialMap; materials (m1 m2); }"
    counts = np.array([2, 2], np.int32); mat = np.array([2, 1, 1, 1], np.int32); w = np.ones(4); x = np.zeros(4); out = np.zeros(2)
    n = orc.orc_shannon_sequence(text.encode(), b"m1 1 m2 2", 2, ol.ip(counts), ol.ip(mat), ol.dp(x), ol.dp(w), ol.dp(out))
    assert n == 2 + 1 + 2, ol.err(orc)
    assert out[0] == pytest.approx(1.0, abs=1e-7) and out[1] == pytest.approx(0.0, abs=1e-7)
    # unequal weights over a space map, and a third cycle beyond `cycles` that must not be scored
    text = "type shannonEntropyClerk; cycles 2; map { type spaceMap; axis x; grid lin; min 0.0; max 4.0; N 4; }"
    counts = np.array([3, 4, 2], np.int32); x = np.array([0.5, 1.5, 1.6, 0.1, 1.1, 2.1, 3.1, 0.5, 9.0]); w = np.array([1.0, 2.0, 1.0, 1, 1, 1, 1, 1, 1.0])
    mat = np.zeros(9, np.int32); out = np.zeros(3)
    assert orc.orc_shannon_sequence(text.encode(), b"", 3, ol.ip(counts), ol.ip(mat), ol.dp(x), ol.dp(w), ol.dp(out)) == 4 + 1 + 2
    p = np.array([0.25, 0.75])
    assert out[0] == pytest.approx(-(p * np.log2(p)).sum(), abs=1e-12) and out[1] == pytest.approx(2.0, abs=1e-12)


def test_response_known_answers(orc):
    # TallyResponses/Tests/macroResponse_test.f90:29,62-68 (database: total 6, elastic 3, inelastic 0, capture 2, fission 1, nuFission 1.5,
    # kappa 9): total 6, disappearance (capture) 2, fission 1, nuFission 1.5, absorbtion 3, kappa-fission 9; fluxResponse_test.f90: 1
    xs = np.array([6.0, 3.0, 0.0, 2.0, 1.0, 1.5, 9.0])
    for mt, ref in ((-1, 6.0), (-2, 2.0), (-6, 1.0), (-7, 1.5), (-21, 3.0), (-80, 9.0)):
        assert orc.orc_response_value(("type macroResponse; MT %d;" % mt).encode(), ol.dp(xs)) == pytest.approx(ref, abs=1e-9)
    for mt, ref in ((1, 6.0), (101, 2.0), (18, 1.0), (27, 3.0), (301, 9.0)):           # ENDF MT numbers are translated to the macroscopic ones
        assert orc.orc_response_value(("type macroResponse; MT %d;" % mt).encode(), ol.dp(xs)) == pytest.approx(ref, abs=1e-9)
    assert orc.orc_response_value(b"type fluxResponse;", ol.dp(xs)) == 1.0
