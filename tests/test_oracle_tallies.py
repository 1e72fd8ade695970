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
