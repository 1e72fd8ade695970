"""Oracle particleDungeon against ParticleObjects/Tests/particleDungeon_test.f90
(normSize sizes 5 and 15: :247-248,292-293; sortByBroodID: :307-340) plus properties the
reference relies on: the transcribed in-place permutation is a stable sort, and the
reproducible resampling keeps exactly the sites whose random number exceeds the threshold.
"""
import numpy as np

from tests import oracle_lib as ol

G = 2806196910506780709
MASK = (1 << 63) - 1


def lcg_reals(state, n):
    out = []
    for _ in range(n):
        state = ((G * state) & MASK) + 1 & MASK
        out.append(float(state) * 2.0 ** -63)
    return np.array(out)


def test_sort_by_brood_is_stable(orc):
    rng = np.random.default_rng(1)
    for n in (1, 10, 1000):
        brood = rng.integers(1, max(2, n // 3), n).astype(np.int32)
        tag = np.arange(n, dtype=np.int32)
        assert orc.orc_dungeon_sort(n, ol.ip(brood), ol.ip(tag)) == 0
        ref = np.argsort(brood, kind="stable")
        np.testing.assert_array_equal(tag, ref)
    # reverse order case of the reference test
    brood = np.arange(10, 0, -1).astype(np.int32)
    tag = np.arange(10, dtype=np.int32)
    orc.orc_dungeon_sort(10, ol.ip(brood), ol.ip(tag))
    np.testing.assert_array_equal(tag, np.arange(9, -1, -1))


def test_norm_size_down(orc):
    n, tot = 10, 5
    brood = np.arange(1, n + 1, dtype=np.int32)
    tag = np.arange(n, dtype=np.int32)
    m = orc.orc_dungeon_norm_size(n, ol.ip(brood), ol.ip(tag), n, tot, 0x5C3A84C9)
    assert m == tot
    rn = lcg_reals(0x5C3A84C9, n)
    thr = np.sort(rn)[n - tot - 1]          # (excess)-th smallest
    np.testing.assert_array_equal(tag[:m], np.nonzero(rn > thr)[0])


def test_norm_size_up(orc):
    n, tot = 10, 15
    brood = np.arange(1, n + 1, dtype=np.int32)
    tag = np.zeros(2 * tot, dtype=np.int32); tag[:n] = np.arange(n)
    m = orc.orc_dungeon_norm_size(n, ol.ip(brood), ol.ip(tag), 2 * tot, tot, 0x5C3A84C9)
    assert m == tot
    rn = lcg_reals(0x5C3A84C9, n)
    thr = np.sort(rn)[tot - n - 1]
    dup = np.nonzero(rn <= thr)[0]
    ref = np.sort(np.concatenate([np.arange(n), dup]), kind="stable")
    np.testing.assert_array_equal(tag[:m], ref)


def test_norm_size_up_with_copies(orc):
    # massive undersampling: 4 sites -> 11 (2 full copies + 3 duplicates)
    n, tot = 4, 11
    brood = np.array([2, 2, 5, 7], dtype=np.int32)
    tag = np.zeros(64, dtype=np.int32); tag[:n] = np.arange(n)
    m = orc.orc_dungeon_norm_size(n, ol.ip(brood), ol.ip(tag), 64, tot, 999)
    assert m == tot
    rn = lcg_reals(999, n)
    thr = np.sort(rn)[3 - 1]
    dup = np.nonzero(rn <= thr)[0]
    seq = np.concatenate([np.arange(n), np.arange(n), dup])
    ref = seq[np.argsort(brood[seq], kind="stable")]
    np.testing.assert_array_equal(tag[:m], ref)


def test_heap_queue_known_answers(orc):
    # DataStructures/Tests/heapQueue_test.f90 testBelowMaximum :8-19, testAboveMaximum :21-33 (the bounded max-heap normSize_Repr
    # uses to find the k-th smallest random number)
    import ctypes as C
    import numpy as np
    from tests import oracle_lib as ol
    n = C.c_int()
    seq = np.array([2.0, 3.0, 1.0, 4.0, 5.0])
    assert orc.orc_heap_queue(8, len(seq), ol.dp(seq), 0, C.byref(n)) == 5.0 and n.value == 5
    seq = np.array([1000.0, 2.0, 3.0, 1.0, 1.4, 5.0])
    assert orc.orc_heap_queue(3, len(seq), ol.dp(seq), 1, C.byref(n)) == 2.0 and n.value == 3
