"""Oracle RNG against the reference's known answers.

RandomNumbers/Tests/RNG_test.f90:15-70 (range; skip(N) == N steps; skip(-(N+1)) returns to start)
RandomNumbers/RNG_class.f90:61-123      (tabulated g^(2^i) mod 2^63, a golden vector in the reference)
"""
import numpy as np

SEED = 0x5C3A84C9
MASK = (1 << 63) - 1
G = 2806196910506780709

# first and last few entries of pow_of_gsq (RNG_class.f90:61-123)
POW_OF_GSQ_HEAD = [4118111548459160921, 6263099103742179569, 5434410004014125793, 3900069298110130625]
POW_OF_GSQ_TAIL = [6917529027641081857, 4611686018427387905, 1, 1, 1]


def test_range(orc):
    s = SEED
    for _ in range(1000):
        s = orc.orc_rng_next(s)
        r = orc.orc_rng_real(s)
        assert 0.0 <= r <= 1.0


def test_lcg_recurrence_matches_python(orc):
    s = SEED
    for _ in range(100):
        s2 = orc.orc_rng_next(s)
        assert s2 == ((G * s) & MASK) + 1 & MASK
        s = s2


def test_pow_table_golden():
    g = G
    tab = []
    for _ in range(63):
        g = (g * g) & MASK
        tab.append(g)
    assert tab[:4] == POW_OF_GSQ_HEAD
    assert tab[-5:] == POW_OF_GSQ_TAIL


def test_skip_forward_and_back(orc):
    N = 13456757
    s = orc.orc_rng_next(SEED)          # r_start
    r_start = orc.orc_rng_real(s)
    for _ in range(N):                   # python loop over the LCG itself (independent of the oracle)
        s = ((G * s) & MASK) + 1 & MASK
    r_end = orc.orc_rng_real(s)
    s2 = orc.orc_rng_skip(SEED, N)
    s2 = orc.orc_rng_next(s2)
    assert orc.orc_rng_real(s2) == r_end
    s2 = orc.orc_rng_skip(s2, -(N + 1))
    s2 = orc.orc_rng_next(s2)
    assert orc.orc_rng_real(s2) == r_start


def test_stride(orc):
    # stride(n) == skip(152917 * n)   RNG_class.f90:305-316
    for n in (1, 7, 100000, 2**31 - 1):
        assert orc.orc_rng_stride(SEED, n) == orc.orc_rng_skip(SEED, 152917 * n)
