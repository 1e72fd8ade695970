"""GPU parity of fixed-source calculations (fixedSourcePhysicsPackage: pointSource, secondaries followed within the history from a
private last-in-first-out buffer) against the CPU oracle: with the shared deterministic log/sin/cos both sides follow the same
histories, so segment and collision counts are EQUAL and every tally bin agrees to summation-order rounding."""
import ctypes as C
import os

import numpy as np
import pytest

import scone_b200
from tests import oracle_lib as ol

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MG = os.path.join(ROOT, "decks", "fixed", "mg_sphere")
CE = os.path.join(ROOT, "decks", "fixed", "ce_sphere")


@pytest.mark.parametrize("deck,ov", [
    (MG, "pop 6000; cycles 3; seed 5;"),
    (MG, "pop 4000; cycles 2; seed 6; transportOperator { type transportOperatorST; } source { type pointSource; r (0.1 0.2 0.3); G 2; dir (0.0 1.0 1.0); }"),
    (MG, "pop 4000; cycles 2; seed 7; transportOperator { type transportOperatorHT; cutoff 0.7; } source { type pointSource; r (-1.0 0.0 0.5); probG (0.5 0.2 0.1 0.1 0.05 0.03 0.02); }"),
    (CE, "pop 6000; cycles 3; seed 8;"),
    (CE, "pop 4000; cycles 2; seed 9; transportOperator { type transportOperatorDT; } source { type pointSource; r (1.0 1.0 0.0); E 2.0; }"),
    # materialSource (materialSource_class.f90): uniform in the fuel by rejection over the geometry's bounding box / a given box
    (MG, "pop 5000; cycles 2; seed 10; source { type materialSource; mat UO2; data mg; G 3; }"),
    (MG, "pop 4000; cycles 2; seed 11; source { type materialSource; mat water; data mg; G 1; boundingBox (-5.0 -5.0 -5.0 5.0 5.0 0.0); }"),
    (CE, "pop 4000; cycles 2; seed 12; source { type materialSource; mat fuel; E 1.5; }")])
def test_fixed_source_batches_against_oracle(orc, deck, ov):
    orc.orc_set_math_mode(1)
    try:
        e = orc.orc_eigen_load(deck.encode(), ov.encode())
        assert e, ol.err(orc)
        assert orc.orc_eigen_is_fixed(e) == 1
        pp = scone_b200.FixedSourcePhysicsPackage(deck, ov, device=0)
        assert pp.is_fixed_source
        segs = colls = 0
        for _ in range(pp.n_active):
            assert orc.orc_fixed_cycle(e) == 0, ol.err(orc)
            res = pp.fixed_cycle()
            segs += res.n_segments; colls += res.n_collisions
            assert pp.rng_state == orc.orc_eigen_rng_state(e)
        seg, coll, hist = C.c_long(), C.c_long(), C.c_long()
        orc.orc_eigen_stats(e, C.byref(seg), C.byref(coll), C.byref(hist))
        assert segs == seg.value and colls == coll.value           # same histories, including every secondary
        assert colls > hist.value * 0.5
        n = orc.orc_eigen_tally_size(e, 1)
        cs, cs2, nb = pp.tally(True)
        ocs = np.zeros(n); ocs2 = np.zeros(n); b = C.c_int()
        orc.orc_eigen_tally(e, 1, ol.dp(ocs), ol.dp(ocs2), C.byref(b))
        assert len(cs) == n and nb == b.value == pp.n_active
        np.testing.assert_allclose(cs, ocs, rtol=1e-10, atol=1e-300)
        np.testing.assert_allclose(cs2, ocs2, rtol=1e-10, atol=1e-300)
        assert cs.sum() > 0
        pp.close(); orc.orc_eigen_free(e)
    finally:
        orc.orc_set_math_mode(0)


def test_secondary_buffer_overflow_is_the_reference_error():
    """A buffer of one entry cannot hold the sites of a fission with nu > 2: 'Run out of space for particles'."""
    pp = scone_b200.FixedSourcePhysicsPackage(MG, "pop 20000; cycles 1; seed 5; buffer 1;", device=0)
    with pytest.raises(scone_b200.EngineError, match="Run out of space for particles"):
        pp.fixed_cycle()
    pp.close()


def test_material_source_without_its_material_in_the_box_is_the_reference_error():
    ov = "pop 100; cycles 1; seed 1; source { type materialSource; mat UO2; data mg; G 1; boundingBox (4.0 4.0 4.0 4.9 4.9 4.9); }"
    pp = scone_b200.FixedSourcePhysicsPackage(MG, ov, device=0)
    with pytest.raises(scone_b200.EngineError, match="Infinite loop in sampling source"):
        pp.fixed_cycle()
    pp.close()
