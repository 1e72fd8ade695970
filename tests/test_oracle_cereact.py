"""CPU: the oracle's continuous-energy reaction restatement (oracle/cereact.hpp, cephysics.hpp) against the reference's own
known answers, the ACE fixtures against the reference's files, and the product's host-side card processing against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/IntegrationTestFiles/"
ACE = os.path.join(ROOT, "data", "ace")
have_ref = os.path.exists(REF + "1001JEF311.ace")
DECK = os.path.join(ROOT, "decks", "ce", "pincell")
FILES = [("1001JEF311", 1779), ("92233JEF311", 1), ("52126JEF311", 1), ("91231JEF311", 1), ("91232JEF311", 1)]


def _a(x, t=np.float64):
    return np.ascontiguousarray(x, t)


def test_tabular_pdf_sample_known_answers(orc):
    # NuclearData/NuclearDataStructures/Tests/tabularPdf_test.f90:28-36,125-160 (TOL 1e-9)
    grid = _a([1.0, 2.0, 3.0, 4.0])
    pdf_l, cdf_l = _a([0.0, 0.5, 0.5, 0.0]), _a([0.0, 0.25, 0.75, 1.0])
    pdf_h, cdf_h = _a([0.3, 0.5, 0.2, 0.2]), _a([0.0, 0.3, 0.8, 1.0])
    lin = lambda r: orc.orc_tabpdf_sample(4, ol.dp(grid), ol.dp(pdf_l), ol.dp(cdf_l), 2, r)
    his = lambda r: orc.orc_tabpdf_sample(4, ol.dp(grid), ol.dp(pdf_h), ol.dp(cdf_h), 1, r)
    for r, x in ((0.5, 2.5), (0.7, 2.9), (0.9, 3.367544468), (0.25, 2.0), (0.0, 1.0), (1.0, 4.0)):
        assert lin(r) == pytest.approx(x, abs=1e-9)
    for r, x in ((0.5, 2.4), (0.8, 3.0), (0.0, 1.0), (1.0, 4.0)):
        assert his(r) == pytest.approx(x, abs=1e-9)


def test_endf_table_known_answers(orc):
    # NuclearData/NuclearDataStructures/Tests/endfTable_test.f90:107-130,163-182,228-253 (TOL 1e-6)
    x, y = _a([-1.0, 2.0, 2.5, 3.5]), _a([17.0, -2.0, 0.0, 1.5])
    none = np.zeros(1, np.int32)
    at = lambda v: orc.orc_endftable_at(4, ol.dp(x), ol.dp(y), 0, ol.ip(none), ol.ip(none), v)
    for v, r in ((-1.0, 17.0), (3.5, 1.5), (2.0, -2.0), (0.0, 32.0 / 3), (3.0, 0.75)):
        assert at(v) == pytest.approx(r, abs=1e-6)
    b, f = _a([4], np.int32), _a([1], np.int32)
    at = lambda v: orc.orc_endftable_at(4, ol.dp(x), ol.dp(y), 1, ol.ip(b), ol.ip(f), v)
    for v, r in ((-1.0, 17.0), (3.5, 0.0), (2.0, -2.0), (0.0, 17.0), (3.0, 0.0)):
        assert at(v) == pytest.approx(r, abs=1e-6)
    x, y = _a([-1.0, 2.0, 2.5, 2.5, 3.5]), _a([17.0, 2.0, 0.0, 1.0, 1.5])
    b, f = _a([2, 4, 5], np.int32), _a([4, 3, 5], np.int32)          # logLin, linLog, logLog
    at = lambda v: orc.orc_endftable_at(5, ol.dp(x), ol.dp(y), 3, ol.ip(b), ol.ip(f), v)
    for v, r in ((-1.0, 17.0), (3.5, 1.5), (2.0, 2.0), (0.0, 8.329954), (2.1, 1.5627016), (3.2, 1.346458993)):
        assert at(v) == pytest.approx(r, abs=1e-6)


@pytest.mark.skipif(not have_ref, reason="reference ACE files are not on this box")
def test_ace_fixtures_are_the_reference_cards(orc):
    """data/ace/*.acebin hold exactly what the reference's text cards hold (nuclide built from either is identical)."""
    for name, line in FILES:
        a = orc.orc_ce_nuclide_from_ace((REF + name + ".ace").encode(), line)
        b = orc.orc_ce_nuclide_from_acebin(os.path.join(ACE, name + ".acebin").encode())
        assert a and b, ol.err(orc)
        out = []
        for h in (a, b):
            n, rows, m, kT = C.c_int(), C.c_int(), C.c_double(), C.c_double()
            orc.orc_ce_nuclide_info(h, C.byref(n), C.byref(rows), C.byref(m), C.byref(kT))
            g = np.zeros(n.value); d = np.zeros(n.value * rows.value)
            orc.orc_ce_nuclide_data(h, ol.dp(g), ol.dp(d))
            mts = np.zeros(64, np.int32); fi = np.zeros(64, np.int32)
            k = orc.orc_ce_nuclide_mt_list(h, ol.ip(mts), ol.ip(fi))
            out.append((g, d, mts[:k].copy(), fi[:k].copy(), m.value, kT.value))
        for u, v in zip(out[0], out[1]):
            assert np.array_equal(u, v)
        orc.orc_ce_nuclide_free(a); orc.orc_ce_nuclide_free(b)


def test_reaction_sampling_is_physical(orc):
    """Every reaction of every bundled nuclide samples: |mu| <= 1, phi in [0, 2 pi), E_out > 0; U-233 fission spectrum mean ~ 2 MeV."""
    for name, _ in FILES:
        h = orc.orc_ce_nuclide_from_acebin(os.path.join(ACE, name + ".acebin").encode())
        assert h, ol.err(orc)
        mts = np.zeros(64, np.int32); fi = np.zeros(64, np.int32)
        k = orc.orc_ce_nuclide_mt_list(h, ol.ip(mts), ol.ip(fi))
        out = np.zeros(4)
        for s in range(200):
            assert orc.orc_ce_nuclide_sample(h, 0, 0, 10.0 ** (-9 + 10 * s / 200.0), 1000 + s, ol.dp(out)) > 0, ol.err(orc)
            assert -1.0 <= out[0] <= 1.0 and 0.0 <= out[1] < 2 * np.pi + 1e-12
        for i in range(k):
            for s in range(20):
                assert orc.orc_ce_nuclide_sample(h, 1, i, 19.0, 77 + s, ol.dp(out)) > 0, (name, mts[i], ol.err(orc))
                assert -1.0 <= out[0] <= 1.0 and out[2] > 0.0
            assert orc.orc_ce_nuclide_mt_release(h, i, 19.0) >= 1.0
        if name == "92233JEF311":
            E = []
            for s in range(4000):
                assert orc.orc_ce_nuclide_sample(h, 2, 0, 1.0e-6, 5 + 7 * s, ol.dp(out)) > 0, ol.err(orc)
                E.append(out[2])
            assert 1.8 < np.mean(E) < 2.3
            # (n,2n) of U-233 releases two neutrons, (n,3n) three (neutronScattering_iTest.f90 checks the same for O-16)
            rel = {int(mts[i]): orc.orc_ce_nuclide_mt_release(h, i, 19.0) for i in range(k)}
            assert rel[16] == 2.0 and rel[17] == 3.0 and rel[51] == 1.0
        orc.orc_ce_nuclide_free(h)


def test_ce_eigen_oracle_runs_and_is_reproducible(orc):
    ov = b"pop 600; inactive 2; active 2; seed 99;"
    ks = []
    for mode in (0, 0, 1):
        orc.orc_set_math_mode(mode)
        e = orc.orc_eigen_load(DECK.encode(), ov)
        assert e, ol.err(orc)
        assert orc.orc_eigen_run(e) == 0, ol.err(orc)
        ks.append(orc.orc_eigen_keff0(e))
        n = orc.orc_eigen_bank_size(e)
        E = np.zeros(n); orc.orc_eigen_bank_E(e, ol.dp(E))
        assert n == 600 and np.all(E > 0) and np.all(E <= 20.0)
        orc.orc_eigen_free(e)
    orc.orc_set_math_mode(0)
    assert ks[0] == pytest.approx(ks[1], rel=1e-12)        # same seed, same histories whatever the thread schedule (sums differ in order only)
    assert 0.8 < ks[0] < 1.5 and abs(ks[2] - ks[0]) < 0.2  # sbmath mode: different last bits of log/sin/cos, same physics


def test_host_card_processing_matches_oracle(orc):
    """The product's ACE card -> nuclide (scone_b200/csrc/sb_cekin.cuh ceProcessCard) equals the oracle's independent
    restatement of aceNeutronNuclide%init bit for bit: energy grid, main data, MT order of invertInelastic."""
    import scone_b200
    pp = scone_b200.EigenPhysicsPackage(DECK, "pop 100;", device=-1)
    L = pp.L
    nn, nm = C.c_int32(), C.c_int32()
    L.sbh_ce_info(pp.h, C.byref(nn), C.byref(nm))
    assert nn.value == 5 and nm.value == 2
    # deck order of first appearance: 92233, 52126, 91231, 91232, 1001
    order = ["92233JEF311", "52126JEF311", "91231JEF311", "91232JEF311", "1001JEF311"]
    for i, name in enumerate(order, start=1):
        gs, rows, nmt = C.c_int32(), C.c_int32(), C.c_int32()
        assert L.sbh_ce_card_process(pp.h, i, C.byref(gs), C.byref(rows), C.byref(nmt), None, None, None, None) == 0, pp._err()
        g = np.zeros(gs.value); d = np.zeros(gs.value * rows.value); mt = np.zeros(max(1, nmt.value), np.int32); ak = np.zeros(2)
        assert L.sbh_ce_card_process(pp.h, i, C.byref(gs), C.byref(rows), C.byref(nmt), ol.dp(g), ol.dp(d), mt.ctypes.data_as(C.POINTER(C.c_int32)), ol.dp(ak)) == 0
        h = orc.orc_ce_nuclide_from_acebin(os.path.join(ACE, name + ".acebin").encode())
        n, r, m, kT = C.c_int(), C.c_int(), C.c_double(), C.c_double()
        orc.orc_ce_nuclide_info(h, C.byref(n), C.byref(r), C.byref(m), C.byref(kT))
        assert (n.value, r.value) == (gs.value, rows.value) and (m.value, kT.value) == (ak[0], ak[1])
        og = np.zeros(n.value); od = np.zeros(n.value * r.value)
        orc.orc_ce_nuclide_data(h, ol.dp(og), ol.dp(od))
        assert np.array_equal(og, g) and np.array_equal(od, d), name
        omt = np.zeros(64, np.int32); ofi = np.zeros(64, np.int32)
        k = orc.orc_ce_nuclide_mt_list(h, ol.ip(omt), ol.ip(ofi))
        assert k == nmt.value and np.array_equal(omt[:k], mt[:k])
        orc.orc_ce_nuclide_free(h)
    pp.close()


def test_fixed_source_oracle_and_host_model(orc):
    """fixedSourcePhysicsPackage decks: the oracle runs source batches reproducibly; the product's host model parses the same decks."""
    import scone_b200
    for deck, nbins in (("mg_sphere", 1 + 40), ("ce_sphere", 1 + 200)):
        path = os.path.join(ROOT, "decks", "fixed", deck)
        out = []
        for _ in range(2):
            e = orc.orc_eigen_load(path.encode(), b"pop 1500; cycles 2; seed 3;")
            assert e, ol.err(orc)
            assert orc.orc_eigen_is_fixed(e) == 1
            for _c in range(2):
                assert orc.orc_fixed_cycle(e) == 0, ol.err(orc)
            n = orc.orc_eigen_tally_size(e, 1)
            assert n == nbins
            cs = np.zeros(n); cs2 = np.zeros(n); b = C.c_int()
            orc.orc_eigen_tally(e, 1, ol.dp(cs), ol.dp(cs2), C.byref(b))
            seg, coll, hist = C.c_long(), C.c_long(), C.c_long()
            orc.orc_eigen_stats(e, C.byref(seg), C.byref(coll), C.byref(hist))
            out.append((cs, seg.value, coll.value))
            assert b.value == 2 and hist.value == 3000 and cs.sum() > 0
            orc.orc_eigen_free(e)
        assert out[0][1:] == out[1][1:]
        np.testing.assert_allclose(out[0][0], out[1][0], rtol=1e-12)
        pp = scone_b200.EigenPhysicsPackage(path, "pop 100;", device=-1)
        assert pp.is_fixed_source and pp.is_ce == (deck == "ce_sphere")
        pp.close()


def test_ace_card_header_and_fission_flags(orc):
    # NuclearData/DataDecks/Tests/aceCard_iTest.f90:26-57: Pa-231 (8 precursor groups, prompt + delayed + total nu-bar) and
    # Pa-232 (no delayed data) of JEFF 3.1.1
    ref = {"91231JEF311": (229.05, 2.5852e-08, "91231.03c", [8, 1, 1, 1, 1]),
           "91232JEF311": (230.045, 2.5852e-08, "91232.03c", [0, 1, 1, 0, 1])}
    for name, (aw, tz, zaid, flags) in ref.items():
        out = np.zeros(2); fl = np.zeros(5, np.int32); z = C.create_string_buffer(16)
        assert orc.orc_ace_card_info(os.path.join(ACE, name + ".acebin").encode(), ol.dp(out), ol.ip(fl), z) == 0, ol.err(orc)
        assert out[0] == pytest.approx(aw, rel=1e-6) and out[1] == pytest.approx(tz, rel=1e-6)
        assert z.value.decode().strip() == zaid
        assert fl.tolist() == flags


def test_elastic_scattering_of_te126_is_isotropic_in_cm(orc):
    # NuclearData/Reactions/Tests/elasticScattering_iTest.f90:83-107 (LOCB == 0 card, Te-126): probOf = 1 / (4 pi) for every mu, i.e.
    # mu is sampled as 2 r - 1 and E_out = E_in; :53-56 elastic scattering is in the CM frame with one neutron out
    h = orc.orc_ce_nuclide_from_acebin(os.path.join(ACE, "52126JEF311.acebin").encode())
    assert h, ol.err(orc)
    assert orc.orc_ce_nuclide_elastic_isotropic(h) == 1
    out = np.zeros(4)
    state = 987654321
    n = orc.orc_ce_nuclide_sample(h, 0, 0, 6.7, state, ol.dp(out))
    assert n == 2                                                      # one number for mu, one for phi
    r1 = orc.orc_rng_real(orc.orc_rng_next(state)); r2 = orc.orc_rng_real(orc.orc_rng_next(orc.orc_rng_next(state)))
    assert out[0] == 2.0 * r1 - 1.0 and out[1] == r2 * (2.0 * np.pi) and out[2] == 6.7
    orc.orc_ce_nuclide_free(h)
    h = orc.orc_ce_nuclide_from_acebin(os.path.join(ACE, "92233JEF311.acebin").encode())      # a card with LOCB > 0: tabular mu
    assert orc.orc_ce_nuclide_elastic_isotropic(h) == 0
    orc.orc_ce_nuclide_free(h)
