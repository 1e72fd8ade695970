"""Bank interchange files on the GPU path against the CPU oracle:
  * printSource (eigenPhysicsPackage_class.f90:278-281): particleDungeon%printToFile dumps of the normalised bank after every
    cycle - the binary dumps are compared BYTE FOR BYTE (positions, directions, E, G, broodID, weight), the text dumps value for
    value after parsing;
  * fileSource (ParticleObjects/Source/fileSource_class.f90) driving fixed-source batches from such a dump, text or binary."""
import ctypes as C
import os

import numpy as np
import pytest

import scone_b200
from tests import oracle_lib as ol
from tests.gpu_util import DECK

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MG = os.path.join(ROOT, "decks", "fixed", "mg_sphere")
CE = os.path.join(ROOT, "decks", "fixed", "ce_sphere")


def eigen_deck_of(fixed_deck, tmp_path):
    """The fixed-source deck's geometry and data as an eigenvalue problem (source line dropped, data paths made absolute)."""
    text = open(fixed_deck).read()
    text = text.replace("type fixedSourcePhysicsPackage;", "type eigenPhysicsPackage;")
    text = "\n".join(l for l in text.split("\n") if not l.startswith("source "))
    text = text.replace("../c5g7/xs", os.path.join(ROOT, "decks", "c5g7", "xs"))
    text += "\ninactiveTally { }\nactiveTally { }\n"
    path = str(tmp_path / "eigen_deck")
    open(path, "w").write(text)
    return path


def run_both(orc, deck, ov_common, tmp_path, mode, ninact, nact):
    go, oo = str(tmp_path / "gpu"), str(tmp_path / "orc")
    ov = ov_common + " printSource %d; " % mode
    e = orc.orc_eigen_load(deck.encode(), (ov + "outputFile %s;" % oo).encode())
    assert e, ol.err(orc)
    pp = scone_b200.EigenPhysicsPackage(deck, ov + "outputFile %s;" % go, device=0)
    assert orc.orc_eigen_init_source(e) == 0, ol.err(orc)
    pp.generateInitialState()
    k_o = orc.orc_eigen_keff0(e)
    for cyc in range(ninact + nact):
        active = cyc >= ninact
        pp.cycle(active)
        k_o = orc.orc_eigen_cycle(e, 1 if active else 0, k_o)
        assert not np.isnan(k_o), ol.err(orc)
    pp.close(); orc.orc_eigen_free(e)
    return go, oo


@pytest.mark.parametrize("deck,pop", [("c5g7", 3000), ("ce_pin", 2000)])
def test_binary_source_dumps_are_byte_identical(orc, tmp_path, deck, pop):
    orc.orc_set_math_mode(1)
    try:
        ninact, nact = 3, 2
        go, oo = run_both(orc, DECK[deck], "pop %d; inactive %d; active %d; seed 99;" % (pop, ninact, nact), tmp_path, 2, ninact, nact)
        # the cycle index restarts with the active phase and the files are replaced: 3 files survive, numbered 1..3
        names = sorted(os.listdir(tmp_path))
        assert names == sorted(["%s_source%d_rank0.bin" % (s, i) for s in ("gpu", "orc") for i in (1, 2, 3)])
        for i in (1, 2, 3):
            g = open("%s_source%d_rank0.bin" % (go, i), "rb").read()
            o = open("%s_source%d_rank0.bin" % (oo, i), "rb").read()
            assert len(g) == pop * 80
            assert g == o, "source dump of cycle %d differs" % i
        rows = np.frombuffer(g, dtype=np.float64).reshape(pop, 10)
        assert np.allclose((rows[:, 3:6] ** 2).sum(1), 1.0)
        assert (rows[:, 9] == 1.0).all()
        brood = rows[:, 8]
        assert (np.diff(brood) >= 0).all() and brood.min() >= 1 and brood.max() <= pop       # bank is in broodID order; parents 1..pop
        if deck == "c5g7":
            assert set(np.unique(rows[:, 7])) <= set(range(1, 8)) and (rows[:, 6] == 0).all()
        else:
            assert (rows[:, 7] == 0).all() and (rows[:, 6] > 1e-11).all() and (rows[:, 6] < 20.0).all()
    finally:
        orc.orc_set_math_mode(0)


def test_text_source_dump_reads_back_to_the_same_doubles(orc, tmp_path):
    orc.orc_set_math_mode(1)
    try:
        go, oo = run_both(orc, DECK["c5g7"], "pop 1500; inactive 1; active 1; seed 3;", tmp_path, 1, 1, 1)
        g = np.loadtxt(go + "_source1_rank0.txt"); o = np.loadtxt(oo + "_source1_rank0.txt")
        assert g.shape == (1500, 10)
        assert np.array_equal(g, o)
    finally:
        orc.orc_set_math_mode(0)


def test_print_source_value_is_checked():
    with pytest.raises(scone_b200.EngineError, match="printSource must be 0"):
        scone_b200.EigenPhysicsPackage(DECK["c5g7"], "pop 100; printSource 3;", device=0)


def make_rows(n, seed, ce):
    """A synthetic printToFile dump inside the 5 cm box / 6 cm sphere of the fixed-source decks, with unequal weights."""
    rng = np.random.default_rng(seed)
    rows = np.zeros((n, 10))
    rows[:, 0:3] = rng.uniform(-3.3, 3.3, size=(n, 3))
    u = rng.normal(size=(n, 3)); rows[:, 3:6] = u / np.sqrt((u * u).sum(1))[:, None]
    if ce:
        rows[:, 6] = 10.0 ** rng.uniform(-8, 1, size=n)
    else:
        rows[:, 7] = rng.integers(1, 8, size=n)
    rows[:, 8] = rng.integers(1, 50, size=n)          # broodID: ignored by fileSource
    rows[:, 9] = rng.uniform(0.5, 1.5, size=n)
    return rows


@pytest.mark.parametrize("deck,ce,binary,extra", [
    (MG, False, True, ""),
    (MG, False, False, "transportOperator { type transportOperatorST; }"),
    (CE, True, True, ""),
    (CE, True, False, "transportOperator { type transportOperatorDT; }")])
def test_file_source_batches_against_oracle(orc, tmp_path, deck, ce, binary, extra):
    rows = make_rows(777, 11, ce)
    path = str(tmp_path / ("src.bin" if binary else "src.txt"))
    if binary:
        rows.tofile(path)
    else:
        with open(path, "w") as f:
            for r in rows:
                f.write("  " + " ".join("%.17g" % v for v in r) + "\n")
    ov = "pop 5000; cycles 2; seed 12; %s source { type fileSource; path %s; data %s; binary %d; }" % (extra, path, "ce" if ce else "mg", 1 if binary else 0)
    orc.orc_set_math_mode(1)
    try:
        e = orc.orc_eigen_load(deck.encode(), ov.encode())
        assert e, ol.err(orc)
        pp = scone_b200.FixedSourcePhysicsPackage(deck, ov, device=0)
        segs = colls = 0
        for _ in range(pp.n_active):
            assert orc.orc_fixed_cycle(e) == 0, ol.err(orc)
            res = pp.fixed_cycle()
            segs += res.n_segments; colls += res.n_collisions
            assert pp.rng_state == orc.orc_eigen_rng_state(e)
        seg, coll, hist = C.c_long(), C.c_long(), C.c_long()
        orc.orc_eigen_stats(e, C.byref(seg), C.byref(coll), C.byref(hist))
        assert segs == seg.value and colls == coll.value
        n = orc.orc_eigen_tally_size(e, 1)
        cs, cs2, nb = pp.tally(True)
        ocs = np.zeros(n); ocs2 = np.zeros(n); b = C.c_int()
        orc.orc_eigen_tally(e, 1, ol.dp(ocs), ol.dp(ocs2), C.byref(b))
        assert nb == b.value == 2
        np.testing.assert_allclose(cs, ocs, rtol=1e-10, atol=1e-300)
        np.testing.assert_allclose(cs2, ocs2, rtol=1e-10, atol=1e-300)
        assert cs.sum() > 0
        pp.close(); orc.orc_eigen_free(e)
    finally:
        orc.orc_set_math_mode(0)


def test_file_source_site_outside_geometry_is_the_reference_error(tmp_path):
    rows = make_rows(10, 1, False)
    rows[:, 0] = 7.0                                   # beyond the 5 cm box
    path = str(tmp_path / "out.bin"); rows.tofile(path)
    pp = scone_b200.FixedSourcePhysicsPackage(MG, "pop 100; cycles 1; seed 1; source { type fileSource; path %s; data mg; binary 1; }" % path, device=0)
    with pytest.raises(scone_b200.EngineError, match="outside of geometry"):
        pp.fixed_cycle()
    pp.close()


def test_file_source_data_type_must_match(tmp_path):
    rows = make_rows(10, 1, False)
    path = str(tmp_path / "s.bin"); rows.tofile(path)
    with pytest.raises(scone_b200.EngineError, match="inconsistent with nuclear database"):
        scone_b200.FixedSourcePhysicsPackage(MG, "pop 100; cycles 1; seed 1; source { type fileSource; path %s; binary 1; }" % path, device=0)


def test_eigen_dump_feeds_a_file_source(orc, tmp_path):
    """Round trip of the interchange format: a printSource dump of an eigenvalue run (the subcritical ball of the fixed-source
    deck, driven as an eigenvalue problem) is a valid fileSource for fixed-source batches in the same geometry."""
    orc.orc_set_math_mode(1)
    try:
        ov = "pop 2000; inactive 2; active 0; seed 5; printSource 2; outputFile %s;" % str(tmp_path / "eig")
        pp = scone_b200.EigenPhysicsPackage(eigen_deck_of(MG, tmp_path), ov, device=0)
        pp.generateInitialState()
        pp.cycle(False); pp.cycle(False)
        pp.close()
        path = str(tmp_path / "eig_source2_rank0.bin")
        assert os.path.getsize(path) == 2000 * 80
        fov = "pop 3000; cycles 2; seed 6; source { type fileSource; path %s; data mg; binary 1; }" % path
        fp = scone_b200.FixedSourcePhysicsPackage(MG, fov, device=0)
        e = orc.orc_eigen_load(MG.encode(), fov.encode())
        assert e, ol.err(orc)
        for _ in range(2):
            fp.fixed_cycle(); assert orc.orc_fixed_cycle(e) == 0, ol.err(orc)
        n = orc.orc_eigen_tally_size(e, 1)
        cs, cs2, nb = fp.tally(True)
        ocs = np.zeros(n); ocs2 = np.zeros(n); b = C.c_int()
        orc.orc_eigen_tally(e, 1, ol.dp(ocs), ol.dp(ocs2), C.byref(b))
        np.testing.assert_allclose(cs, ocs, rtol=1e-10, atol=1e-300)
        assert cs.sum() > 0
        fp.close(); orc.orc_eigen_free(e)
    finally:
        orc.orc_set_math_mode(0)


@pytest.mark.parametrize("deck,ce", [(MG, False), (CE, True)])
def test_fixed_source_batches_are_dumped_byte_identical(orc, tmp_path, deck, ce):
    """fixedSourcePhysicsPackage prints the batch its source has just generated: <outputFile>_source<i> (no rank suffix,
    fixedSourcePhysicsPackage_class.f90:191-194); broodID of source particles is 0."""
    orc.orc_set_math_mode(1)
    try:
        go, oo = str(tmp_path / "gpu"), str(tmp_path / "orc")
        ov = "pop 1500; cycles 2; seed 3; printSource 2; "
        e = orc.orc_eigen_load(deck.encode(), (ov + "outputFile %s;" % oo).encode())
        assert e, ol.err(orc)
        pp = scone_b200.FixedSourcePhysicsPackage(deck, ov + "outputFile %s;" % go, device=0)
        for _ in range(2):
            assert orc.orc_fixed_cycle(e) == 0, ol.err(orc)
            pp.fixed_cycle()
        for i in (1, 2):
            g = open("%s_source%d.bin" % (go, i), "rb").read(); o = open("%s_source%d.bin" % (oo, i), "rb").read()
            assert len(g) == 1500 * 80 and g == o
        rows = np.frombuffer(g, dtype=np.float64).reshape(1500, 10)
        assert (rows[:, 8] == 0).all() and (rows[:, 9] == 1.0).all()
        assert ((rows[:, 6] > 0).all() and (rows[:, 7] == 0).all()) if ce else ((rows[:, 6] == 0).all() and (rows[:, 7] >= 1).all())
        pp.close(); orc.orc_eigen_free(e)
    finally:
        orc.orc_set_math_mode(0)
