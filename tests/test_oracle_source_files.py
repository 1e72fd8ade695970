"""Oracle-only checks of the bank interchange files (no GPU): printToFile layout (particleDungeon_class.f90:1077-1112) and
fileSource sampling (fileSource_class.f90:151-196: row int(rand*N)+1, weight / group from the row, broodID ignored)."""
import ctypes as C
import os

import numpy as np

from tests import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MG = os.path.join(ROOT, "decks", "fixed", "mg_sphere")
STRIDE = 152917


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


def lcg_skip(seed, n):
    """RNG%skip (RNG_class.f90): state after n draws of the 63-bit LCG."""
    g, c, mask = 2806196910506780709, 1, (1 << 63) - 1
    n &= mask
    gn, cn = 1, 0
    while n:
        if n & 1:
            gn = gn * g & mask; cn = (cn * g + c) & mask
        c = (g + 1) * c & mask; g = g * g & mask; n >>= 1
    return (gn * seed + cn) & mask


def lcg_get(state):
    s = (2806196910506780709 * state + 1) & ((1 << 63) - 1)
    return s, s / float(1 << 63)


def test_print_source_layout_and_brood_order(orc, tmp_path):
    out = str(tmp_path / "o")
    ov = "pop 500; inactive 2; active 1; seed 5; printSource 2; outputFile %s;" % out
    e = orc.orc_eigen_load(eigen_deck_of(MG, tmp_path).encode(), ov.encode())
    assert e, ol.err(orc)
    assert orc.orc_eigen_init_source(e) == 0
    k = orc.orc_eigen_keff0(e)
    for c in range(3):
        k = orc.orc_eigen_cycle(e, 1 if c == 2 else 0, k)
    # bank after the last cycle equals the last dump written (active cycle 1 replaced inactive cycle 1)
    n = 500
    r = np.zeros((n, 3)); d = np.zeros((n, 3)); w = np.zeros(n); G = np.zeros(n, np.int32); brood = np.zeros(n, np.int32)
    assert orc.orc_eigen_bank(e, ol.dp(r), ol.dp(d), ol.dp(w), G.ctypes.data_as(C.POINTER(C.c_int)), brood.ctypes.data_as(C.POINTER(C.c_int))) == n
    rows = np.fromfile(out + "_source1_rank0.bin").reshape(n, 10)
    assert np.array_equal(rows[:, 0:3], r) and np.array_equal(rows[:, 3:6], d)
    assert np.array_equal(rows[:, 6], np.zeros(n)) and np.array_equal(rows[:, 7], G) and np.array_equal(rows[:, 8], brood) and np.array_equal(rows[:, 9], w)
    assert sorted(os.listdir(tmp_path)) == ["eigen_deck", "o_source1_rank0.bin", "o_source2_rank0.bin"]
    orc.orc_eigen_free(e)


def test_file_source_text_and_binary_give_the_same_batches_and_follow_the_rng(orc, tmp_path):
    rng = np.random.default_rng(4)
    n = 321
    rows = np.zeros((n, 10))
    rows[:, 0:3] = rng.uniform(-4.0, 4.0, size=(n, 3))
    u = rng.normal(size=(n, 3)); rows[:, 3:6] = u / np.sqrt((u * u).sum(1))[:, None]
    rows[:, 7] = rng.integers(1, 8, size=n); rows[:, 8] = 5; rows[:, 9] = rng.uniform(0.5, 1.5, size=n)
    pb, pt = str(tmp_path / "s.bin"), str(tmp_path / "s.txt")
    rows.tofile(pb)
    with open(pt, "w") as f:
        for row in rows:
            f.write(" ".join("%.17g" % v for v in row) + "\n")
    res = []
    for path, binary in ((pb, 1), (pt, 0)):
        ov = "pop 400; cycles 1; seed 77; source { type fileSource; path %s; data mg; binary %d; }" % (path, binary)
        e = orc.orc_eigen_load(MG.encode(), ov.encode())
        assert e, ol.err(orc)
        assert orc.orc_fixed_cycle(e) == 0, ol.err(orc)
        m = orc.orc_eigen_tally_size(e, 1)
        cs = np.zeros(m); cs2 = np.zeros(m); b = C.c_int()
        orc.orc_eigen_tally(e, 1, ol.dp(cs), ol.dp(cs2), C.byref(b))
        res.append(cs)
        orc.orc_eigen_free(e)
    np.testing.assert_allclose(res[0], res[1], rtol=1e-12)        # thread order of the score sums
    assert res[0].sum() > 0
    # which rows the batch used: particle i draws row int(rand*N) with the package RNG skipped by stride*i
    picks = [int(lcg_get(lcg_skip(77, STRIDE * i))[1] * n) for i in range(1, 401)]
    assert 0 <= min(picks) and max(picks) < n and len(set(picks)) > 200


def test_file_source_row_outside_geometry_is_fatal(orc, tmp_path):
    rows = np.zeros((4, 10)); rows[:, 0] = 7.0; rows[:, 3] = 1.0; rows[:, 7] = 1; rows[:, 9] = 1.0
    path = str(tmp_path / "out.bin"); rows.tofile(path)
    e = orc.orc_eigen_load(MG.encode(), ("pop 50; cycles 1; seed 1; source { type fileSource; path %s; data mg; binary 1; }" % path).encode())
    assert e, ol.err(orc)
    assert orc.orc_fixed_cycle(e) != 0
    assert "outside of geometry" in ol.err(orc)
    orc.orc_eigen_free(e)
