"""GPU parity of the multi-rank path: a run whose bank is shared between 2 / 3 ranks (one engine each; here they share
cuda:0 and talk over gloo, on a multi-GPU box the same code runs one rank per GPU over NCCL) follows exactly the
histories of the single-rank run -- SCONE's reproducibility property (eigenPhysicsPackage_class.f90:216-218,
particleDungeon_class.f90:431-602): the concatenation of the ranks' banks is bit-identical after every cycle."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import scone_b200
from tests.gpu_util import DECK

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def run_ranks(ws, backend, deck, ov, ninact, nact, out, extra=()):
    port = str(_free_port())
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dist_worker.py"), ROOT, port, str(r), str(ws), backend,
                               deck, ov, str(ninact), str(nact), str(out), *extra], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(ws)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and ("ok %d" % r) in o, o[-3000:]


@pytest.mark.parametrize("deck,pop,ws", [("c5g7", 20001, 2), ("c5g7", 9000, 3), ("slab", 6000, 2), ("ce_pin", 3001, 2)])
def test_ranked_run_reproduces_single_rank_histories(tmp_path, deck, pop, ws):
    ninact, nact = 3, 2
    ov = "pop %d; inactive %d; active %d; seed 31337;" % (pop, ninact, nact)
    run_ranks(ws, "gloo", DECK[deck], ov, ninact, nact, tmp_path)
    pp = scone_b200.EigenPhysicsPackage(DECK[deck], ov, device=0)
    pp.generateInitialState()
    fin = [np.load(os.path.join(tmp_path, "final_r%d.npz" % r)) for r in range(ws)]
    seg_single = []
    for c in range(ninact + nact):
        res = pp.cycle(c >= ninact)
        seg_single.append(res.n_segments)
        parts = [np.load(os.path.join(tmp_path, "bank_c%d_r%d.npz" % (c, r))) for r in range(ws)]
        sizes = [len(p["w"]) for p in parts]
        assert sizes == [scone_b200.distributed.workshare(pop, ws, r)[0] for r in range(ws)]      # load balancing restored the shares
        r1, d1, w1, G1 = pp.bank()
        for key, ref in (("r", r1), ("d", d1), ("w", w1), ("G", G1)):      # 4th array: G (multigroup) or E (continuous energy)
            assert np.array_equal(np.concatenate([p[key] for p in parts]), ref), "bank differs after cycle %d (%s)" % (c, key)
        for f in fin:
            assert f["k"][c] == pytest.approx(pp.k, rel=1e-12)          # same k on every rank (sums differ in rounding only)
        assert sum(int(f["seg"][c]) for f in fin) == res.n_segments
    # user tallies: per-rank accumulation, summed at the end (scoreMemory%collectDistributed)
    cs, cs2, nb = pp.tally(True)
    assert sum(int(f["nb"]) for f in fin) == ws * nb           # batch counts are summed over ranks
    # scoreMemory%collectDistributed: the master ends with the sums over ranks, the others with a batch count of 0
    col = [np.load(os.path.join(tmp_path, "collected_r%d.npz" % r)) for r in range(ws)]
    assert int(col[0]["nb"]) == ws * nb and all(int(c["nb"]) == 0 for c in col[1:])
    np.testing.assert_allclose(col[0]["cs"], sum(f["cs"] for f in fin), rtol=1e-13, atol=1e-300)
    np.testing.assert_allclose(col[0]["cs2"], sum(f["cs2"] for f in fin), rtol=1e-13, atol=1e-300)
    if len(cs) and deck == "c5g7":
        # the un-normalised fission map is additive over ranks (a tally with `norm` is normalised per rank, as in SCONE)
        tot = sum(f["cs"] for f in fin)
        np.testing.assert_allclose(tot, cs, rtol=1e-10, atol=1e-300)
    pp.close()


@pytest.mark.parametrize("exchange", ["gloo", "peer"])
def test_ranked_user_keff_clerks_hold_the_ranks_own_scores(tmp_path, exchange):
    """A keffImplicitClerk named in the deck's own activeTally is not synchronised per cycle in the reference (only the
    attachment tally is mpiSync): every rank accumulates ITS scores and collectDistributed adds the ranks up at the end.
    The four score bins of the ranks must then add up to the single-rank run's; with the rank-global sums in them the total
    would come out n_ranks times too large."""
    pop, ws, ninact, nact = 12000, 2, 2, 3
    ov = ("pop %d; inactive %d; active %d; seed 777; activeTally { kimp { type keffImplicitClerk; } "
          "fis { type collisionClerk; response (f); f { type macroResponse; MT -6; } } }" % (pop, ninact, nact))
    run_ranks(ws, "gloo", DECK["c5g7"], ov, ninact, nact, tmp_path, extra=(("peer",) if exchange == "peer" else ()))
    pp = scone_b200.EigenPhysicsPackage(DECK["c5g7"], ov, device=0)
    pp.generateInitialState()
    for c in range(ninact + nact):
        pp.cycle(c >= ninact)
    cs, cs2, nb = pp.tally(True)
    fin = [np.load(os.path.join(tmp_path, "final_r%d.npz" % r)) for r in range(ws)]
    tot = sum(f["cs"] for f in fin)
    np.testing.assert_allclose(tot[:4], cs[:4], rtol=1e-11)            # IMP_PROD, IMP_ABS, SCATTER_PROD, ANA_LEAK: additive over ranks
    np.testing.assert_allclose(tot[5:], cs[5:], rtol=1e-10, atol=1e-300)   # the collision clerk behind it
    for f in fin:                                                       # each rank's own k estimate is a k-eff, not a multiple of one
        assert 0.5 < f["cs"][4] / nb < 2.0
    pp.close()


def test_ranked_source_dumps_concatenate_to_the_single_rank_dump(tmp_path):
    """printSource with several ranks: every rank writes <outputFile>_source<i>_rank<r> AFTER the load balancing
    (eigenPhysicsPackage_class.f90:275-281); the rank files in rank order hold the sites of the single-rank dump.  broodID is the
    parent's index within its rank's share, carried through the exchange (MPI_PARTICLE_STATE holds it, mpi_func.f90:88-117)."""
    pop, ws, ninact, nact = 7001, 2, 2, 1
    base = "pop %d; inactive %d; active %d; seed 99; printSource 2; " % (pop, ninact, nact)
    run_ranks(ws, "gloo", DECK["c5g7"], base + "outputFile %s;" % (tmp_path / "ranked"), ninact, nact, tmp_path)
    pp = scone_b200.EigenPhysicsPackage(DECK["c5g7"], base + "outputFile %s;" % (tmp_path / "single"), device=0)
    pp.generateInitialState()
    for c in range(ninact + nact):
        pp.cycle(c >= ninact)
    pp.close()
    shares = [scone_b200.distributed.workshare(pop, ws, r)[0] for r in range(ws)]
    for i in (1, 2):
        one = np.fromfile(str(tmp_path / ("single_source%d_rank0.bin" % i))).reshape(-1, 10)
        parts = [np.fromfile(str(tmp_path / ("ranked_source%d_rank%d.bin" % (i, r)))).reshape(-1, 10) for r in range(ws)]
        assert [len(p) for p in parts] == shares
        cat = np.concatenate(parts)
        cols = [0, 1, 2, 3, 4, 5, 6, 7, 9]
        assert np.array_equal(cat[:, cols], one[:, cols])
        for r, p in enumerate(parts):
            assert p[:, 8].min() >= 1 and p[:, 8].max() <= max(shares)
        # rank 0 starts the global history numbering, so its own sites keep the brood IDs of the single-rank run
        kept = min(len(parts[0]), len(one))
        same = parts[0][:kept, 8] == one[:kept, 8]
        assert same[: kept // 2].all()


@pytest.mark.parametrize("deck,pop,ws", [("c5g7", 20001, 2), ("c5g7", 9000, 3), ("ce_pin", 3001, 2)])
def test_peer_memory_exchange_reproduces_single_rank_histories(tmp_path, deck, pop, ws):
    """The ranks exchange score sums, bank sizes and the balanced sites by storing into each other's memory (CUDA IPC; NVLink on a
    multi-GPU node) with no collective and one host synchronisation per cycle: same banks, bit for bit, as the single-rank run."""
    ninact, nact = 3, 2
    ov = "pop %d; inactive %d; active %d; seed 4242;" % (pop, ninact, nact)
    run_ranks(ws, "gloo", DECK[deck], ov, ninact, nact, tmp_path, extra=("peer",))
    pp = scone_b200.EigenPhysicsPackage(DECK[deck], ov, device=0)
    pp.generateInitialState()
    fin = [np.load(os.path.join(tmp_path, "final_r%d.npz" % r)) for r in range(ws)]
    for c in range(ninact + nact):
        res = pp.cycle(c >= ninact)
        parts = [np.load(os.path.join(tmp_path, "bank_c%d_r%d.npz" % (c, r))) for r in range(ws)]
        assert [len(p["w"]) for p in parts] == [scone_b200.distributed.workshare(pop, ws, r)[0] for r in range(ws)]
        for key, ref in zip(("r", "d", "w", "G"), pp.bank()):
            assert np.array_equal(np.concatenate([p[key] for p in parts]), ref), "bank differs after cycle %d (%s)" % (c, key)
        for f in fin:
            assert f["k"][c] == pytest.approx(pp.k, rel=1e-12)
        assert sum(int(f["seg"][c]) for f in fin) == res.n_segments
    assert len({int(f["rng"][0]) for f in fin}) == ws          # every rank keeps its own stream offset
    pp.close()


def test_peer_exchange_times_out_when_a_rank_does_not_show_up(tmp_path):
    """A rank that never posts its cycle data must not hang the others: the waiting kernels give up after the time-out and the call
    returns the error (the spin is bounded on the device, nothing is left running)."""
    run_ranks(2, "gloo", DECK["c5g7"], "pop 4000; inactive 1; active 1; seed 1;", 1, 1, tmp_path, extra=("peer_absent",))


@pytest.mark.parametrize("extra", [(), ("peer",)])
def test_ranked_surface_tracking_with_entropy_clerk(tmp_path, extra):
    """Several ranks with surface tracking, a Shannon-entropy clerk and a very subcritical system (decks/mg/can, k = 0.18: every cycle the
    bank is multiplied up with copies by normSize_Repr, so most sites change rank in the balancing): same banks as one rank, over the
    process group and over peer memory."""
    pop, ws, ninact, nact = 6002, 3, 2, 2
    tal = ("%s { ent { type shannonEntropyClerk; cycles 2; map { type spaceMap; axis z; grid lin; min -6.0; max 6.0; N 6; } } "
           "fl { type collisionClerk; response (f); f { type fluxResponse; } } }")
    ov = "pop %d; inactive %d; active %d; seed 77; transportOperator { type transportOperatorST; cache 1; } %s %s" % (
        pop, ninact, nact, tal % "inactiveTally", tal % "activeTally")
    run_ranks(ws, "gloo", DECK["can"], ov, ninact, nact, tmp_path, extra=extra)
    pp = scone_b200.EigenPhysicsPackage(DECK["can"], ov, device=0)
    pp.generateInitialState()
    fin = [np.load(os.path.join(tmp_path, "final_r%d.npz" % r)) for r in range(ws)]
    for c in range(ninact + nact):
        res = pp.cycle(c >= ninact)
        parts = [np.load(os.path.join(tmp_path, "bank_c%d_r%d.npz" % (c, r))) for r in range(ws)]
        assert [len(p["w"]) for p in parts] == [scone_b200.distributed.workshare(pop, ws, r)[0] for r in range(ws)]
        for key, ref in zip(("r", "d", "w", "G"), pp.bank()):
            assert np.array_equal(np.concatenate([p[key] for p in parts]), ref), "bank differs after cycle %d (%s)" % (c, key)
        for f in fin:
            assert f["k"][c] == pytest.approx(pp.k, rel=1e-12)
        assert sum(int(f["seg"][c]) for f in fin) == res.n_segments
    cs, cs2, nb = pp.tally(True)
    tot = sum(f["cs"] for f in fin)
    assert tot[-1] == pytest.approx(cs[-1], rel=1e-10)                  # the flux clerk is additive over ranks
    ent_single = cs[7:9]
    assert (ent_single > 0).all()
    for f in fin:                                                       # every rank scores the entropy of ITS share of the bank (no mpiSync)
        assert (f["cs"][7:9] > 0).all() and (f["cs"][:7] == 0).all()
    pp.close()


@pytest.mark.parametrize("deck,ws", [("mg_sphere", 2), ("ce_sphere", 3)])
def test_ranked_fixed_source_is_the_single_rank_run_split(tmp_path, deck, ws):
    """Fixed-source batches over several ranks (fixedSourcePhysicsPackage_class.f90:131,347): the ranks' histories together are the
    histories of the single-rank run (same random streams), so segment counts add up exactly and un-normalised tallies to rounding."""
    path = os.path.join(ROOT, "decks", "fixed", deck)
    ov = "pop 4001; cycles 3; seed 17; tally { f { type collisionClerk; map { type spaceMap; axis x; grid lin; min -5.0; max 5.0; N 10; } response (fl); fl { type fluxResponse; } } }"
    port = str(_free_port())
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "fixed_dist_worker.py"), ROOT, port, str(r), str(ws), path, ov, str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(ws)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and ("ok %d" % r) in o, o[-3000:]
    fin = [np.load(os.path.join(tmp_path, "fixed_r%d.npz" % r)) for r in range(ws)]
    assert [int(f["pop"]) for f in fin] == [scone_b200.distributed.workshare(4001, ws, r)[0] for r in range(ws)]
    pp = scone_b200.FixedSourcePhysicsPackage(path, ov, device=0)
    segs = sum(pp.fixed_cycle().n_segments for _ in range(3))
    cs, cs2, nb = pp.tally(True)
    assert sum(int(f["seg"]) for f in fin) == segs
    np.testing.assert_allclose(sum(f["cs"] for f in fin), cs, rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(fin[0]["ccs"], cs, rtol=1e-10, atol=1e-300)          # collectDistributed: the master holds the sums
    assert int(fin[0]["cnb"]) == ws * nb
    pp.close()


@pytest.mark.parametrize("deck,pop,ws,extra", [
    ("can", 402, 4, ()), ("can", 402, 4, ("peer",)),        # k = 0.18: strongly fluctuating per-rank fission banks (2 x share is the capacity, as in the reference)
    ("c5g7", 130, 4, ()), ("c5g7", 130, 4, ("peer",))])
def test_ranked_runs_with_tiny_shares(tmp_path, deck, pop, ws, extra):
    """A few histories per rank: fission banks of single ranks can be empty, the resampling multiplies sites up with copies and nearly every
    site changes rank in the balancing; both exchanges must reproduce the single-rank banks."""
    ninact, nact = 2, 1
    ov = "pop %d; inactive %d; active %d; seed 55; inactiveTally { } activeTally { }" % (pop, ninact, nact)
    run_ranks(ws, "gloo", DECK[deck], ov, ninact, nact, tmp_path, extra=extra)
    pp = scone_b200.EigenPhysicsPackage(DECK[deck], ov, device=0)
    pp.generateInitialState()
    for c in range(ninact + nact):
        pp.cycle(c >= ninact)
        parts = [np.load(os.path.join(tmp_path, "bank_c%d_r%d.npz" % (c, r))) for r in range(ws)]
        assert [len(p["w"]) for p in parts] == [scone_b200.distributed.workshare(pop, ws, r)[0] for r in range(ws)]
        for key, ref in zip(("r", "d", "w", "G"), pp.bank()):
            assert np.array_equal(np.concatenate([p[key] for p in parts]), ref), "bank differs after cycle %d (%s)" % (c, key)
    pp.close()
