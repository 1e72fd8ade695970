"""Where a long history's round goes (needs a library built with -DSB_PROFILE_ROUNDS: build.sh -DSB_PROFILE_ROUNDS -o ../libscone_b200_prof.so,
run with SB_LIBRARY=.../libscone_b200_prof.so). Per lane the kernel accumulates clock64() differences per region of the rounds in which
the lane's history was alone in its warp (with / without the draw window) or shared it with 1 - 3 others; this prints the lane that had
the most rounds alone (the cycle's longest history) and the sum over all lanes."""
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, ctypes as C
import scone_b200
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
pop = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
pp = scone_b200.EigenPhysicsPackage(os.path.join(ROOT, "decks/c5g7/c5g7_2d"), "pop %d; inactive 3; active 30; seed 7;" % pop, device=0)
pp.generateInitialState(); pp.cycles(False, 3)
L, eng = pp.L, pp.engine
L.sb_profile_rounds.argtypes = [C.c_void_p, C.c_void_p]
regions = ["refill+window", "draw+move", "placement", "XS+accept+scoring", "channel", "site slots", "scatter draw+group", "rotate+update+end"]
kinds = ["alone, window", "alone, no window", "with 1-3 others"]
for cyc in range(2):
    res = pp.cycle(True)
    raw = np.zeros(27 * 148 * 384, np.int64)
    assert L.sb_profile_rounds(eng, raw.ctypes.data) == 0
    d = raw.reshape(-1, 27)
    ln = int(np.argmax(d[:, 24] + d[:, 25]))
    print("cycle %d: longest history %d flights" % (cyc, res.max_history_segments))
    for k, nm in enumerate(kinds):
        n1, nA = d[ln, 24 + k], d[:, 24 + k].sum()
        if nA == 0: continue
        print("  %-18s lane of the longest history: %4d rounds, %6.0f cycles/round | all lanes: %8d rounds, %6.0f cycles/round" % (
            nm, n1, d[ln, 8 * k:8 * k + 8].sum() / max(n1, 1), nA, d[:, 8 * k:8 * k + 8].sum() / nA))
        print("       " + ", ".join("%s %.0f (%.0f)" % (r, d[ln, 8 * k + j] / max(n1, 1), d[:, 8 * k + j].sum() / nA) for j, r in enumerate(regions)))
