"""Round timings of the history kernel per number of histories alive in the warp (needs a library built with -DSB_PROFILE_ROUNDS)."""
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
names = ["lone(window)", "lone(no window)", "2-4 alive", "5-16", "17-31", "32"]
for cyc in range(3):
    res = pp.cycle(True)
    raw = np.zeros(8 * 1024 * 40, np.int64)
    assert L.sb_profile_rounds(eng, raw.ctypes.data) == 0
    buf = raw[:8 * 1024 * 16].reshape(8 * 1024, 16); reg = raw[8 * 1024 * 16:].reshape(8 * 1024, 6, 4)
    tot = buf[:, 12]; w = int(np.argmax(tot))
    print("cycle %d: longest history %d flights; slowest warp %d ran %.3f ms (at 1.965 GHz)" % (cyc, res.max_history_segments, w, tot[w] / 1.965e6))
    for i, nm in enumerate(names):
        T, N = buf[w, i], buf[w, 6 + i]
        Ta, Na = buf[:, i].sum(), buf[:, 6 + i].sum()
        print("   %-16s slowest warp: %5d rounds, %7.0f cycles/round (%.3f ms) | all warps: %8d rounds, %7.0f cycles/round | slowest warp per round: flight+geometry %5.0f, scoring %5.0f, channel+slots %5.0f, sites+scattering %5.0f" % (
            nm, N, T / max(N, 1), T / 1.965e6, Na, Ta / max(Na, 1), *(reg[w, i] / max(N, 1))))
