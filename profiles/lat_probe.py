import os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import scone_b200, ctypes as C
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
for pop in (32, 256, 2000, 20000, 100000):
    pp = scone_b200.EigenPhysicsPackage(os.path.join(ROOT, "decks/c5g7/c5g7_2d"), "pop %d; inactive 3; active 30; seed 7;" % pop, device=0)
    pp.generateInitialState()
    pp.cycles(False, 3)
    L, eng = pp.L, pp.engine
    tot_ms = 0.0; tot_max = 0; tot_seg = 0
    L.sb_profile_enable(eng, 1)
    for _ in range(20):
        res = pp.cycle(True)
        tot_max += res.max_history_segments; tot_seg += res.n_segments
    ms = C.c_double(); a = C.c_int64(); b = C.c_int64(); c = C.c_int64()
    L.sb_profile_read(eng, C.byref(ms), C.byref(a), C.byref(b), C.byref(c))
    print("pop %6d: kernel %.3f ms/cycle, longest history %.0f flights (reported >= 256), us per flight of the longest %.2f, seg/hist %.1f" % (
        pop, ms.value / 20, tot_max / 20, 1e3 * (ms.value / 20) / max(1, tot_max / 20), tot_seg / 20 / pop))
    pp.close()
