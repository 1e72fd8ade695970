"""Latency of one flight + collision round for a history that runs ALONE in its warp: cycles of one history each
(the kernel time is that history's dependent chain), CUDA-event time of the history kernel / flights."""
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import scone_b200, ctypes as C
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
for pop in (1, 2, 4, 8, 16, 32):
    pp = scone_b200.EigenPhysicsPackage(os.path.join(ROOT, "decks/c5g7/c5g7_2d"), "pop %d; inactive 3; active 400; seed 11;" % pop, device=0)
    pp.generateInitialState()
    pp.cycles(False, 3)
    L, eng = pp.L, pp.engine
    L.sb_profile_enable(eng, 1)
    seg = 0; n = 300
    for _ in range(n):
        res = pp.cycle(True); seg += res.n_segments
    ms = C.c_double(); a = C.c_int64(); b = C.c_int64(); c = C.c_int64()
    L.sb_profile_read(eng, C.byref(ms), C.byref(a), C.byref(b), C.byref(c))
    print("pop %3d: kernel %.4f ms/cycle, segments/cycle %.1f, us per segment (all lanes of one warp) %.3f" % (pop, ms.value / n, seg / n, 1e3 * ms.value / max(1, seg)))
    pp.close()
