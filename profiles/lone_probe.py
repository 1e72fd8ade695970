"""Latency of one flight + collision round of a history that is alone in its warp: cycles of FOUR histories, one per warp (SB_LANES=1: only
lane 0 of a warp takes histories), so that every history is alone from its first flight and 8 warps of the CTA have nothing to do.
Prints the sum of the kernel times over the sum of the longest histories.  SB_ASSIST=0 / 1 compares the plain lone mode with helpers."""
import os, sys
os.environ.setdefault("SB_LANES", "1")
os.environ["SB_MAXSEG_MIN"] = "0"
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import scone_b200, ctypes as C
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
pop = int(sys.argv[1]) if len(sys.argv) > 1 else 4
pp = scone_b200.EigenPhysicsPackage(os.path.join(ROOT, "decks/c5g7/c5g7_2d"), "pop %d; inactive 3; active 400; seed 11;" % pop, device=0)
pp.generateInitialState(); pp.cycles(False, 3)
L, eng = pp.L, pp.engine
L.sb_profile_enable(eng, 1)
tot_max = 0; n = 0
try:
    for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 200):
        res = pp.cycle(True); tot_max += res.max_history_segments; n += 1
except Exception as e:
    print("stopped after", n, "cycles:", e)
ms = C.c_double(); a = C.c_int64(); b = C.c_int64(); c = C.c_int64()
L.sb_profile_read(eng, C.byref(ms), C.byref(a), C.byref(b), C.byref(c))
print("pop %d, %d cycles: kernel %.4f ms/cycle, longest history %.1f flights on average, %.2f us = %.0f cycles (1.965 GHz) per flight of the longest" % (
    pop, n, ms.value / max(n, 1), tot_max / max(n, 1), 1e3 * ms.value / max(1, tot_max), 1965 * ms.value / max(1, tot_max)))
