"""Cycles of FOUR histories (one warp, soon one lane): ncu target for the latency of a history that runs alone."""
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import scone_b200
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
pp = scone_b200.EigenPhysicsPackage(os.path.join(ROOT, "decks/c5g7/c5g7_2d"), "pop 4; inactive 3; active 400; seed 11;", device=0)
pp.generateInitialState()
pp.cycles(False, 3)
seg = 0
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
    res = pp.cycle(True); seg += res.n_segments
print("segments", seg)
