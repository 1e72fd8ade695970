"""One warp of histories alone on the GPU (ncu target): what a single warp's flight costs in instructions and stall cycles."""
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import scone_b200
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
pp = scone_b200.EigenPhysicsPackage(os.path.join(ROOT, "decks/c5g7/c5g7_2d"), "pop 32; inactive 3; active 4; seed 7;", device=0)
pp.generateInitialState()
pp.cycles(False, 3)
tot = 0
for _ in range(4):
    res = pp.cycle(True); tot += res.n_segments
print("segments per cycle", tot / 4)
