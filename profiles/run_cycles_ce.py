"""Small driver for ncu captures of the continuous-energy history kernel (not a benchmark)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scone_b200  # noqa: E402

pop = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
extra = sys.argv[2] if len(sys.argv) > 2 else ""
deck = sys.argv[3] if len(sys.argv) > 3 else "decks/ce/pincell"
pp = scone_b200.EigenPhysicsPackage(os.path.join(ROOT, deck), "pop %d; inactive 2; active 3; seed 1; %s" % (pop, extra), device=0)
pp.generateInitialState()
pp.cycles(False, 2)
res = pp.cycles(True, 3)
print("k", pp.k, "segments/cycle", res.n_segments, "collisions/cycle", res.n_collisions, "launches", pp.launch_count())
