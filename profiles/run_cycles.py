"""Small driver for ncu captures: a few cycles of a deck (not a benchmark; numbers under a profiler are not bench values)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scone_b200  # noqa: E402

deck = sys.argv[1] if len(sys.argv) > 1 else "decks/c5g7/c5g7_2d"
pop = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
extra = sys.argv[3] if len(sys.argv) > 3 else ""
pp = scone_b200.EigenPhysicsPackage(os.path.join(ROOT, deck), "pop %d; inactive 3; active 5; seed 1; %s" % (pop, extra), device=0)
pp.generateInitialState()
pp.cycles(False, 3)
res = pp.cycles(True, 3)
print("k", pp.k, "segments/cycle", res.n_segments, "launches", pp.launch_count())
