"""The engine driven through its C ABI alone, from C: tests/c_driver/run_cycle.c (strict C99, sb_* entry points only, no host
driver) loads the committed flat model of the C5G7 deck, runs eigenvalue cycles as eigenPhysicsPackage%cycles does, and must
report, cycle by cycle, the k-eff, site and segment counts of the same run through the C++ host driver."""
import os
import re
import subprocess

import pytest

import scone_b200
from tests.gpu_util import DECK

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_driver_runs_the_same_cycles(tmp_path):
    exe = str(tmp_path / "run_cycle")
    libdir = os.path.join(ROOT, "scone_b200")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_driver", "run_cycle.c"),
                           "-L", libdir, "-lscone_b200", "-Wl,-rpath," + libdir, "-o", exe])
    pop, ni, na = 5000, 2, 2
    out = subprocess.check_output([exe, os.path.join(ROOT, "tests", "golden", "c5g7_flat.bin"), str(pop), str(ni), str(na)], text=True)
    rows = [(float(m.group(1)), int(m.group(2)), int(m.group(3))) for m in re.finditer(r"k_cum (\S+) sites (\d+) segments (\d+)", out)]
    assert len(rows) == ni + na
    pp = scone_b200.EigenPhysicsPackage(DECK["c5g7"], "pop %d; inactive %d; active %d; seed 20261017;" % (pop, ni, na), device=0)
    pp.generateInitialState()
    for cyc, (k, sites, seg) in enumerate(rows):
        res = pp.cycle(cyc >= ni)
        assert (res.n_sites, res.n_segments) == (sites, seg)
        assert res.k_cum == k                     # %.17g round-trips a double
    pp.close()
