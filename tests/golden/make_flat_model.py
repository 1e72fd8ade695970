#!/usr/bin/env python3
"""Writes tests/golden/c5g7_flat.bin: the flat model of decks/c5g7/c5g7_2d (the arrays that cross the C ABI) as the host-side
model builder produces it, for the plain-C driver test (tests/c_driver/run_cycle.c).  Needs no GPU: the package handle is
created without a device.  Seed and population of the run are the driver's arguments / the values below."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import scone_b200  # noqa: E402

OVERRIDES = "pop 5000; inactive 2; active 2; seed 20261017;"
if __name__ == "__main__":
    L = scone_b200.load_library()
    h = L.sbh_eigen_create(os.path.join(ROOT, "decks", "c5g7", "c5g7_2d").encode(), OVERRIDES.encode(), -1, 0, 1)
    assert h, L.sbh_last_error(None).decode()
    out = os.path.join(ROOT, "tests", "golden", "c5g7_flat.bin")
    assert L.sbh_model_dump(h, out.encode()) == 0, L.sbh_last_error(h).decode()
    L.sbh_eigen_destroy(h)
    print("wrote", out, os.path.getsize(out), "bytes")
