#!/usr/bin/env python3
"""Generates tests/golden/ce_nuclides.npz from the reference's bundled ACE files (run where /root/reference exists).

For each of the five nuclides that ship with the reference (IntegrationTestFiles/*.ace) the oracle's ACE loader
(oracle/cedata.hpp, pinned on the reference's regression values) produces eGrid and mainData exactly as
aceNeutronNuclide%init would; those arrays, plus the oracle's answers for a fixed set of energies, are the fixture
the GPU box uses (it has no /root/reference)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import oracle_lib as ol  # noqa: E402

REF = "/root/reference/IntegrationTestFiles/"
FILES = [("1001", "1001JEF311.ace", 1779), ("92233", "92233JEF311.ace", 1), ("52126", "52126JEF311.ace", 1),
         ("91231", "91231JEF311.ace", 1), ("91232", "91232JEF311.ace", 1)]


def main():
    orc = ol.load()
    out = {}
    probeE = np.array([1.0e-11, 1.1e-6, 5.6e-3, 3.6e-1, 1.6, 6.525, 17.0, 19.9, 20.0])
    for name, f, line in FILES:
        h = orc.orc_ce_nuclide_from_ace((REF + f).encode(), line)
        assert h, ol.err(orc)
        n, rows, m, kT = C.c_int(), C.c_int(), C.c_double(), C.c_double()
        orc.orc_ce_nuclide_info(h, C.byref(n), C.byref(rows), C.byref(m), C.byref(kT))
        g = np.zeros(n.value); d = np.zeros(n.value * rows.value)
        orc.orc_ce_nuclide_data(h, ol.dp(g), ol.dp(d))
        out["grid_" + name] = g
        out["data_" + name] = d.reshape(n.value, rows.value)
        mic = np.zeros((len(probeE), 8))
        for i, E in enumerate(probeE):
            assert orc.orc_ce_nuclide_micro(h, float(E), ol.dp(mic[i])) == 0
        out["micro_" + name] = mic
        out["awr_kT_" + name] = np.array([m.value, kT.value])
        orc.orc_ce_nuclide_free(h)
    out["probeE"] = probeE
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ce_nuclides.npz"), **out)
    print("wrote tests/golden/ce_nuclides.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
