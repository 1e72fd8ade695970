"""Small driver for ncu captures of the CE lookup kernel (not a benchmark)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from scone_b200.ce import CeDatabase  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nuclides, material = bench.ce_nuclides_and_material(20)
eng = CeDatabase(nuclides, [material], device=0)
rng = np.random.default_rng(3)
E = torch.from_numpy(np.exp(rng.uniform(np.log(1e-11), np.log(20.0), n))).cuda()
mat = torch.ones(n, dtype=torch.int32, device="cuda")
tot = torch.zeros(n, dtype=torch.float64, device="cuda")
for _ in range(3):
    ms = eng.lookup_device(E.data_ptr(), mat.data_ptr(), tot.data_ptr(), 0, 0, n=n)
print("k_ce_lookup", n, "lookups", ms, "ms")
