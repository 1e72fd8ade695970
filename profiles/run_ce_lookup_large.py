"""The CE lookup kernel on a library that does not fit in L2: ~300 nuclides of 2e4 - 6e4 grid points each (the reference's five
bundled nuclides refined and cloned with seeded energy shifts), materials of 20 nuclides, random unsorted energies.
usage: run_ce_lookup_large.py [n_lookups] [n_nuclides] [sorted]"""
import os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
from scone_b200.ce import CeDatabase, large_library

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nn = int(sys.argv[2]) if len(sys.argv) > 2 else 300
nuclides, materials = large_library(nn)
t0 = time.time()
db = CeDatabase(nuclides, materials, device=0)
raw, idx, tab = db.memory()
print("library: %d nuclides, %.0f MB of nuclide tables, %.0f MB of lookup structures (%.2fx), union table %s, built in %.1f s" % (nn, raw / 1e6, idx / 1e6, (raw + idx) / raw, tab, time.time() - t0))
rng = np.random.default_rng(3)
E = torch.from_numpy(np.exp(rng.uniform(np.log(1e-11), np.log(19.0), n))).cuda()
mat = torch.from_numpy(rng.integers(1, len(materials) + 1, n).astype(np.int32)).cuda()
mode = sys.argv[3] if len(sys.argv) > 3 else ""
if mode == "sorted":                     # pre-sorted by the caller (energy order)
    E, order = torch.sort(E); mat = mat[order]
tot = torch.zeros(n, dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ms = []
for it in range(6):
    flush.zero_(); torch.cuda.synchronize()
    ms.append(db.lookup_device(E.data_ptr(), mat.data_ptr(), tot.data_ptr(), 0, 0, n=n, sort=(mode == "engine-sort")))
k = sum(ms[2:]) / len(ms[2:])
alg = (36 * 20 + 20) * n
print("%d lookups: %.3f ms, %.3g lookups/s, algorithmic %.0f GB/s" % (n, k, n / (k * 1e-3), alg / (k * 1e-3) / 1e9))
