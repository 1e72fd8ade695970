"""Where the per-cycle time of a multi-rank run goes (torchrun, NCCL): stage timings of scone_b200.distributed.cycle (diagnostic)."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scone_b200  # noqa: E402
from scone_b200 import distributed as D  # noqa: E402
from scone_b200.lib import CycleResult  # noqa: E402

rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = D.TorchComm(device=torch.device("cuda", local))
pp = scone_b200.EigenPhysicsPackage(os.path.join(ROOT, "decks/c5g7/c5g7_2d"), "pop %d; inactive 5; active 100; seed 1;" % (100000 * ws), device=local, rank=rank, n_ranks=ws)
pp.generateInitialState()
for _ in range(8):
    pp.cycle(True, comm=comm)
L = pp.L
T = np.zeros(6)
N = 30
for _ in range(N):
    t = [time.perf_counter()]
    n_sites = C.c_int32()
    L.sbh_eigen_cycle_begin(pp.h, 1, pp.k, comm.sums.data_ptr(), C.byref(n_sites)); t.append(time.perf_counter())
    g = comm.all_gather_sums(); t.append(time.perf_counter())
    sums = np.zeros(6)
    for r in range(ws):
        sums = sums + g[r, :6]
    sizes = np.ascontiguousarray(g[:, 6].astype(np.int32)); new_sizes = np.zeros(ws, np.int32)
    res, k = CycleResult(), C.c_double(pp.k)
    L.sbh_eigen_cycle_end_resample_ranked(pp.h, 1, sums.ctypes.data_as(C.POINTER(C.c_double)), sizes.ctypes.data_as(C.POINTER(C.c_int32)),
                                          new_sizes.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(k), C.byref(res)); t.append(time.perf_counter())
    pp.k = k.value
    D.load_balance(pp, comm, [int(x) for x in new_sizes]); t.append(time.perf_counter())
    T[:4] += np.diff(t)
if rank == 0:
    names = ["cycle_begin (transport + sort + sums, 1 sync)", "all-gather of 8 doubles + copy to host", "cycle_end + resample (1 sync)", "load balance (export, send/recv, splice)"]
    for n, v in zip(names, T[:4] / N * 1e3):
        print("%-55s %.3f ms" % (n, v))
    print("total %.3f ms per cycle" % (T[:4].sum() / N * 1e3))
pp.close()
dist.barrier(); dist.destroy_process_group()
