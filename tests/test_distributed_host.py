"""Host logic of the multi-rank path on CPU: work shares, the loadBalancing plan and the neighbour exchange over a
world_size-2/3 gloo group (no GPU: banks are numpy arrays, the exchange uses the same TorchComm the GPU path uses)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import scone_b200
from scone_b200 import distributed as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_workshare_matches_mpi_func():
    # getWorkshare = (N + rank) / worldSize ; getOffset = N/ws*rank + max(0, mod(N, ws) + rank - ws)   (mpi_func.f90:133-159)
    for N in (10, 100000, 100003, 7):
        for ws in (1, 2, 3, 4, 8):
            shares = [D.workshare(N, ws, r) for r in range(ws)]
            assert sum(s for s, _ in shares) == N
            off = 0
            for r, (s, o) in enumerate(shares):
                assert s == (N + r) // ws
                assert o == off
                off += s
    assert [D.workshare(10, 4, r) for r in range(4)] == [(2, 0), (2, 2), (3, 4), (3, 7)]


def test_unbalanceable_distribution_is_seen_by_every_rank():
    """sizes (0, 0, 12) over three ranks: rank 1 would have to pass on sites it has not received yet; nearest-neighbour transfers
    cannot do that. load_balance works out the plan of every rank on every rank, so all of them refuse before anyone waits."""
    import pytest
    from scone_b200.lib import EngineError
    sizes = [0, 0, 12]
    failing = []
    for r in range(3):
        try:
            D.balance_plan(12, 3, r, sizes)
        except EngineError:
            failing.append(r)
    assert failing == [2] or failing                         # at least one rank's plan is impossible ...
    for me in range(3):                                      # ... and the loop every rank runs in load_balance finds it whoever `me` is
        with pytest.raises(EngineError):
            for r in range(3):
                D.balance_plan(12, 3, r, sizes)


def test_balance_plan_restores_target_offsets():
    rng = np.random.default_rng(5)
    for ws in (2, 3, 4, 8):
        for _ in range(50):
            tot = int(rng.integers(1000, 5000))
            target = np.array([D.workshare(tot, ws, r)[0] for r in range(ws)])
            # perturb sizes a little (what normSize_Repr leaves behind), keeping the total
            d = rng.integers(-20, 21, ws); d[-1] -= d.sum()
            sizes = target + d
            plans = [D.balance_plan(tot, ws, r, sizes) for r in range(ws)]
            for r in range(ws):
                su, ru, sd, rd = plans[r]
                assert su == 0 or ru == 0
                assert sd == 0 or rd == 0
                if r + 1 < ws:
                    assert su == plans[r + 1][3] and ru == plans[r + 1][2]      # my up == neighbour's down
                else:
                    assert su == 0 and ru == 0
                if r == 0:
                    assert sd == 0 and rd == 0
                assert sizes[r] - su - sd + ru + rd == target[r]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from scone_b200 import distributed as D
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=int(sys.argv[4]))
rank, ws = dist.get_rank(), dist.get_world_size()
comm = D.TorchComm(device=torch.device("cpu"))
tot = 1000
rng = np.random.default_rng(11)
target = np.array([D.workshare(tot, ws, r)[0] for r in range(ws)])
d = rng.integers(-15, 16, ws); d[-1] -= d.sum()
sizes = target + d
start = int(sizes[:rank].sum())
bank = np.arange(start, start + sizes[rank], dtype=np.int64)          # global site numbers stand in for sites
# the three exchanges
comm.sums.copy_(torch.arange(8, dtype=torch.float64) * (rank + 1)); comm.all_reduce_sums()
assert torch.allclose(comm.sums, torch.arange(8, dtype=torch.float64) * sum(range(1, ws + 1)))
comm.sums.copy_(torch.arange(8, dtype=torch.float64) + 10 * rank)
g = comm.all_gather_sums()
assert g.shape == (ws, 8) and all(np.array_equal(g[r], np.arange(8) + 10.0 * r) for r in range(ws))
got = comm.all_gather_int(int(sizes[rank])); assert got == [int(x) for x in sizes]
su, ru, sd, rd = D.balance_plan(tot, ws, rank, sizes)
as_u8 = lambda a: torch.from_numpy(a.view(np.uint8).copy())
bu, bd = torch.empty(8 * ru, dtype=torch.uint8), torch.empty(8 * rd, dtype=torch.uint8)
sends, recvs = [], []
if su: sends.append((rank + 1, as_u8(bank[len(bank) - su:])))
if sd: sends.append((rank - 1, as_u8(bank[:sd])))
if ru: recvs.append((rank + 1, bu))
if rd: recvs.append((rank - 1, bd))
comm.exchange(sends, recvs)
new = np.concatenate([bd.numpy().view(np.int64), bank[sd:len(bank) - su], bu.numpy().view(np.int64)])
share, off = D.workshare(tot, ws, rank)
assert len(new) == share and np.array_equal(new, np.arange(off, off + share)), (rank, new[:5], off, share)
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


@pytest.mark.parametrize("ws", [2, 3])
def test_neighbour_exchange_over_gloo(ws, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(_free_port())
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), str(ws)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(ws)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and ("ok %d" % r) in o, o
