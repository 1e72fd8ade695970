"""Worker of tests/test_gpu_distributed.py: one rank of an eigenvalue run whose bank is shared between ranks.
argv: root port rank world backend deck overrides ncycles_inactive ncycles_active outdir [peer]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

root, port, rank, ws, backend, deck, ov, ninact, nact, out = sys.argv[1:11]
rank, ws, ninact, nact = int(rank), int(ws), int(ninact), int(nact)
sys.path.insert(0, root)
import scone_b200  # noqa: E402
from scone_b200 import distributed as D  # noqa: E402

dev = rank if backend == "nccl" else 0
torch.cuda.set_device(dev)
dist.init_process_group(backend, init_method="tcp://127.0.0.1:%s" % port, rank=rank, world_size=ws)
comm = D.TorchComm(device=torch.device("cuda", dev))
pp = scone_b200.EigenPhysicsPackage(deck, ov, device=dev, rank=rank, n_ranks=ws)
if len(sys.argv) > 11 and sys.argv[11] == "peer":          # exchange through peer memory instead of the process group
    assert D.enable_peer(pp, comm), "peer memory could not be attached: " + comm.peer_error
if len(sys.argv) > 11 and sys.argv[11] == "peer_absent":    # rank 1 never runs its cycle: rank 0 must come back with the time-out error
    assert D.enable_peer(pp, comm), "peer memory could not be attached: " + comm.peer_error
    pp.generateInitialState()
    if rank == 0:
        pp.L.sb_peer_set_timeout(pp.engine, 0.5)
        try:
            pp.cycle(False, comm=comm)
            print("no error raised")
        except scone_b200.EngineError as ex:
            print("raised:", ex)
            if "did not post its cycle data in time" in str(ex):
                print("ok", rank)
    else:
        print("ok", rank)
    dist.barrier()
    pp.close()
    dist.destroy_process_group()
    sys.exit(0)
pp.generateInitialState()
ks, segs = [], []
for c in range(ninact + nact):
    res = pp.cycle(c >= ninact, comm=comm)
    r, d, w, G = pp.bank()
    np.savez(os.path.join(out, "bank_c%d_r%d.npz" % (c, rank)), r=r, d=d, w=w, G=G)
    ks.append(pp.k); segs.append(res.n_segments)
cs, cs2, nb = pp.tally(True)
ccs, ccs2, cnb = D.collect_distributed(pp, comm, True)          # tally%collectDistributed: the master holds the sums over ranks
np.savez(os.path.join(out, "collected_r%d.npz" % rank), cs=ccs, cs2=ccs2, nb=cnb)
np.savez(os.path.join(out, "final_r%d.npz" % rank), k=np.array(ks), seg=np.array(segs), cs=cs, cs2=cs2, nb=nb, rng=np.array([pp.rng_state], np.uint64))
pp.close()
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
