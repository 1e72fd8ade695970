"""Worker of tests/test_gpu_distributed.py: one rank of a fixed-source run (argv: root port rank world deck overrides outdir)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

root, port, rank, ws, deck, ov, out = sys.argv[1:8]
rank, ws = int(rank), int(ws)
sys.path.insert(0, root)
import scone_b200  # noqa: E402
from scone_b200 import distributed as D  # noqa: E402

torch.cuda.set_device(0)
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % port, rank=rank, world_size=ws)
comm = D.TorchComm(device=torch.device("cuda", 0))
pp = scone_b200.FixedSourcePhysicsPackage(deck, ov, device=0, rank=rank, n_ranks=ws)
segs = 0
for _ in range(pp.n_active):
    segs += pp.fixed_cycle().n_segments
cs, cs2, nb = pp.tally(True)
ccs, ccs2, cnb = D.collect_distributed(pp, comm, True)
np.savez(os.path.join(out, "fixed_r%d.npz" % rank), cs=cs, cs2=cs2, nb=nb, ccs=ccs, ccs2=ccs2, cnb=cnb, seg=segs, pop=pp.pop)
pp.close()
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
