"""Several ranks, one engine (GPU) each: the per-cycle exchanges SCONE does over MPI, on torch.distributed.

  particleDungeon%normSize_Repr / loadBalancing   ParticleObjects/particleDungeon_class.f90:431-698
  scoreMemory%reduceBins (mpiSync clerks)          Tallies/scoreMemory_class.f90:404-431
  k_eff broadcast                                  PhysicsPackages/eigenPhysicsPackage_class.f90:302
  getWorkshare / getOffset                         SharedModules/mpi_func.f90:133-159

Per cycle: ONE all-gather of 8 doubles per rank (the 6 k-eff score sums and the fission-bank size) and the send/recv of
the few sites that cross the boundaries between neighbouring ranks. The resampling threshold and every rank's bank
size after normalisation are recomputed on every rank (pure functions of the sizes and the master RNG state): no
gather + 3 broadcasts + 2 all-gathers as in the MPI code.
With the nccl backend the buffers are device memory and move over NVLink; with gloo (CPU tests, or several
ranks sharing one GPU) they are staged through host memory.
"""
import ctypes as C

import numpy as np

from .lib import CycleResult, EngineError, load_library


def workshare(tot_pop, n_ranks, rank):
    """(share, offset) of `rank` -- mpi_func.f90:133-159."""
    L = load_library()
    s, o = C.c_int32(), C.c_int32()
    L.sbh_workshare(tot_pop, n_ranks, rank, C.byref(s), C.byref(o))
    return s.value, o.value


def balance_plan(tot_pop, n_ranks, rank, pop_sizes):
    """loadBalancing counts of `rank`: (send_up, recv_up, send_down, recv_down).
    up = rank + 1 (sites leave from / arrive at the END of the bank), down = rank - 1 (the BEGINNING)."""
    L = load_library()
    sizes = np.ascontiguousarray(pop_sizes, np.int32)
    out = np.zeros(4, np.int32)
    rc = L.sbh_balance_plan(tot_pop, n_ranks, rank, sizes.ctypes.data_as(C.POINTER(C.c_int32)), out.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        raise EngineError("loadBalancing: rank %d would have to send more sites than it holds (sizes %s)" % (rank, list(pop_sizes)))
    return tuple(int(x) for x in out)


SITE_BYTES = 8 * 8 + 8      # r, dir, w, E (f64) + G, broodID (i32), as sb_site_buffer_bytes


def site_buffer_bytes(k):
    return k * SITE_BYTES + 8


class TorchComm:
    """The three exchanges on a torch.distributed process group."""

    def __init__(self, group=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        self.device = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))
        self.on_device = self.backend == "nccl"
        self.sums = torch.zeros(8, dtype=torch.float64, device=self.device)      # filled by the engine: 6 score sums, bank size, 0
        self._stage = torch.zeros(8, dtype=torch.float64)
        self._gath = torch.zeros(8 * self.size, dtype=torch.float64, device=self.device if self.on_device else "cpu")
        self._sizes = torch.zeros(self.size, dtype=torch.int32, device=self.device if self.on_device else "cpu")
        self._one = torch.zeros(1, dtype=torch.int32, device=self.device if self.on_device else "cpu")
        self._bufs = {}
        self.peer = False          # set by enable_peer(): the engines exchange through peer memory, this object is not used per cycle

    def sync(self):
        if self.device.type == "cuda":
            self.torch.cuda.synchronize(self.device)

    def all_reduce_sums(self):
        """In-place sum over ranks of self.sums (device)."""
        if self.on_device:
            self.dist.all_reduce(self.sums, group=self.group)
            self.sync()
        else:
            self._stage.copy_(self.sums)
            self.dist.all_reduce(self._stage, group=self.group)
            self.sums.copy_(self._stage)
            self.sync()

    def all_gather_sums(self):
        """All-gather of self.sums (8 doubles per rank) -> numpy [size][8] on the host."""
        if self.on_device:
            self.dist.all_gather_into_tensor(self._gath, self.sums, group=self.group)
            return self._gath.cpu().numpy().reshape(self.size, 8).copy()
        self._stage.copy_(self.sums)
        self.dist.all_gather(list(self._gath.split(8)), self._stage, group=self.group)
        return self._gath.numpy().reshape(self.size, 8).copy()

    def all_gather_int(self, v):
        self._one[0] = int(v)
        self.dist.all_gather_into_tensor(self._sizes, self._one, group=self.group) if self.on_device else \
            self.dist.all_gather(list(self._sizes.split(1)), self._one, group=self.group)
        return [int(x) for x in self._sizes.cpu().tolist()]

    def buffer(self, key, nbytes):
        b = self._bufs.get(key)
        if b is None or b.numel() < nbytes:
            b = self.torch.empty(max(nbytes, 1 << 16), dtype=self.torch.uint8, device=self.device)
            self._bufs[key] = b
        return b

    def exchange(self, sends, recvs):
        """sends / recvs: lists of (peer, device uint8 tensor view of exact length). Nearest neighbours only."""
        dist = self.dist
        if self.on_device:
            ops = [dist.P2POp(dist.isend, t, p, group=self.group) for p, t in sends] + [dist.P2POp(dist.irecv, t, p, group=self.group) for p, t in recvs]
            if ops:
                for r in dist.batch_isend_irecv(ops):
                    r.wait()
                self.sync()
        else:
            reqs, staged = [], []
            for p, t in sends:
                reqs.append(dist.isend(t.cpu(), p, group=self.group))
            for p, t in recvs:
                c = self.torch.empty(t.numel(), dtype=self.torch.uint8)
                staged.append((t, c))
                reqs.append(dist.irecv(c, p, group=self.group))
            for r in reqs:
                r.wait()
            for t, c in staged:
                t.copy_(c)
            self.sync()


def enable_peer(pp, comm):
    """Connect the engines of the ranks of one node through peer memory (include/scone_b200.h, sb_peer_*): every rank exports its
    mailbox region with CUDA IPC, the handles go round once over the process group; afterwards a cycle needs no collective.
    Returns True if every rank attached (False, with the NCCL / gloo exchange kept, if IPC is not available)."""
    import torch
    L, eng = pp.L, pp.engine
    handle = C.create_string_buffer(64)
    caps = [None] * comm.size
    comm.dist.all_gather_object(caps, int(L.sb_bank_capacity(eng)), group=comm.group)
    ok = min(caps) > 0 and L.sb_peer_create(eng, comm.size, comm.rank, max(caps), handle) == 0
    why = "" if ok else pp._eng_err()
    mine = (bytes(handle.raw), int(L.sb_peer_capacity(eng)) if ok else -1, socket_host())
    everyone = [None] * comm.size
    comm.dist.all_gather_object(everyone, mine, group=comm.group)
    ok = ok and all(c == mine[1] and c > 0 for _, c, _ in everyone) and all(hn == mine[2] for _, _, hn in everyone)
    if ok:
        blob = b"".join(hd for hd, _, _ in everyone)
        caps = np.array([c for _, c, _ in everyone], np.int32)
        ok = L.sb_peer_attach(eng, blob, caps.ctypes.data_as(C.POINTER(C.c_int32))) == 0
        why = "" if ok else pp._eng_err()
    elif not why:
        why = "ranks on different hosts or with different bank capacities: %r" % [(c, hn) for _, c, hn in everyone]
    flags = [None] * comm.size
    comm.dist.all_gather_object(flags, bool(ok), group=comm.group)
    comm.peer = all(flags)
    comm.peer_error = why
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    return comm.peer


def socket_host():
    import socket
    return socket.gethostname()


def cycle_peer(pp, active, comm):
    """One cycle with the exchange done by the engines themselves over peer memory: one call, one synchronisation."""
    L = pp.L
    final = np.zeros(comm.size, np.int32)
    res, k = CycleResult(), C.c_double(pp.k)
    if L.sbh_eigen_cycle_peer(pp.h, 1 if active else 0, C.byref(k), final.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(res)) != 0:
        raise EngineError(pp._err())
    pp.k = k.value
    if int(final.sum()) != pp.total_pop:
        raise EngineError("Normalisation failed!")
    return res


def cycle(pp, active, comm):
    """One cycle of `pp` (EigenPhysicsPackage created with rank / n_ranks) in step with the other ranks.
    eigenPhysicsPackage_class.f90:203-307 with MPI defined."""
    L = pp.L
    if getattr(comm, "peer", False):
        return cycle_peer(pp, active, comm)
    n_sites = C.c_int32()
    if L.sbh_eigen_cycle_begin(pp.h, 1 if active else 0, pp.k, comm.sums.data_ptr(), C.byref(n_sites)) != 0:
        raise EngineError(pp._err())
    # ONE collective per cycle: every rank gets every rank's 6 score sums and bank size, and adds the sums in rank order
    # (same bits on every rank, independent of the collective's internal order)
    g = comm.all_gather_sums()
    sums = np.zeros(6)
    for r in range(comm.size):
        sums = sums + g[r, :6]
    sizes = np.ascontiguousarray(g[:, 6].astype(np.int32))
    new_sizes = np.zeros(comm.size, np.int32)
    res, k = CycleResult(), C.c_double(pp.k)
    if L.sbh_eigen_cycle_end_resample_ranked(pp.h, 1 if active else 0, sums.ctypes.data_as(C.POINTER(C.c_double)),
                                             sizes.ctypes.data_as(C.POINTER(C.c_int32)), new_sizes.ctypes.data_as(C.POINTER(C.c_int32)),
                                             C.byref(k), C.byref(res)) != 0:
        raise EngineError(pp._err())
    pp.k = k.value
    sizes2 = [int(x) for x in new_sizes]
    if sum(sizes2) != pp.total_pop:
        raise EngineError("Normalisation failed!")
    if comm.size > 1:
        load_balance(pp, comm, sizes2)
        # printSource: nextCycle%printToFile comes after normSize_Repr, whose last step is the balancing
        if L.sbh_eigen_print_source(pp.h, 1 if active else 0) != 0:
            raise EngineError(pp._err())
    return res


def load_balance(pp, comm, sizes):
    """particleDungeon%loadBalancing on the device banks."""
    L, eng = pp.L, pp.engine
    # every rank works out the plan of EVERY rank from the same sizes, so that a distribution nearest-neighbour transfers cannot
    # fix (a rank would have to pass on sites it has not received yet) is refused by all ranks together, before anyone waits in
    # a receive; the peer-memory path does the same on the device (k_peer_plan, SB_ERR_BALANCE)
    for r in range(comm.size):
        if r != comm.rank:
            balance_plan(pp.total_pop, comm.size, r, sizes)
    send_up, recv_up, send_down, recv_down = balance_plan(pp.total_pop, comm.size, comm.rank, sizes)
    if not (send_up or recv_up or send_down or recv_down):
        return
    view = lambda key, k: comm.buffer(key, site_buffer_bytes(k))[:site_buffer_bytes(k)]
    su, sd, ru, rd = view("su", send_up), view("sd", send_down), view("ru", recv_up), view("rd", recv_down)
    if L.sb_bank_export(eng, send_down, sd.data_ptr(), send_up, su.data_ptr()) != 0:
        raise EngineError(pp._eng_err())
    sends, recvs = [], []
    if send_up:
        sends.append((comm.rank + 1, su))
    if send_down:
        sends.append((comm.rank - 1, sd))
    if recv_up:
        recvs.append((comm.rank + 1, ru))
    if recv_down:
        recvs.append((comm.rank - 1, rd))
    comm.exchange(sends, recvs)
    if L.sb_bank_splice(eng, send_down, send_up, recv_down, rd.data_ptr(), recv_up, ru.data_ptr()) != 0:
        raise EngineError(pp._eng_err())


def collect_distributed(pp, comm, active=True):
    """scoreMemory%collectDistributed (Tallies/scoreMemory_class.f90:443-467) at the end of a run whose ranks tallied independently:
    the cumulative sums, the sums of squares and the batch counts are summed over ranks; the master (rank 0) holds the result and the
    other ranks are left with a batch count of 0, as in the reference. Returns (csum, csum2, batch_n) as numpy arrays / int."""
    torch = comm.torch
    cs, cs2, nb = pp.tally(active)
    dev = comm.device if comm.on_device else torch.device("cpu")
    buf = torch.from_numpy(np.concatenate([cs, cs2, [float(nb)]])).to(dev)
    comm.dist.all_reduce(buf, group=comm.group)              # sum; reduce-to-master semantics are applied below
    out = buf.cpu().numpy()
    n = len(cs)
    if comm.rank == 0:
        return out[:n].copy(), out[n:2 * n].copy(), int(round(out[2 * n]))
    return cs, cs2, 0
