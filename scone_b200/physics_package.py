"""Python face of the host driver: mirrors SCONE's eigenPhysicsPackage (init / generateInitialState /
cycles / run; PhysicsPackages/eigenPhysicsPackage_class.f90:135-366) on top of the C ABI."""
import ctypes as C

import numpy as np

from .lib import CycleResult, EngineError, load_library


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class _Base:
    def _err(self):
        return self.L.sbh_last_error(self.h).decode()

    @property
    def engine(self):
        return self.L.sbh_engine(self.h)

    def _eng_err(self):
        return self.L.sb_last_error(self.engine).decode()

    def geom_query(self, r, u, dist=None):
        """placeCoord (dist None) or teleport (dist given) for n points on the device; returns mat, uniqueID, r, u."""
        r = np.ascontiguousarray(r, np.float64).copy(); u = np.ascontiguousarray(u, np.float64).copy()
        n = r.shape[0]
        mat = np.zeros(n, np.int32); uid = np.zeros(n, np.int32)
        d = None if dist is None else np.ascontiguousarray(dist, np.float64)
        if self.L.sb_geom_query(self.engine, n, _dp(r), _dp(u), None if d is None else _dp(d), _ip(mat), _ip(uid)) != 0:
            raise EngineError(self._eng_err())
        return mat, uid, r, u

    def launch_count(self):
        return int(self.L.sb_launch_count(self.engine))


class GeometryHandle(_Base):
    """Flattened geometry (csg -> device graph) without nuclear data, for geometry queries."""

    def __init__(self, text, is_path=False, device=0):
        self.L = load_library()
        self.h = self.L.sbh_geom_create(text.encode(), 1 if is_path else 0, device)
        if not self.h:
            raise EngineError(self.L.sbh_last_error(None).decode())

    def info(self):
        v = [C.c_int32() for _ in range(8)]
        self.L.sbh_geom_info(self.h, *[C.byref(x) for x in v])
        keys = ["nSurf", "nCell", "nUni", "nGraph", "uniqueCells", "rootIdx", "borderIdx", "nesting"]
        return dict(zip(keys, [x.value for x in v]))

    def graph(self):
        n = self.info()["nGraph"]
        idx = np.zeros(n, np.int32); gid = np.zeros(n, np.int32)
        self.L.sbh_model_graph(self.h, _ip(idx), _ip(gid))
        return idx, gid

    def uni_fill(self, uni_idx):
        out = np.zeros(1 << 16, np.int32)
        n = self.L.sbh_geom_uni_fill(self.h, uni_idx, _ip(out), len(out))
        return out[:n].tolist()

    def active_mats(self):
        out = np.zeros(4096, np.int32)
        n = self.L.sbh_geom_active_mats(self.h, _ip(out), len(out))
        return out[:n].tolist()

    def close(self):
        if self.h:
            self.L.sbh_eigen_destroy(self.h)
            self.h = None


class FixedSourcePhysicsPackage:
    """fixedSourcePhysicsPackage on the B200 engine: the same handle as EigenPhysicsPackage, created from a deck of that type."""

    def __new__(cls, deck, overrides="", device=0, rank=0, n_ranks=1):
        pp = EigenPhysicsPackage(deck, overrides, device, rank=rank, n_ranks=n_ranks)
        if not pp.is_fixed_source:
            pp.close()
            raise EngineError("%s is not a fixedSourcePhysicsPackage deck" % deck)
        return pp


class EigenPhysicsPackage(_Base):
    """eigenPhysicsPackage on the B200 engine.

    deck      -- path of a SCONE input file (type eigenPhysicsPackage; dataType mg)
    overrides -- dictionary text whose top-level entries replace the deck's (pop, active, seed, ...)
    rank/n_ranks -- bank share of this process (mpi_func.f90 getWorkshare/getOffset)
    """

    def __init__(self, deck, overrides="", device=0, rank=0, n_ranks=1):
        self.L = load_library()
        self.h = self.L.sbh_eigen_create(str(deck).encode(), overrides.encode(), device, rank, n_ranks)
        if not self.h:
            raise EngineError(self.L.sbh_last_error(None).decode())
        v = [C.c_int32() for _ in range(7)]
        self.L.sbh_eigen_info(self.h, *[C.byref(x) for x in v])
        self.pop, self.n_inactive, self.n_active, self.n_groups, self.n_mat, self.n_graph, self.unique_cells = [x.value for x in v]
        self.total_pop = int(self.L.sbh_eigen_total_pop(self.h))
        self.rank, self.n_ranks = rank, n_ranks
        self.k = self.L.sbh_eigen_keff0(self.h)

    # -- reference-named procedures ----------------------------------------------------------
    def generateInitialState(self):
        if self.L.sbh_eigen_generate_initial_state(self.h) != 0:
            raise EngineError(self._err())

    def cycle(self, active, host_buffers=False, comm=None):
        """One cycle; returns the CycleResult. self.k is k_new afterwards.
        comm: a scone_b200.distributed.TorchComm when the bank is shared between ranks."""
        if comm is not None:
            from . import distributed
            if host_buffers:                    # dungeons kept in (pinned) host memory: upload, cycle, download
                if self.L.sbh_eigen_upload_bank(self.h) != 0:
                    raise EngineError(self._err())
            res = distributed.cycle(self, active, comm)
            if host_buffers:
                if self.L.sbh_eigen_download_bank(self.h) != 0:
                    raise EngineError(self._err())
                self.last_bins(active)
            return res
        if self.n_ranks > 1:
            raise EngineError("this package owns a share of the bank (n_ranks > 1): pass comm= to cycle()")
        k = C.c_double(self.k)
        res = CycleResult()
        fn = self.L.sbh_eigen_cycle_host_buffers if host_buffers else self.L.sbh_eigen_cycle
        if fn(self.h, 1 if active else 0, C.byref(k), C.byref(res)) != 0:
            raise EngineError(self._err())
        self.k = k.value
        return res

    def cycles(self, active, n, comm=None):
        out = None
        for _ in range(n):
            out = self.cycle(active, comm=comm)
        return out

    def run(self):
        self.generateInitialState()
        self.cycles(False, self.n_inactive)
        return self.cycles(True, self.n_active)

    # -- fixedSourcePhysicsPackage (decks of that type; n_active = cycles) -------------------------------
    @property
    def is_fixed_source(self):
        return bool(self.L.sbh_eigen_is_fixed(self.h))

    def fixed_cycle(self):
        """One source batch: pointSource, histories with their secondaries, tallies closed. Returns the CycleResult."""
        res = CycleResult()
        if self.L.sbh_fixed_cycle(self.h, C.byref(res)) != 0:
            raise EngineError(self._err())
        return res

    # -- data access ------------------------------------------------------------------------------
    @property
    def rng_state(self):
        return int(self.L.sbh_eigen_rng_state(self.h))

    @rng_state.setter
    def rng_state(self, s):
        self.L.sbh_eigen_set_rng_state(self.h, s)

    @property
    def is_ce(self):
        return bool(self.L.sbh_eigen_is_ce(self.h))

    def bank(self):
        """The current bank: (r, dir, w, G) for multigroup decks, (r, dir, w, E) for continuous-energy decks."""
        cap = 2 * self.pop + 1024
        n = C.c_int32()
        r = np.zeros((cap, 3)); d = np.zeros((cap, 3)); w = np.zeros(cap)
        if self.is_ce:
            E = np.zeros(cap)
            if self.L.sb_bank_download_ce(self.engine, cap, C.byref(n), _dp(r), _dp(d), _dp(w), _dp(E)) != 0:
                raise EngineError(self._eng_err())
            m = n.value
            return r[:m], d[:m], w[:m], E[:m]
        G = np.zeros(cap, np.int32)
        if self.L.sb_bank_download(self.engine, cap, C.byref(n), _dp(r), _dp(d), _dp(w), _ip(G)) != 0:
            raise EngineError(self._eng_err())
        m = n.value
        return r[:m], d[:m], w[:m], G[:m]

    def set_bank(self, r, d, w, G):
        r = np.ascontiguousarray(r, np.float64); d = np.ascontiguousarray(d, np.float64)
        w = np.ascontiguousarray(w, np.float64); G = np.ascontiguousarray(G, np.int32)
        if self.L.sb_bank_upload(self.engine, len(w), _dp(r), _dp(d), _dp(w), _ip(G)) != 0:
            raise EngineError(self._eng_err())

    def tally(self, active=True):
        ph = 1 if active else 0
        n = int(self.L.sb_tally_size(self.engine, ph))
        cs = np.zeros(max(1, n)); cs2 = np.zeros(max(1, n)); b = C.c_int32()
        if self.L.sb_tally_read(self.engine, ph, _dp(cs), _dp(cs2), C.byref(b)) != 0:
            raise EngineError(self._eng_err())
        return cs[:n], cs2[:n], b.value

    def last_bins(self, active=True):
        ph = 1 if active else 0
        n = int(self.L.sb_tally_size(self.engine, ph))
        out = np.zeros(max(1, n))
        if self.L.sb_tally_last_bins(self.engine, ph, _dp(out)) != 0:
            raise EngineError(self._eng_err())
        return out[:n]

    def stats(self):
        a, b, c, t = C.c_longlong(), C.c_longlong(), C.c_longlong(), C.c_double()
        self.L.sbh_eigen_stats(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(t))
        return dict(seg_inactive=a.value, seg_active=b.value, histories=c.value, t_transport=t.value)

    def host_bytes(self, active=True):
        a, b = C.c_longlong(), C.c_longlong()
        self.L.sbh_eigen_host_bytes(self.h, 1 if active else 0, C.byref(a), C.byref(b))
        return a.value, b.value

    def model_graph(self):
        idx = np.zeros(self.n_graph, np.int32); gid = np.zeros(self.n_graph, np.int32)
        self.L.sbh_model_graph(self.h, _ip(idx), _ip(gid))
        return idx, gid

    def model_xs(self):
        data = np.zeros((self.n_mat, self.n_groups, 6)); maj = np.zeros(self.n_groups)
        self.L.sbh_model_xs(self.h, _dp(data), _dp(maj))
        return data, maj

    def mg_query(self, mat, G):
        mat = np.ascontiguousarray(mat, np.int32); G = np.ascontiguousarray(G, np.int32)
        tot = np.zeros(len(mat)); maj = np.zeros(len(mat))
        if self.L.sb_mg_query(self.engine, len(mat), _ip(mat), _ip(G), _dp(tot), _dp(maj)) != 0:
            raise EngineError(self._eng_err())
        return tot, maj

    def close(self):
        if self.h:
            self.L.sbh_eigen_destroy(self.h)
            self.h = None
