#!/usr/bin/env python3
"""bench.py -- active-cycle throughput of the SCONE hot path on B200 (and of the CPU reference arm).

Metric (BASELINE.json): active-cycle neutrons/s (and segments/s) on the C5G7 MOX 2-D 7-group eigenvalue
problem, delta tracking, 1e5 neutrons per cycle per GPU.  A "step" is one active cycle: every history of the
current fission bank is transported to its death (XS lookup, delta-tracking flights, collisions, tallies,
fission-site banking), the cycle's tallies are closed and the next bank is normalised (normSize_Repr).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--pop P] [--deck c5g7|c5g7_3d|inf|slab]

value : whole-job neutrons/s with the bank resident in HBM (CUDA-event time, max over ranks)
e2e   : the same through the host-buffer path of the C ABI (bank uploaded from pinned host memory and the
        next bank + tally bins read back every step; wall clock, max over ranks)
--impl reference : the CPU restatement of SCONE's own OpenMP history loop (oracle/, the reference itself is
        Fortran and cannot be compiled in this image) on all host threads, same deck / pop / metric.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DECKS = {"c5g7": "decks/c5g7/c5g7_2d", "c5g7_3d": "decks/c5g7/c5g7_3d_rodded", "inf": "decks/urr/inf", "slab": "decks/urr/slab",
         "ce_pin": "decks/ce/pincell", "ce_asm": "decks/ce/assembly17", "can": "decks/mg/can"}
WORKLOAD = {"c5g7": "C5G7 MOX 2D 7-group eigenvalue, delta tracking (InputFiles/Benchmarks/Multigroup/C5G7 as decks/c5g7/c5g7_2d)",
            "c5g7_3d": "C5G7 3D rodded-A 7-group eigenvalue with 34x34x9 flux+fission mesh, delta tracking",
            "inf": "SCONE_Inf URRa-2-1-IN 2-group infinite medium", "slab": "SCONE_Slab URRa-2-1-SL 2-group slab (P1)",
            "ce_asm": "synthetic continuous-energy 17x17 assembly, 20 nuclides per fuel material (5 bundled ACE nuclides + 15 energy-shifted clones), delta tracking, k-eff only (BASELINE configs[4] at a single-GPU population)",
            "can": "7-group finite can bounded by truncated cylinders (reflective bottom, vacuum top), surface tracking, strongly subcritical",
            "ce_pin": "continuous-energy U-233 / H-1 pin cell from the reference's bundled ACE nuclides (BASELINE configs[2] stand-in), 300-bin energy x material flux tally"}
ALG_BYTES_PER_SEGMENT = 124      # particle SoA read+write per flight segment (multigroup)
ALG_BYTES_PER_SCORE = 16         # f64 read-modify-write per tally score
# continuous energy, SURVEY.md section 8(d) / DESIGN.md section 4: a flight segment moves the particle record (132 B with E) and one
# total-cross-section lookup of 36 B per nuclide of the material + 20 B; a real collision walks the nuclides again (36 B each) and
# reads one micro set (148 B)
CE_BYTES_PER_SEGMENT = 132 + 20
CE_BYTES_PER_NUCLIDE_TERM = 36
CE_BYTES_PER_COLLISION = 148
DATA = {"c5g7": "NEA C5G7 benchmark 7-group constants re-authored from the published specification (decks/c5g7); uniform initial source",
        "c5g7_3d": "NEA C5G7 benchmark 7-group constants, 3-D rodded configuration A authored from the published specification; uniform initial source",
        "inf": "analytic benchmark constants URRa-2-1-IN (Sood et al.) as in the reference's InputFiles/SCONE_Inf", "slab": "analytic benchmark constants URRa-2-1-SL (Sood et al.) as in the reference's InputFiles/SCONE_Slab",
        "ce_pin": "the reference's bundled JEF-3.1.1 ACE nuclides (data/ace); authored pin-cell deck", "can": "NEA C5G7 7-group constants in an authored finite geometry",
        "ce_asm": "synthetic: the reference's bundled ACE nuclides plus seeded energy-shifted clones (decks/ce/synth); authored 17x17 deck"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.rows, self.stop_flag = device, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons, "samples": len(self.rows)}


def tracking_override(tracking):
    return "" if not tracking else " transportOperator { type transportOperator%s; }" % tracking


def oracle_rate(deck, pop, n_inactive, seconds, threads, tracking=None):
    """CPU arm: oracle (C++ restatement of SCONE's OpenMP loop) active-cycle neutrons/s on `threads` threads."""
    os.environ["OMP_NUM_THREADS"] = str(threads)
    from tests import oracle_lib as ol
    orc = ol.load()
    ov = "pop %d; inactive %d; active 1000000; seed 20261017;%s" % (pop, n_inactive, tracking_override(tracking))
    e = orc.orc_eigen_load(os.path.join(ROOT, deck).encode(), ov.encode())
    if not e:
        raise RuntimeError(ol.err(orc))
    orc.orc_eigen_init_source(e)
    k = orc.orc_eigen_keff0(e)
    for _ in range(n_inactive):
        k = orc.orc_eigen_cycle(e, 0, k)
    seg0, c0, h0 = C.c_long(), C.c_long(), C.c_long()
    orc.orc_eigen_stats(e, C.byref(seg0), C.byref(c0), C.byref(h0))
    t0 = time.perf_counter(); n = 0
    while True:
        k = orc.orc_eigen_cycle(e, 1, k); n += 1
        if time.perf_counter() - t0 > seconds:
            break
    dt = time.perf_counter() - t0
    seg1, c1, h1 = C.c_long(), C.c_long(), C.c_long()
    orc.orc_eigen_stats(e, C.byref(seg1), C.byref(c1), C.byref(h1))
    kc, ks = C.c_double(), C.c_double()
    orc.orc_eigen_keff(e, 1, C.byref(kc), C.byref(ks))
    orc.orc_eigen_free(e)
    return dict(nps=pop * n / dt, sps=(seg1.value - seg0.value) / dt, cycles=n, seconds=dt, k=k, k_cum=kc.value, k_std=ks.value)


def ce_nuclides_and_material(n_nuc=20):
    """BASELINE configs[4]-style stress: ~20 nuclides per fuel material built from the golden nuclide tables."""
    import numpy as np
    from scone_b200.ce import synthetic_nuclides
    g = np.load(os.path.join(ROOT, "tests", "golden", "ce_nuclides.npz"))
    base = [(g["grid_" + n], g["data_" + n]) for n in ("1001", "92233", "52126", "91231", "91232")]
    nuclides = synthetic_nuclides(base, n_nuc, seed=7)
    rng = np.random.default_rng(1)
    material = [(k + 1, float(rng.uniform(1e-5, 5e-2))) for k in range(n_nuc)]
    return nuclides, material


def ce_lookup_bench(device, n_lookups, cpu_seconds, with_cpu):
    """The CE XS-lookup kernel on its own: Sigma_t of 20-nuclide materials for n_lookups particles of random energy (unsorted =
    worst case for the gathers), on a library that does NOT fit in L2 (300 nuclides, > 1 GB of tables and lookup structures:
    the gathers come from HBM), unsorted and sorted by energy; and, as `l2_resident`, on the 20-nuclide library of round 1
    (4.5 MB of tables: an L2 figure). Returns the roofline object of the kernel."""
    import numpy as np
    import torch
    from scone_b200.ce import CeDatabase, large_library
    peak, peak_src = measured_peak()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % device)
    rng = np.random.default_rng(3)
    E_h = np.exp(rng.uniform(np.log(1e-11), np.log(19.0), n_lookups))

    def timed(eng, E, mat, tot, sort=False):
        ms = []
        for it in range(8):
            flush.zero_(); torch.cuda.synchronize()
            ms.append(eng.lookup_device(E.data_ptr(), mat.data_ptr(), tot.data_ptr(), 0, 0, n=n_lookups, sort=sort))
        return sum(ms[3:]) / len(ms[3:])

    # ---- the library that does not fit in L2 --------------------------------------------------------------------------
    nuclides, materials = large_library(300)
    eng = CeDatabase(nuclides, materials, device=device)
    raw, idx, tab = eng.memory()
    m_h = rng.integers(1, len(materials) + 1, n_lookups).astype(np.int32)
    E = torch.from_numpy(E_h).to(flush.device); mat = torch.from_numpy(m_h).to(flush.device)
    tot = torch.zeros(n_lookups, dtype=torch.float64, device=flush.device)
    k_ms = timed(eng, E, mat, tot)
    tots = torch.zeros_like(tot)
    k_ms_sorted = timed(eng, E, mat, tots, sort=True)          # the engine bins the lookups by (material, energy) itself: sort + lookup timed
    assert torch.equal(tots.nan_to_num(1e300), tot.nan_to_num(1e300)), "binned lookups differ from unsorted ones"
    alg = (36 * 20 + 20) * n_lookups                                 # 36 B per nuclide + 20 B per lookup
    E_p = torch.from_numpy(E_h).pin_memory(); m_p = torch.from_numpy(m_h).pin_memory(); t_p = torch.zeros(n_lookups, dtype=torch.float64).pin_memory()
    eng.lookup_into(E_p.numpy(), m_p.numpy(), t_p.numpy())
    t_e2e = 1e30
    for _ in range(3):
        t0 = time.perf_counter(); eng.lookup_into(E_p.numpy(), m_p.numpy(), t_p.numpy()); t_e2e = min(t_e2e, time.perf_counter() - t0)
    assert torch.equal(t_p, tot.cpu()), "host-buffer lookup differs from the device-resident one"
    out = {"kernel": "k_ce_lookup", "workload": "Sigma_t of 20-nuclide materials, %d lookups, random (unsorted) energies in [1e-11, 19] MeV; library of 300 nuclides "
                     "(the reference's 5 bundled ACE nuclides refined to 3.6e3 - 5.9e4 grid points and cloned with seeded energy shifts): %.0f MB of nuclide tables + "
                     "%.0f MB of lookup structures, no [union interval][nuclide] table (per-nuclide hashed index)" % (n_lookups, raw / 1e6, idx / 1e6),
           "bound": "hbm", "achieved": alg / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (k_ms * 1e-3) / 1e9 / peak,
           "traffic": None, "peak_source": peak_src, "kernel_ms_per_launch": k_ms, "algorithmic_bytes_per_launch": alg,
           "lookups_per_s": n_lookups / (k_ms * 1e-3), "e2e_lookups_per_s_host_buffers": n_lookups / t_e2e, "e2e_bytes_per_lookup": {"h2d": 12, "d2h": 8},
           "memory": {"nuclide_tables_bytes": raw, "lookup_structures_bytes": idx, "ratio_to_nuclide_tables": (raw + idx) / raw},
           "binned_by_material_and_energy": {"ms_per_call": k_ms_sorted, "achieved": alg / (k_ms_sorted * 1e-3) / 1e9, "frac": alg / (k_ms_sorted * 1e-3) / 1e9 / peak,
                                "lookups_per_s": n_lookups / (k_ms_sorted * 1e-3), "traffic": None,
                                "note": "sb_ce_lookup_sorted_device on the same unsorted batch: counting sort of the lookups by (material, energy bin) on the device + the same lookup "
                                        "kernel in that order (both inside the timed region); neighbouring lanes gather from the same sectors, results bit-identical"},
           "l2": "flushed before every timed launch (256 MiB memset)",
           "parity": "grid indices and cross sections bit-identical to the CPU restatement's binary searches (tests/test_gpu_ce.py::test_large_library_bit_exact)"}
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        out["traffic"] = tr.get("k_ce_lookup_large_dram_bytes_per_launch")
        out["binned_by_material_and_energy"]["traffic"] = tr.get("k_ce_lookup_large_binned_dram_bytes_per_call")
    except Exception:
        pass
    if with_cpu:
        from tests import ce_util
        from tests import oracle_lib as ol
        orc = ol.load()
        db, _ = ce_util.oracle_db(orc, nuclides, materials)
        m = min(n_lookups, 2_000_000)
        Ec = np.ascontiguousarray(E_h[:m]); mc = np.ascontiguousarray(m_h[:m]); oc = np.zeros(m)
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < cpu_seconds:
            orc.orc_ce_db_total_n(db, m, ol.dp(Ec), ol.ip(mc), ol.dp(oc)); reps += 1
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": m * reps / dt, "unit": "lookups/s", "cores": os.cpu_count() or 1, "kind": "port",
                               "sample": "%d x %d lookups in %.1f s (oracle: binary search per nuclide as aceNeutronNuclide%%search, OpenMP over particles)" % (reps, m, dt)}
        assert np.array_equal(oc, tot[:m].cpu().numpy()), "CE lookup differs from the oracle"
        orc.orc_ce_db_free(db)
    eng.close()
    del E, mat, tot, tots

    # ---- the 20-nuclide library of round 1: tables resident in L2 -----------------------------------------------------
    nuclides, material = ce_nuclides_and_material(20)
    eng = CeDatabase(nuclides, [material], device=device)
    E = torch.from_numpy(E_h).to(flush.device); mat = torch.ones(n_lookups, dtype=torch.int32, device=flush.device)
    tot = torch.zeros(n_lookups, dtype=torch.float64, device=flush.device)
    k2 = timed(eng, E, mat, tot)
    out["l2_resident"] = {"workload": "the same kernel on one 20-nuclide material of 4.5 MB of tables ([union interval][nuclide] table): the gathers are served by L2, the figure is not an HBM figure",
                          "kernel_ms_per_launch": k2, "achieved": alg / (k2 * 1e-3) / 1e9, "frac_of_hbm_peak": alg / (k2 * 1e-3) / 1e9 / peak, "lookups_per_s": n_lookups / (k2 * 1e-3)}
    eng.close()
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(threads)
    from tests import oracle_lib as ol
    orc = ol.load()
    deck = DECKS[args.deck]
    pop = args.pop
    ov = "pop %d; inactive %d; active 1000000; seed 20261017;%s" % (pop, args.inactive, tracking_override(args.tracking))
    e = orc.orc_eigen_load(os.path.join(ROOT, deck).encode(), ov.encode())
    if not e:
        raise RuntimeError(ol.err(orc))
    orc.orc_eigen_init_source(e)
    k = orc.orc_eigen_keff0(e)
    for _ in range(args.inactive):
        k = orc.orc_eigen_cycle(e, 0, k)
    for _ in range(args.warmup):
        k = orc.orc_eigen_cycle(e, 1, k)
    s0, c0, h0 = C.c_long(), C.c_long(), C.c_long()
    orc.orc_eigen_stats(e, C.byref(s0), C.byref(c0), C.byref(h0))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        k = orc.orc_eigen_cycle(e, 1, k)
    dt = time.perf_counter() - t0
    s1, c1, h1 = C.c_long(), C.c_long(), C.c_long()
    orc.orc_eigen_stats(e, C.byref(s1), C.byref(c1), C.byref(h1))
    val = pop * args.steps / dt
    kc, ks = C.c_double(), C.c_double()
    orc.orc_eigen_keff(e, 1, C.byref(kc), C.byref(ks))
    line = {
        "impl": "reference", "metric": "active-cycle neutrons/s", "value": val, "unit": "neutrons/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": DATA[args.deck],
        "config": {"workload": WORKLOAD[args.deck], "deck": deck, "pop_per_cycle": pop, "tracking": args.tracking or "as the deck (%s)" % ("ST, cache" if args.deck in ("ce_pin", "can") else "DT"),
                   "note": "CPU reference arm: one step = one active cycle of pop histories on the host cores"},
        "segments_per_s": (s1.value - s0.value) / dt, "keff": kc.value, "keff_std": ks.value,
        "cpu_baseline": {"value": val, "unit": "neutrons/s", "cores": threads, "kind": "port",
                         "sample": "%d active cycles of %d histories (oracle: C++/OpenMP restatement of SCONE's history loop; "
                                   "SCONE itself is Fortran and no Fortran compiler exists in the image)" % (args.steps, pop)},
        "e2e": {"value": val, "unit": "neutrons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own")
    ap.add_argument("--deck", default="c5g7", choices=sorted(DECKS))
    ap.add_argument("--pop", type=int, default=100000, help="histories per cycle PER GPU (weak scaling)")
    ap.add_argument("--inactive", type=int, default=10, help="untimed inactive cycles before the active phase")
    ap.add_argument("--tracking", default=None, choices=["DT", "ST", "HT"], help="transportOperator; default: what the deck says (delta tracking for the MG decks as BASELINE configs[0] names it, surface tracking with cache for the CE pin cell)")
    ap.add_argument("--no-rank-parity", action="store_true", help="N > 1: skip the check that the ranks reproduce the single-rank k-eff")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="N > 1: per-cycle exchange between the ranks through peer memory (engine kernels storing into the other GPUs' HBM over NVLink) or through NCCL collectives")
    ap.add_argument("--no-l2-flush", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the large-population and CE-lookup measurements")
    ap.add_argument("--large-pop", type=int, default=1000000)
    ap.add_argument("--ce-lookups", type=int, default=10000000)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import scone_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    deck = os.path.join(ROOT, DECKS[args.deck])
    pop = args.pop
    total_pop = pop * world
    ov = "pop %d; inactive %d; active %d; seed 20261017;%s" % (
        total_pop, args.inactive, args.warmup + 2 * args.steps + 4, tracking_override(args.tracking))
    sampler = ClockSampler(local); sampler.start()
    pp = scone_b200.EigenPhysicsPackage(deck, ov, device=local, rank=rank, n_ranks=world)
    comm = scone_b200.distributed.TorchComm(device=torch.device("cuda", local)) if world > 1 else None
    L = pp.L
    eng = pp.engine
    exchange = ""
    if world > 1:
        peer = args.exchange == "peer" and scone_b200.distributed.enable_peer(pp, comm)
        exchange = ("; per cycle: score sums, bank sizes and the load-balancing sites stored by the engines' kernels into the other GPUs' memory "
                    "(CUDA IPC over NVLink), no collective, one host synchronisation") if peer else \
                   "; per cycle: one all-gather of 7 f64 (score sums + bank size), neighbour send/recv of boundary sites (NCCL)"
    pp.generateInitialState()
    pp.cycles(False, args.inactive, comm=comm)
    for _ in range(max(3, args.warmup)):
        pp.cycle(True, comm=comm)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def maxreduce(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumreduce(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    flush = not args.no_l2_flush
    FLUSH_BYTES = 256 << 20
    # ---- device-resident arm: K active cycles, CUDA events on the engine stream --------------------------
    L.sb_profile_enable(eng, 1)
    launches0 = pp.launch_count()
    barrier()
    ms_total = 0.0; seg = 0; scores = 0; nsites = 0; coll = 0; xs_terms = 0
    ms = C.c_double()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        if flush:
            L.sb_flush_l2(eng, FLUSH_BYTES)
        L.sb_timer_begin(eng)
        res = pp.cycle(True, comm=comm)
        L.sb_timer_end(eng, C.byref(ms))
        ms_total += ms.value; seg += res.n_segments; scores += res.n_scores; nsites += res.n_sites; coll += res.n_collisions; xs_terms += res.n_xs_terms
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = pp.launch_count() - launches0 - (args.steps if flush else 0) * 0
    msk, nl, segp, scp = C.c_double(), C.c_int64(), C.c_int64(), C.c_int64()
    L.sb_profile_read(eng, C.byref(msk), C.byref(nl), C.byref(segp), C.byref(scp))
    ms_wait, ms_tail = C.c_double(), C.c_double()
    L.sb_profile_peer_stages(eng, C.byref(ms_wait), C.byref(ms_tail))
    L.sb_profile_enable(eng, 0)
    ms_max = maxreduce(ms_total)
    value = total_pop * args.steps / (ms_max * 1e-3)
    seg_all = sumreduce(float(seg))
    k_dev, k_std = pp.k, res.k_cum_std
    max_seg = int(maxreduce(float(res.max_history_segments)))

    # ---- end-to-end arm: bank in pinned HOST memory, copied in and out every step ------------------------
    pp.cycle(True, host_buffers=True, comm=comm)          # warm the host buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pp.cycle(True, host_buffers=True, comm=comm)
    barrier()
    t_e2e = maxreduce(time.perf_counter() - t0)
    h2d, d2h = pp.host_bytes(True)
    e2e_val = total_pop * args.steps / t_e2e
    sampler.stop_flag = True; sampler.join(timeout=2)

    # ---- N > 1: the ranks together against ONE rank on the same histories (hardware check of the exchange) ----
    rank_parity = None
    if world > 1 and not args.no_rank_parity:
        ovp = "pop %d; inactive 3; active 2; seed 20261017;%s" % (20000 * world, tracking_override(args.tracking))
        comm2 = scone_b200.distributed.TorchComm(device=torch.device("cuda", local))
        ppr = scone_b200.EigenPhysicsPackage(deck, ovp, device=local, rank=rank, n_ranks=world)
        if args.exchange == "peer":
            scone_b200.distributed.enable_peer(ppr, comm2)
        ppr.generateInitialState()
        ks = []
        for c in range(5):
            ppr.cycle(c >= 3, comm=comm2); ks.append(ppr.k)
        ppr.close()
        if rank == 0:
            pp1 = scone_b200.EigenPhysicsPackage(deck, ovp, device=local)
            pp1.generateInitialState()
            ks1 = []
            for c in range(5):
                pp1.cycle(c >= 3); ks1.append(pp1.k)
            pp1.close()
            diff = max(abs(a - b) / abs(b) for a, b in zip(ks, ks1))
            rank_parity = {"cycles": 5, "histories_per_cycle": 20000 * world, "ranks": world, "k_of_the_ranks": ks[-1], "k_of_one_rank": ks1[-1],
                           "max_rel_k_difference": diff, "bound": 1e-12, "ok": bool(diff <= 1e-12)}
            if diff > 1e-12:
                print("bench.py: the %d ranks and the single rank disagree on k-eff: %r vs %r" % (world, ks, ks1), file=sys.stderr)

    # ---- roofline of the dominant kernel (k_histories) -----------------------------------------------------
    peak, peak_src = measured_peak()
    is_ce = args.deck.startswith("ce_")
    if is_ce:          # mean nuclides per lookup from the engine's counter; the collision walks the same material again
        nbar = xs_terms / max(1, seg)
        alg_bytes = (CE_BYTES_PER_SEGMENT * seg + CE_BYTES_PER_NUCLIDE_TERM * xs_terms + (CE_BYTES_PER_COLLISION + CE_BYTES_PER_NUCLIDE_TERM * nbar) * coll
                     + ALG_BYTES_PER_SCORE * scores) / max(1, args.steps)
    else:
        alg_bytes = (ALG_BYTES_PER_SEGMENT * segp.value + ALG_BYTES_PER_SCORE * scp.value) / max(1, nl.value)
    k_ms = msk.value / max(1, nl.value)
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(("k_histories_ce" if args.deck.startswith("ce_") else "k_histories") + "_dram_bytes_per_launch")     # (k_histories: with k_lone)
    except Exception:
        pass
    kname = "k_histories_ce" if args.deck.startswith("ce_") else ("k_histories + k_lone" if args.tracking in (None, "DT") else "k_histories_track")
    roofline = {"kernel": kname, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel_ms_per_launch": k_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_share_of_step": (msk.value / ms_total) if ms_total > 0 else None,
                "algorithmic_bytes": ("per segment 132 + 20 B + 36 B per nuclide term of the total cross section (%.2f terms per segment counted by the engine); per real collision 148 B + the nuclide walk again; 16 B per score"
                                      % (xs_terms / max(1, seg))) if is_ce else "124 B per flight segment + 16 B per tally score",
                # what actually bounds the kernel at this population: the dependent chain of the cycle's longest history
                "latency": {"longest_history_flights": max_seg, "kernel_us_per_flight_of_the_longest_history": (1e3 * k_ms / max_seg) if max_seg > 256 else None,
                            "mean_flights_per_history": seg / max(1, pop * args.steps),
                            "note": "the kernel ends when the longest history of the cycle ends: its flights x the latency of one flight + collision round of a history that runs alone in its warp (profiles/README.md, round 2)"},
                "note": ("lane-resident histories in CTA-lockstep phases: bound by instruction fetch, barrier wait and memory latency at 16-32 warps per SM, not by HBM (DESIGN.md section 4, profiles/README.md)"
                         if not kname.startswith("k_histories +") else "register-resident histories (k_histories) and the last histories of every warp one per warp (k_lone, running beside it): latency bound, not HBM bound; kernel_ms_per_launch is from the first launch of k_histories to the end of k_lone and k_finish_sites (DESIGN.md section 4)")}

    # ---- extras (N = 1 only): the same deck at a GPU-sized population, and the CE XS-lookup kernel ----------------
    large = None; ce = None
    if world == 1 and not args.no_extras:
        pp.close()
        ov2 = "pop %d; inactive 6; active 100; seed 20261017;%s" % (args.large_pop, tracking_override(args.tracking))
        pl = scone_b200.EigenPhysicsPackage(deck, ov2, device=local)
        pl.generateInitialState(); pl.cycles(False, 6)
        for _ in range(3):
            pl.cycle(True)
        tms = 0.0; segl = 0; nst = 8
        for _ in range(nst):
            L.sb_flush_l2(pl.engine, FLUSH_BYTES); L.sb_timer_begin(pl.engine)
            rl = pl.cycle(True)
            L.sb_timer_end(pl.engine, C.byref(ms)); tms += ms.value; segl += rl.n_segments
        large = {"pop_per_cycle": args.large_pop, "value": args.large_pop * nst / (tms * 1e-3), "unit": "neutrons/s", "ms_per_step": tms / nst,
                 "segments_per_s": segl / (tms * 1e-3), "keff": pl.k, "steps": nst,
                 "note": "same deck and engine at a population that fills the GPU (BASELINE configs[3] runs 1.25e6 histories per GPU); "
                         "at 1e5 histories per cycle the kernel time is the critical path of the longest history"}
        pl.close()
        ce = ce_lookup_bench(local, args.ce_lookups, min(args.cpu_seconds, 6.0), not args.no_cpu_baseline)
        pp = None

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            r = oracle_rate(DECKS[args.deck], pop, 3, args.cpu_seconds, threads, args.tracking)
            comb = (r["k_std"] ** 2 + k_std ** 2) ** 0.5
            cpu = {"value": r["nps"], "unit": "neutrons/s", "cores": threads, "kind": "port", "segments_per_s": r["sps"],
                   "keff": r["k_cum"], "keff_std": r["k_std"],
                   "keff_delta_in_combined_sigma": (k_dev - r["k_cum"]) / comb if comb > 0 else None,
                   "sample": "%d active cycles of %d histories in %.1f s (oracle: C++/OpenMP restatement of SCONE's history loop; "
                             "SCONE is Fortran and cannot be compiled in this image)" % (r["cycles"], pop, r["seconds"])}
        line = {
            "metric": "active-cycle neutrons/s", "value": value, "unit": "neutrons/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": DATA[args.deck],
            "config": {"workload": WORKLOAD[args.deck], "deck": DECKS[args.deck], "pop_per_cycle_per_gpu": pop, "pop_per_cycle_total": total_pop,
                       "tracking": args.tracking or "as the deck (%s)" % ("ST, cache" if args.deck in ("ce_pin", "can") else "DT"), "inactive_cycles_before": args.inactive,
                       "l2": "flushed between timed steps (256 MiB memset, untimed)" if flush else "not flushed (steady-state cycles)",
                       "parallelism": "bank sharded by history index over %d GPU(s)%s" % (
                           world, exchange)},
            "longest_history_segments": max_seg, "virtual_per_real_collision": (seg - coll) / max(1, coll),
            "segments_per_s": seg_all / (ms_max * 1e-3), "segments_per_history": seg / max(1, pop * args.steps),
            "keff": k_dev, "keff_std": k_std, "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps,
            "e2e": {"value": e2e_val, "unit": "neutrons/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "ms_per_step": 1e3 * t_e2e / args.steps},
            "gpu_launches": launches, "roofline": roofline, "clocks": sampler.summary(),
        }
        if rank_parity is not None:
            line["rank_parity"] = rank_parity
        if world > 1:
            line["stage_ms_rank0"] = {"histories": msk.value / max(1, nl.value), "waiting_for_every_rank_sums": ms_wait.value / args.steps,
                                      "close_resample_balance": ms_tail.value / args.steps}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if large is not None:
            line["large_population"] = large
        if ce is not None:
            line["roofline_ce_lookup"] = ce
    if pp is not None:
        pp.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
