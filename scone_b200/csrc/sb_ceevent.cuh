// Continuous-energy histories as EVENT QUEUES inside a persistent CTA (the north star's formulation, sized to what was measured).
//
// k_histories_ce (sb_cehist.cuh) binds a history to a lane: with 1e5 histories on 75 776 lanes, 6.5 of 32 lanes are active per
// instruction (profiles/cehist_r01g_summary.txt) -- a warp pays for an event if a single one of its lanes needs it.  Here the
// history lives in a SLOT in global memory (L2 resident: scalars + coordList + distance cache, 1.1 KB) and every event phase first
// builds, in shared memory, the queue of slots that need it, then all threads of the CTA drain the queue: warps are full, and all
// lanes of a warp run the same event.  Per round:
//
//   refill   dead slots claim the next histories of the bank (warp-aggregated atomics)            -> queue ALIVE
//   flight   ALIVE: tracking XS, distance, delta-tracking teleport or surface-tracking move       -> queues VIRT, COLL (leak: history ends)
//   tally    VIRT: tallyAdmin%reportInColl of the virtual collisions
//   collide  COLL: nuclide, channel, reportInColl, number of implicit fission sites               -> queues FISS, EFIX, EGAS, INEL (capture: ends)
//   fission  FISS: sites into the next-cycle bank
//   scatter  EFIX, EGAS, INEL drained one after the other (target at rest / free gas / MT laws), cut-off, new union interval
//
// Phases are separated by CTA barriers; CTAs are independent (no grid barrier): a CTA leaves when the bank is exhausted and its
// slots are dead.  The arithmetic of every event is the device function k_histories_ce calls, so results are bit-identical; the
// order in which slots are served does not matter (per-history random stream, sites keyed (history, sequence), atomic tallies).
// References: as sb_cehist.cuh.
#pragma once
#include "sb_cehist.cuh"

namespace sbc {

// slot storage: scalars as structure of arrays (a warp drains consecutive queue entries = mostly consecutive slots: coalesced),
// coordList + distance cache as one record per slot (indexed dynamically by nesting level)
struct CeSlotGeom { sbt::Coords c; sbt::DistCache cache; };
struct CeSlots {
  double *E, *w, *w0, *trackXS, *majXS, *sigTot, *sProd, *sAbs, *sScat;
  uint64_t* rng;
  int *hi, *nSite, *hSeg, *mode, *u, *alive, *MT, *nuc0, *nNew, *site0;
  CeSlotGeom* g;
  static constexpr int N_F64 = 10, N_I32 = 10;       // rng counted with the doubles
  static size_t bytes(size_t n) { return n * (N_F64 * 8 + N_I32 * 4 + sizeof(CeSlotGeom)) + 256; }
  void carve(char* base, size_t n) {
    double* d = (double*)base;
    E = d; w = d + n; w0 = d + 2 * n; trackXS = d + 3 * n; majXS = d + 4 * n; sigTot = d + 5 * n; sProd = d + 6 * n; sAbs = d + 7 * n; sScat = d + 8 * n;
    rng = (uint64_t*)(d + 9 * n);
    int* i = (int*)(d + 10 * n);
    hi = i; nSite = i + n; hSeg = i + 2 * n; mode = i + 3 * n; u = i + 4 * n; alive = i + 5 * n; MT = i + 6 * n; nuc0 = i + 7 * n; nNew = i + 8 * n; site0 = i + 9 * n;
    size_t off = n * (N_F64 * 8 + N_I32 * 4); off = (off + 15) & ~(size_t)15;
    g = (CeSlotGeom*)(base + off);
  }
};

struct CeEventArgs { CeArgs a; CeSlots S; };

enum { Q_ALIVE = 0, Q_VIRT, Q_COLL, Q_FISS, Q_EFIX, Q_EGAS, Q_INEL, Q_COUNT };

template <int SLOTS>
struct EventQueues {
  int n[Q_COUNT];
  unsigned short q[Q_COUNT][SLOTS];
};
template <int SLOTS>
__device__ __forceinline__ void qPush(EventQueues<SLOTS>& Q, int which, bool pred, int slot) {      // warp-aggregated push (all lanes of the warp call)
  const unsigned m = __ballot_sync(0xffffffffu, pred);
  if (m == 0u) return;
  const int lane = threadIdx.x & 31;
  int b = 0;
  if (lane == __ffs(m) - 1) b = atomicAdd(&Q.n[which], __popc(m));
  b = __shfl_sync(0xffffffffu, b, __ffs(m) - 1);
  if (pred) Q.q[which][b + __popc(m & ((1u << lane) - 1u))] = (unsigned short)slot;
}

// history end: what the cycle close reads (per-history scores, site count)
__device__ __forceinline__ void slotDie(const CeArgs& a, const CeSlots& S, int g, double leak) {
  const int hi = S.hi[g], hSeg = S.hSeg[g];
  a.nsites[hi] = S.nSite[g];
  a.hProd[hi] = S.sProd[g]; a.hAbs[hi] = S.sAbs[g]; a.hLeak[hi] = leak; a.hScat[hi] = S.sScat[g];
  if (hSeg > 256) atomicMax(&a.cd->maxSeg, hSeg);
  S.alive[g] = 0;
}

template <int THREADS, int SLOTS>
__global__ void __launch_bounds__(THREADS, 1) k_events_ce(const CeEventArgs ea) {
  const CeArgs& a = ea.a;
  const char* base = a.blob;
  __shared__ CeCtx s_ctx;
  __shared__ EventQueues<SLOTS> Q;
  __shared__ int s_exhausted;
  if (threadIdx.x == 0) {
    s_ctx.M = a.M; s_ctx.T = bind(a.M, base); s_ctx.ce = a.ce; s_ctx.bins = a.bins; s_ctx.phase = a.phase; s_ctx.needMacro = a.needMacro; s_ctx.impScores = a.impScores;
    s_exhausted = 0;
  }
  const CeSlots& S = ea.S;
  const int g0 = blockIdx.x * SLOTS;                  // this CTA's slots are [g0, g0 + SLOTS)
  for (int i = threadIdx.x; i < SLOTS; i += THREADS) S.alive[g0 + i] = 0;
  __syncthreads();
  const CeCtx& ctx = s_ctx;
  const Tables& T = s_ctx.T;
  const Model& M = s_ctx.M;
  const sbce::CeDev& X = s_ctx.ce.xs;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned ltMask = (1u << lane) - 1u;
  const double collisionXS = M.collisionXS;
  unsigned nSeg = 0, nColl = 0, nScore = 0, nTerms = 0;
  constexpr int NPASS = (SLOTS + THREADS - 1) / THREADS;

  for (;;) {
    if (threadIdx.x < Q_COUNT) Q.n[threadIdx.x] = 0;
    const int exhaustedBefore = s_exhausted;          // read before the barrier: the refill phase below is the only writer
    __syncthreads();
    // ---------------- refill: dead slots claim histories ----------------------------------------------------------
    for (int pss = 0; pss < NPASS; ++pss) {
      const int si = pss * THREADS + threadIdx.x;
      const bool inRange = si < SLOTS;
      const int g = g0 + si;
      bool alive = inRange && S.alive[g] != 0;
      const bool want = inRange && !alive && exhaustedBefore == 0;
      const unsigned need = __ballot_sync(FULL, want);
      if (need) {
        int b = 0;
        if (lane == __ffs(need) - 1) b = atomicAdd(&a.cd->nextHistory, __popc(need));
        b = __shfl_sync(FULL, b, __ffs(need) - 1);
        if (b + __popc(need) >= a.n) s_exhausted = 1;                // every writer writes 1; read again after the barrier
        const int my = b + __popc(need & ltMask);
        if (want && my < a.n) {
          CeSlotGeom& sg = S.g[g];
          S.hi[g] = my;
          sg.c.r[0][0] = a.in.rx[my]; sg.c.r[0][1] = a.in.ry[my]; sg.c.r[0][2] = a.in.rz[my];
          sg.c.u[0][0] = a.in.ux[my]; sg.c.u[0][1] = a.in.uy[my]; sg.c.u[0][2] = a.in.uz[my];
          const double w = a.in.w[my], E = a.in.E[my];
          S.w[g] = w; S.w0[g] = w; S.E[g] = E;
          S.rng[g] = sbh::rngSeed(a.seedTab, a.rng0, (unsigned)(a.histOffset + my + 1));
          sg.c.nesting = 1; sg.c.mat = SB_UNDEF_MAT; sg.c.uid = -3; sg.cache.lvl = 0;
          if (!sbt::placeCoord(M, T, sg.c)) atomicMax(&a.cd->error, SB_ERR_NEST);
          S.nSite[g] = 0; S.hSeg[g] = 0; S.sProd[g] = 0.0; S.sAbs[g] = 0.0; S.sScat[g] = 0.0; S.mode[g] = 0; S.trackXS[g] = 1.0; S.sigTot[g] = 0.0;
          int u = sbce::unionSearch(X, E);
          if (u == 0) { atomicMax(&a.cd->error, SB_ERR_CE_ENERGY); u = 1; }
          S.u[g] = u;
          S.majXS[g] = fmax(majorantAt(X, u, E) + 0.0, collisionXS);
          S.alive[g] = 1; alive = true;
        }
      }
      qPush(Q, Q_ALIVE, alive, si);
    }
    __syncthreads();
    const int nAlive = Q.n[Q_ALIVE];
    if (nAlive == 0 && s_exhausted) break;

    // ---------------- flight -----------------------------------------------------------------------------------------
    for (int i0 = 0; i0 < nAlive; i0 += THREADS) {
      const int i = i0 + threadIdx.x;
      const bool on = i < nAlive;
      const int si = on ? Q.q[Q_ALIVE][i] : 0;
      bool realColl = false, scoreVirt = false;
      if (on) {
        const int g = g0 + si;
        CeSlotGeom& s = S.g[g];
        const double E = S.E[g]; const int u = S.u[g];
        const double majXS = S.majXS[g], wgt = S.w[g];
        int mode = S.mode[g];
        bool died = false; double leak = 0.0;
        if (mode == 0) {
          if (a.tracking == SB_TRACK_DT) mode = 1;
          else if (a.tracking == SB_TRACK_ST) mode = 2;
          else {
            double majorant_inv = 1.0 / majXS;
            double sigmaT = (s.c.mat == SB_VOID_MAT) ? 0.0 : ceMatTotal(X, u, E, s.c.mat, nTerms) + 0.0;
            double ratio = sigmaT * majorant_inv;
            mode = (ratio > (1.0 - a.htCutoff)) ? 1 : 2;
          }
          s.cache.lvl = 0;
          S.mode[g] = mode;
        }
        uint64_t rng = S.rng[g];
        if (mode == 1) {
          const double trackXS = majXS;
          S.trackXS[g] = trackXS;
          double majorant_inv = 1.0 / trackXS;
          double distance = -sbk::kLog(rngGet(rng)) * majorant_inv;
          sbt::geomTeleportCoords(M, T, s.c, distance);
          ++nSeg; S.hSeg[g] += 1;
          const int m = s.c.mat;
          if (m == SB_OUTSIDE_MAT) { leak = wgt; died = true; }
          else if (m >= SB_OVERLAP_MAT && m != SB_VOID_MAT) { atomicMax(&a.cd->error, m == SB_UNDEF_MAT ? SB_ERR_UNDEF_MAT : SB_ERR_OVERLAP_MAT); died = true; }
          else {
            bool virt = true;
            double sigTot = 0.0;
            if (m != SB_VOID_MAT) {
              sigTot = ceMatTotal(X, u, E, m, nTerms);
              if (rngGet(rng) < (sigTot + 0.0) * majorant_inv) { realColl = true; virt = false; }
            }
            S.sigTot[g] = sigTot;
            scoreVirt = virt;
          }
        } else {
          const double tol = 1.0E-12;
          int m = s.c.mat;
          const double sigTot = (m == SB_VOID_MAT) ? 0.0 : ceMatTotal(X, u, E, m, nTerms);
          S.sigTot[g] = sigTot;
          double sigmaTrack = (m == SB_VOID_MAT) ? collisionXS : fmax(sigTot + 0.0, collisionXS);
          S.trackXS[g] = sigmaTrack;
          double dist, invSigmaTrack, sigmaT;
          if (sigmaTrack < tol) { dist = INF; invSigmaTrack = INF; sigmaT = 0.0; }
          else {
            invSigmaTrack = 1.0 / sigmaTrack;
            dist = -sbk::kLog(rngGet(rng)) * invSigmaTrack;
            sigmaT = sigTot + 0.0;
          }
          int event;
          sbt::geomMove(M, T, s.c, dist, event, a.stCache ? &s.cache : nullptr);
          ++nSeg; S.hSeg[g] += 1;
          m = s.c.mat;
          if (m == SB_OUTSIDE_MAT) { leak = wgt; died = true; }
          else if (m >= SB_OVERLAP_MAT && m != SB_VOID_MAT) { atomicMax(&a.cd->error, m == SB_UNDEF_MAT ? SB_ERR_UNDEF_MAT : SB_ERR_OVERLAP_MAT); died = true; }
          else if (event == sbt::COLL_EV) {
            if (rngGet(rng) < sigmaT * invSigmaTrack) realColl = true;
            else scoreVirt = true;
          }
        }
        S.rng[g] = rng;
        if (died) slotDie(a, S, g, leak);
      }
      qPush(Q, Q_VIRT, scoreVirt, si);
      qPush(Q, Q_COLL, realColl, si);
    }
    __syncthreads();

    // ---------------- tallyAdmin%reportInColl of the virtual collisions ---------------------------------------------------
    {
      const int nV = Q.n[Q_VIRT];
      for (int i = threadIdx.x; i < nV; i += THREADS) {
        const int g = g0 + Q.q[Q_VIRT][i];
        const CeSlotGeom& s = S.g[g];
        double sProd = S.sProd[g], sAbs = S.sAbs[g];
        scoreInCollCE(ctx, base, s.c.r[0], s.c.mat, S.E[g], S.u[g], S.w[g], S.trackXS[g], S.sigTot[g], true, sProd, sAbs, nScore);
        if (a.impScores) { S.sProd[g] = sProd; S.sAbs[g] = sAbs; }
      }
    }
    __syncthreads();

    // ---------------- collision sampling: nuclide, channel, implicit fission count -------------------------------------------
    {
      const int nC = Q.n[Q_COLL];
      for (int i0 = 0; i0 < nC; i0 += THREADS) {
        const int i = i0 + threadIdx.x;
        const bool on = i < nC;
        const int si = on ? Q.q[Q_COLL][i] : 0;
        int MT = 0, nNew = 0;
        bool fixedEl = false;
        const int g = g0 + si;
        if (on) {
          const CeSlotGeom& s = S.g[g];
          const double E = S.E[g]; const int u = S.u[g]; const int mat = s.c.mat;
          const double sigTot = S.sigTot[g], wgt = S.w[g];
          uint64_t rng = S.rng[g];
          (void)rngGet(rng);
          double rem = (sigTot * 1.0) * rngGet(rng);
          const int k0 = __ldg(X.matOff + mat - 1), k1 = __ldg(X.matOff + mat);
          int nuc0 = -1;
#pragma unroll 1
          for (int k = k0; k < k1; ++k) {
            const int nn = __ldg(X.matNuc + k) - 1;
            const int idx = sbce::nucIndex(X, u, E, nn);
            double E_low, E_top, s_low, s_top;
            sbce::ldPair(X.pairTot + 4 * (__ldg(X.pairOff + nn) + (idx - 1)), E_low, E_top, s_low, s_top);
            const double f = (E - E_low) / (E_top - E_low);
            const double tot = s_top * f + (1.0 - f) * s_low;
            rem = rem - tot * (__ldg(X.matDens + k) * 1.0);
            if (rem < 0.0) { nuc0 = nn; break; }
          }
          if (nuc0 < 0) { atomicMax(&a.cd->error, SB_ERR_SAMPLING); nuc0 = __ldg(X.matNuc + k1 - 1) - 1; }
          const NucPoint p = nucPoint(X, u, E, nuc0);
          double mic[8]; nucMicro(p, mic);
          const double rr = rngGet(rng);
          {
            int C = 1;
            double xs = mic[0] * rr - mic[1];
            if (xs > 0.0) C += 1;
            xs = xs - mic[2];
            if (xs > 0.0) C += 1;
            xs = xs - mic[3];
            if (xs > 0.0) C += 1;
            MT = C;
          }
          ++nColl;
          double sProd = S.sProd[g], sAbs = S.sAbs[g];
          scoreInCollCE(ctx, base, s.c.r[0], mat, E, u, wgt, S.trackXS[g], sigTot, false, sProd, sAbs, nScore);
          if (a.impScores) { S.sProd[g] = sProd; S.sAbs[g] = sAbs; }
          const CeNucRec& N = ctx.ce.nuc[nuc0];
          if (N.fissile) {
            double rand1 = rngGet(rng);
            nNew = (int)(fabs((wgt * mic[5]) / (S.w0[g] * mic[0] * a.k_eff)) + rand1);
            if (nNew < 0) nNew = 0;
          }
          if (MT == 1) fixedEl = (E > N.kT * ctx.ce.threshE) && (N.awr > ctx.ce.threshA);
          S.rng[g] = rng; S.MT[g] = MT; S.nuc0[g] = nuc0; S.nNew[g] = nNew;
        }
        // fission-bank slots: one atomic per warp
        {
          const unsigned spawn = __ballot_sync(FULL, nNew > 0);
          if (spawn) {
            int inc = nNew;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += t; }
            int total = __shfl_sync(FULL, inc, 31);
            int b = 0;
            if (lane == 0) b = atomicAdd(&a.cd->nSites, total);
            b = __shfl_sync(FULL, b, 0);
            int site0 = b + inc - nNew;
            if (b + total > a.cap) { atomicMax(&a.cd->error, SB_ERR_BANK_OVERFLOW); site0 = -1; }
            if (nNew > 0) S.site0[g] = site0;
          }
        }
        qPush(Q, Q_FISS, nNew > 0, si);
        qPush(Q, Q_EFIX, on && MT == 1 && fixedEl, si);
        qPush(Q, Q_EGAS, on && MT == 1 && !fixedEl, si);
        qPush(Q, Q_INEL, on && MT == 2, si);
      }
    }
    __syncthreads();

    // ---------------- implicit fission sites -----------------------------------------------------------------------------------
    {
      const int nF = Q.n[Q_FISS];
      for (int i = threadIdx.x; i < nF; i += THREADS) {
        const int g = g0 + Q.q[Q_FISS][i];
        const CeSlotGeom& s = S.g[g];
        const CeNucRec& N = ctx.ce.nuc[S.nuc0[g]];
        const Tape tp{ctx.ce.tape, N.base};
        int kerr = 0;
        uint64_t rng = S.rng[g];
        const double wSite = fsign(S.w0[g], S.w[g]), Ein = S.E[g];
        const int nNew = S.nNew[g], site0 = S.site0[g], hi = S.hi[g], nSite = S.nSite[g];
#pragma unroll 1
        for (int k = 0; k < nNew; ++k) {
          double mu, phi, E_out;
          sbk::tapeSampleFission(tp, N, Ein, rng, mu, phi, E_out, &kerr);
          double d[3] = {s.c.u[0][0], s.c.u[0][1], s.c.u[0][2]};
          ceRotate(d, mu, phi);
          if (E_out > ctx.ce.maxE) E_out = ctx.ce.maxE;
          if (site0 >= 0) {
            const int o = site0 + k;
            a.out.rx[o] = s.c.r[0][0]; a.out.ry[o] = s.c.r[0][1]; a.out.rz[o] = s.c.r[0][2];
            a.out.ux[o] = d[0]; a.out.uy[o] = d[1]; a.out.uz[o] = d[2];
            a.out.w[o] = wSite * 1.0; a.out.G[o] = 0; a.out.E[o] = E_out; a.out.brood[o] = hi; a.out.seq[o] = nSite + k;
          }
        }
        S.nSite[g] = nSite + nNew;
        S.rng[g] = rng;
        if (kerr) atomicMax(&a.cd->error, SB_ERR_CE_DATA);
      }
    }
    __syncthreads();

    // ---------------- the channel: capture / fission end the history; scattering queues drained one kind at a time -------------
    {
      const int nC = Q.n[Q_COLL];
      for (int i = threadIdx.x; i < nC; i += THREADS) {
        const int g = g0 + Q.q[Q_COLL][i];
        if (S.MT[g] >= 3) slotDie(a, S, g, 0.0);
      }
    }
#pragma unroll 1
    for (int kind = Q_EFIX; kind <= Q_INEL; ++kind) {
      const int nK = Q.n[kind];
      for (int i = threadIdx.x; i < nK; i += THREADS) {
        const int g = g0 + Q.q[kind][i];
        CeSlotGeom& s = S.g[g];
        const int nuc0 = S.nuc0[g];
        const CeNucRec& N = ctx.ce.nuc[nuc0];
        const Tape tp{ctx.ce.tape, N.base};
        int kerr = 0;
        uint64_t rng = S.rng[g];
        double E = S.E[g];
        const double wPre = S.w[g];
        int MTout = 2;
        if (kind == Q_EFIX) {
          double mu = sbk::tapeSampleMu(tp, N.elAng, N.andPos, E, rng, &kerr);
          double phi = rngGet(rng) * sbk::TWO_PI;
          double E_out = E;
          asymptoticScatter(E_out, mu, N.awr);
          double d[3] = {s.c.u[0][0], s.c.u[0][1], s.c.u[0][2]};
          ceRotate(d, mu, phi);
          sbt::coordsRotate(T, s.c, d);
          E = E_out;
        } else if (kind == Q_EGAS) {
          const double A = N.awr, kT = N.kT;
          const double dir_pre[3] = {s.c.u[0][0], s.c.u[0][1], s.c.u[0][2]};
          const double sqE = sqrt(E);
          double V_n[3] = {dir_pre[0] * sqE, dir_pre[1] * sqE, dir_pre[2] * sqE};
          const double Y = sqrt(A * E / kT);
          double Xt, mut;
          sampleTargetVelocity(Y, rng, Xt, mut);
          const double r1 = rngGet(rng);
          const double phit = 2.0 * sbk::PI * r1;
          double V_t[3] = {dir_pre[0], dir_pre[1], dir_pre[2]};
          ceRotate(V_t, mut, phit);
          const double sc = Xt * sqrt(kT / A);
          double V_cm[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) { V_t[k] = V_t[k] * sc; V_cm[k] = (V_n[k] + V_t[k] * A) / (A + 1); V_n[k] = V_n[k] - V_cm[k]; }
          double U_n = sqrt(V_n[0] * V_n[0] + V_n[1] * V_n[1] + V_n[2] * V_n[2]);
#pragma unroll
          for (int k = 0; k < 3; ++k) V_n[k] = V_n[k] / U_n;
          double mu = sbk::tapeSampleMu(tp, N.elAng, N.andPos, E, rng, &kerr);
          double phi = rngGet(rng) * sbk::TWO_PI;
          ceRotate(V_n, mu, phi);
#pragma unroll
          for (int k = 0; k < 3; ++k) { V_n[k] = V_n[k] * U_n; V_n[k] = V_n[k] + V_cm[k]; }
          U_n = sqrt(V_n[0] * V_n[0] + V_n[1] * V_n[1] + V_n[2] * V_n[2]);
          double dir_post[3] = {V_n[0] / U_n, V_n[1] / U_n, V_n[2] / U_n};
          E = U_n * U_n;
          sbt::coordsRotate(T, s.c, dir_post);
        } else {
          const NucPoint p = nucPoint(X, S.u[g], E, nuc0);
          double XS = nucRow(p, 3);
          XS = XS * rngGet(rng);
          int which = -1;
#pragma unroll 1
          for (int k = 0; k < N.nMT; ++k) {
            const CeMtRec& m = ctx.ce.mt[N.mtFirst + k];
            const int idxT = p.idx - m.firstIdx + 1;
            if (idxT < 1) continue;
            const double topXS = tp(m.xsPos + idxT), bottomXS = tp(m.xsPos + idxT - 1);
            XS = XS - topXS * p.f - (1.0 - p.f) * bottomXS;
            if (XS <= 0.0) { which = k; break; }
          }
          if (which < 0) { atomicMax(&a.cd->error, SB_ERR_SAMPLING); which = 0; }
          const CeMtRec& m = ctx.ce.mt[N.mtFirst + which];
          MTout = m.MT;
          double mu = sbk::tapeSampleMu(tp, m.angPos, N.andPos, E, rng, &kerr);
          double E_o = sbk::tapeSampleEnergy(tp, m.lawPos, N.dlwPos, E, rng, &kerr);
          E_o = fmax(E_o, sbk::MIN_E);
          double phi = rngGet(rng) * sbk::TWO_PI;
          if (m.cmFrame) {
            double E_out = E;
            asymptoticInelasticScatter(E_out, mu, E_o, N.awr);
            E = E_out;
          } else E = E_o;
          double d[3] = {s.c.u[0][0], s.c.u[0][1], s.c.u[0][2]};
          ceRotate(d, mu, phi);
          sbt::coordsRotate(T, s.c, d);
          double rel = (double)m.TY;
          if (m.relPos) rel = sbk::tapeTableAtNI(tp, m.relPos, E, &kerr, nullptr);
          S.w[g] = wPre * rel;
          if (a.impScores) {
            double score = 0.0;
            if (MTout == 16 || MTout == 11 || MTout == 24 || MTout == 30 || MTout == 41 || (MTout >= 875 && MTout <= 891)) score = 1.0 * wPre;
            else if (MTout == 17 || MTout == 25 || MTout == 42) score = 2.0 * wPre;
            else if (MTout == 37) score = 3.0 * wPre;
            if (score > 0.0) S.sScat[g] += score;
          }
        }
        if (kerr) atomicMax(&a.cd->error, SB_ERR_CE_DATA);
        S.rng[g] = rng; S.E[g] = E;
        bool died = E < ctx.ce.minE;
        if (!died) {
          S.mode[g] = 0;
          int u = sbce::unionSearch(X, E);
          if (u == 0) { atomicMax(&a.cd->error, SB_ERR_CE_ENERGY); died = true; }
          else { S.u[g] = u; S.majXS[g] = fmax(majorantAt(X, u, E) + 0.0, collisionXS); }
        }
        if (died) slotDie(a, S, g, 0.0);
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    nSeg += __shfl_down_sync(FULL, nSeg, d); nColl += __shfl_down_sync(FULL, nColl, d); nScore += __shfl_down_sync(FULL, nScore, d);
    nTerms += __shfl_down_sync(FULL, nTerms, d);
  }
  if (lane == 0) {
    atomicAdd(&a.cd->nSeg, (unsigned long long)nSeg); atomicAdd(&a.cd->nColl, (unsigned long long)nColl);
    atomicAdd(&a.cd->nScore, (unsigned long long)nScore); atomicAdd(&a.cd->nXsTerms, (unsigned long long)nTerms);
  }
}

}  // namespace sbc
