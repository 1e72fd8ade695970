// The continuous-energy history kernel: eigenvalue cycles with `dataType ce` (delta, surface and hybrid tracking).
//
// Same event rounds, bank handling and cycle close as the multigroup kernels (sb_hist.cuh, sb_track.cuh); what differs:
//   * the particle carries E; cross sections come from the unionised-grid lookup of sb_ce.cuh (one hashed search per
//     energy CHANGE, then one index-table row per lookup);
//   * the collision is neutronCEstd: nuclide sampling, channel inversion on the micro set, implicit fission sites from the
//     nuclide's fission reaction, elastic (target-at-rest / free gas) and inelastic (MT inversion, CM or LAB laws) scattering
//     sampled from the ACE tape (sb_cekin.cuh), energy cut-off;
//   * tallies: energyMap bins on E; macroscopic responses and the implicit k-eff scores from the full macro set at (mat, E).
//
//   TransportOperator/transportOperator{DT,ST,HT}_class.f90                                 tracking loops (as sb_track.cuh)
//   NuclearData/ceNeutronData/ceNeutronDatabase_inter.f90:120-230                           getTrackingXS / getTrackMatXS / getMajorantXS
//   NuclearData/ceNeutronData/ceNeutronMaterial_class.f90:338-455                           sampleNuclide / sampleFission
//   CollisionOperator/CollisionProcessors/collisionProcessor_inter.f90:114-195              collide
//   CollisionOperator/CollisionProcessors/neutronCEstd_class.f90:157-588                    the CE hooks
//   CollisionOperator/scatteringKernels_func.f90:38-378                                     asymptotic + free-gas kernels
//   Tallies/TallyClerks/{collisionClerk,keffImplicitClerk}_class.f90, TallyMaps/Maps1D/energyMap_class.f90
//   ParticleObjects/Source/fissionSource_class.f90:149-271                                  CE source sites
// Not on the device (refused by the host at load time): S(a,b), URR tables, TMS, DBRC, correlated laws.
#pragma once
#include "sb_ce.cuh"
#include "sb_cekin.cuh"
#include "sb_track.cuh"

namespace sbc {
using namespace sbd;
#ifndef SB_CE_NUC_UNROLL
#define SB_CE_NUC_UNROLL 1
#endif
constexpr int NUC_UNROLL = SB_CE_NUC_UNROLL;      // nuclide loops: rolled keeps the code small, unrolled overlaps the gathers of consecutive nuclides

using sbh::rngGet;
using sbk::CeMtRec; using sbk::CeNucRec; using sbk::Tape;

struct CeModelDev {
  sbce::CeDev xs;
  const double* tape; const CeNucRec* nuc; const CeMtRec* mt;
  double minE, maxE, threshE, threshA, sourceE, eLo, eHi;              // neutronCEstd settings, fissionSource E, energyBounds
};

struct CeArgs {
  Model M; const char* blob; CeModelDev ce;
  const ulonglong2* seedTab;
  int n; sbh::Bank in; sbh::Bank out; int cap;
  int* nsites; double *hProd, *hAbs, *hLeak, *hScat;
  double* bins; int phase; int needMacro; int impScores;
  uint64_t rng0; int histOffset; double k_eff;
  sbh::CycleDev* cd;
  int tracking; double htCutoff; int stCache;
  sbt::SecStack stk;
  int nTrackClerks;
};

// what the out-of-line device functions need, kept once per CTA in shared memory: a reference to kernel parameters would make
// every thread copy them to its stack (params live in the constant bank and have no address)
struct CeCtx { Model M; Tables T; CeModelDev ce; double* bins; int phase, needMacro, impScores; };

// ---- cross sections at (E, union interval u) ---------------------------------------------------------------------------
struct NucPoint { int idx; double f; const double* d; int rows; };
__device__ __forceinline__ NucPoint nucPoint(const sbce::CeDev& c, int u, double e, int nuc0) {      // nuclide%search through the index table
  NucPoint p;
  p.idx = sbce::nucIndex(c, u, e, nuc0);
  const double* g = c.grid + __ldg(c.gridOff + nuc0) + (p.idx - 1);
  const double E_low = __ldg(g), E_top = __ldg(g + 1);
  p.f = (e - E_low) / (E_top - E_low);
  p.rows = __ldg(c.rows + nuc0);
  p.d = c.data + __ldg(c.dataOff + nuc0) + (size_t)(p.idx - 1) * p.rows;
  return p;
}
__device__ __forceinline__ double nucRow(const NucPoint& p, int row) {                                // mainData(row, idx+1)*f + (1-f)*mainData(row, idx)
  return __ldg(p.d + p.rows + row - 1) * p.f + (1.0 - p.f) * __ldg(p.d + row - 1);
}
__device__ inline void nucMicro(const NucPoint& p, double xs[8]) {                                    // aceNeutronNuclide%microXSs
#pragma unroll
  for (int r = 0; r < 8; ++r) xs[r] = (r < p.rows) ? nucRow(p, r + 1) : 0.0;
}
__device__ __noinline__ void matMacro(const sbce::CeDev& c, int u, double e, int m, double xs[8]) {         // updateMacroXSs + neutronMacroXSs%add
  const int k0 = __ldg(c.matOff + m - 1), k1 = __ldg(c.matOff + m);
#pragma unroll
  for (int r = 0; r < 8; ++r) xs[r] = 0.0;
#pragma unroll 1
  for (int k = k0; k < k1; ++k) {
    const NucPoint p = nucPoint(c, u, e, __ldg(c.matNuc + k) - 1);
    const double dens = __ldg(c.matDens + k) * 1.0;
    double mic[8]; nucMicro(p, mic);
#pragma unroll
    for (int r = 0; r < 8; ++r) xs[r] = xs[r] + dens * mic[r];
  }
}
__device__ __forceinline__ double majorantAt(const sbce::CeDev& c, int u, double e) {                 // updateMajorantXS
  const int uu = u > c.nUnion - 1 ? c.nUnion - 1 : u;
  const double E_low = __ldg(c.uGrid + uu - 1), E_top = __ldg(c.uGrid + uu);
  const double f = (e - E_low) / (E_top - E_low);
  return __ldg(c.uMaj + uu) * f + (1.0 - f) * __ldg(c.uMaj + uu - 1);
}
// neutronMacroXSs%get (neutronXsPackages_class.f90:143-190) on the 8-vector { total, el, inel, capture, fission, nuFission, kappa, promptNu }
__device__ inline double ceResponse(const double x[8], int MT) {
  switch (MT) {
    case -1: return x[0];
    case -2: return x[3];
    case -3: return x[1];
    case -22: return x[2] + x[4] + x[3];
    case -4: return x[2];
    case -20: return x[1] + x[2];
    case -6: return x[4];
    case -7: return x[5];
    case -80: return x[6];
    case -8: return x[7];
    case -9: return x[5] - x[7];
    case -21: return x[4] + x[3];
    default: return 0.0;
  }
}
// multiMap over spaceMap / materialMap / energyMap (energyMap_class.f90: CE particles are binned on E)
__device__ inline int clerkBinCE(const DClerk& c, const char* blob, const double r[3], int mat, double E) {
  int idx = 1;
  for (int i = 0; i < c.nMaps; ++i) {
    int b;
    if (c.mapType[i] == SB_MAP_SPACE)
      b = gridSearch(c.mapGrid[i], c.mapFirst[i], c.mapStep[i], c.mapN[i], (const double*)(blob + c.mapOff[i]), r[c.mapAxis[i]]);
    else if (c.mapType[i] == SB_MAP_MATERIAL) {
      const int* mb = (const int*)(blob + c.mapOff[i]);
      b = (mat >= 1 && mat <= c.mapGrid[i]) ? mb[mat - 1] : c.mapDef[i];
    } else if (c.mapGrid[i] == SB_GRID_LOG) {                                                          // grid_class.f90:154-176, logarithmic
      b = (int)floor(sbk::kLog(E / c.mapFirst[i]) / c.mapStep[i]) + 1;
      if (b < 1 || b >= c.mapN[i] + 1) b = 0;
    } else b = gridSearch(c.mapGrid[i], c.mapFirst[i], c.mapStep[i], c.mapN[i], (const double*)(blob + c.mapOff[i]), E);
    if (b == 0) return 0;
    idx = idx + (b - 1) * c.mapMul[i];
  }
  return idx;
}
// tallyAdmin%reportInColl for a CE particle
__device__ __noinline__ void scoreInCollCE(const CeCtx& a, const char* base, const double r[3], int mat, double E, int u,
                                     double w, double trackXS, double sigmaTot, bool virt, double& sProd, double& sAbs, unsigned& nScore) {
  const bool isVoid = (mat == SB_VOID_MAT);
  const int nC = a.M.nClerk[a.phase];
  if (nC == 0 && !a.impScores) return;
  double x[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (!isVoid && (a.needMacro || a.impScores)) matMacro(a.ce.xs, u, E, mat, x);
  const double flux = w / trackXS;
  const DClerk* cl = (const DClerk*)(base + a.M.oClerk[a.phase]);
  for (int c = 0; c < nC; ++c) {
    const DClerk& k = cl[c];
    if (k.kind != SB_CLERK_COLLISION) continue;
    if (!k.handleVirtual && (virt || isVoid)) continue;
    int bin = clerkBinCE(k, base, r, mat, E);
    if (bin == 0) continue;
    double f = k.handleVirtual ? flux : w / (sigmaTot + 0.0);
    int addr = k.addr + k.nResp * (bin - 1) - 1;
    for (int i = 0; i < k.nResp; ++i) {
      double resp = (k.respMT[i] == 0) ? 1.0 : (isVoid ? 0.0 : ceResponse(x, k.respMT[i]));
      double s = resp * f;
      if (s != 0.0) { binAdd(a.bins + addr + i, s); ++nScore; }
    }
  }
  if (a.impScores && !isVoid) {                                                                        // keffImplicitClerk%reportInColl
    sProd += x[5] * flux;
    sAbs += (x[3] + x[4]) * flux;
    nScore += 2;
  }
}

// sbce::matTotal with the loop kept rolled (code size; the lookup kernel keeps the unrolled one)
__device__ __noinline__ double ceMatTotal(const sbce::CeDev& c, int u, double e, int m, unsigned& terms) {
  const int k0 = __ldg(c.matOff + m - 1), k1 = __ldg(c.matOff + m);
  terms += (unsigned)(k1 - k0);
  double tot = 0.0;
#pragma unroll NUC_UNROLL
  for (int k = k0; k < k1; ++k) {
    const int nuc = __ldg(c.matNuc + k) - 1;
    const int idx = sbce::nucIndex(c, u, e, nuc);
    double E_low, E_top, s_low, s_top;
    sbce::ldPair(c.pairTot + 4 * (__ldg(c.pairOff + nuc) + (idx - 1)), E_low, E_top, s_low, s_top);
    const double f = (e - E_low) / (E_top - E_low);
    tot = tot + __ldg(c.matDens + k) * (s_top * f + (1.0 - f) * s_low);
  }
  return tot * 1.0;
}
__device__ __noinline__ void ceRotate(double d[3], double mu, double phi) { rotateVector(d, mu, phi); }

// tallyAdmin%reportPath -> trackClerk%reportPath for a CE particle (pre-path position and material, current energy)
__device__ __noinline__ void scorePathCE(const CeCtx& a, const char* base, const double rPre[3], int matPre, double E, int u, double w, double L, unsigned& nScore) {
  const bool isVoid = (matPre == SB_VOID_MAT);
  double x[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (!isVoid && a.needMacro) matMacro(a.ce.xs, u, E, matPre, x);
  const int nC = a.M.nClerk[a.phase];
  const DClerk* cl = (const DClerk*)(base + a.M.oClerk[a.phase]);
  for (int c = 0; c < nC; ++c) {
    const DClerk& k = cl[c];
    if (k.kind != SB_CLERK_TRACK) continue;
    int bin = clerkBinCE(k, base, rPre, matPre, E);
    if (bin == 0) continue;
    int addr = k.addr + k.nResp * (bin - 1) - 1;
    for (int i = 0; i < k.nResp; ++i) {
      double resp = (k.respMT[i] == 0) ? 1.0 : (isVoid ? 0.0 : ceResponse(x, k.respMT[i]));
      double s = resp * w * L;
      if (s != 0.0) { binAdd(a.bins + addr + i, s); ++nScore; }
    }
  }
}

// ---- scattering kernels (scatteringKernels_func.f90) ---------------------------------------------------------------------
__device__ __forceinline__ void asymptoticScatter(double& E, double& mu, double A) {
  const double E_in = E, inv_Ap1 = 1.0 / (A + 1.0);
  E = (1.0 + A * A + 2 * A * mu) * E_in * inv_Ap1 * inv_Ap1;
  mu = (A * mu + 1) * sqrt(E_in / E) * inv_Ap1;
  if (mu > 1.0) mu = 1.0;
}
__device__ __forceinline__ void asymptoticInelasticScatter(double& E, double& mu, double E_out, double A) {
  const double E_in = E, inv_Ap1 = 1.0 / (A + 1.0);
  E = E_out + (E_in + 2.0 * mu * (A + 1.0) * sqrt(E_in * E_out)) * inv_Ap1 * inv_Ap1;
  mu = mu * sqrt(E_out / E) + sqrt(E_in / E) * inv_Ap1;
  if (mu > 1.0) mu = 1.0;
}
// targetVelocity_constXS: returns X (speed in units of sqrt(kT/A)) and mu of the target; the rejection loop is the reference's
__device__ __noinline__ void sampleTargetVelocity(double Y, uint64_t& rng, double& X, double& mu) {
  const double alpha = 2.0 / (Y * sbk::SQRT_PI + 2.0);
  for (;;) {
    const double r1 = rngGet(rng), r2 = rngGet(rng), r3 = rngGet(rng);
    if (r1 > alpha) {                                                                                  // sample_x2expx2
      const double q1 = rngGet(rng), q2 = rngGet(rng), q3 = rngGet(rng);
      double s, c; sbk::kSinCos(0.5 * sbk::PI * q1, &s, &c);
      const double beta = c * c;
      const double gamma05 = -sbk::kLog(q2) * beta;
      X = sqrt(-sbk::kLog(q3) + gamma05);
    } else {                                                                                           // sample_x3expx2
      const double q1 = rngGet(rng), q2 = rngGet(rng);
      X = sqrt(-sbk::kLog(q1) - sbk::kLog(q2));
    }
    mu = 2.0 * r2 - 1.0;
    const double rel_v = sqrt(Y * Y + X * X - 2.0 * X * Y * mu);
    const double P_acc = rel_v / (Y + X);
    if (P_acc > r3) return;
  }
}

// ---- fissionSource, CE branch ----------------------------------------------------------------------------------------------
__global__ void k_source_ce(const Model M, const char* blob, const CeModelDev ce, sbh::Bank out, int n, uint64_t rng0, int offset,
                            double b0, double b1, double b2, double t0, double t1, double t2, sbh::CycleDev* cd) {
  const Tables T = bind(M, blob);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint64_t rng = rng_skip(rng0, RNG_STRIDE * (int64_t)(offset + i + 1));
    const double bottom[3] = {b0, b1, b2}, top[3] = {t0, t1, t2};
    bool ok = false;
    const double E = ce.sourceE;
    const int u = sbce::unionSearch(ce.xs, E);
    if (u == 0) { atomicMax(&cd->error, SB_ERR_CE_ENERGY); return; }
    for (int att = 0; att < 10000 && !ok; ++att) {
      double r3[3]; r3[0] = rngGet(rng); r3[1] = rngGet(rng); r3[2] = rngGet(rng);
      double r[3], uu[3] = {1.0, 0.0, 0.0};
      for (int k = 0; k < 3; ++k) r[k] = (top[k] - bottom[k]) * r3[k] + bottom[k];
      int mat, uid;
      geomPlace(M, T, r, uu, mat, uid);
      if (mat == SB_VOID_MAT || mat == SB_OUTSIDE_MAT) continue;
      if (mat == SB_UNDEF_MAT) { atomicMax(&cd->error, SB_ERR_UNDEF_MAT); break; }
      if (mat == SB_OVERLAP_MAT) { atomicMax(&cd->error, SB_ERR_OVERLAP_MAT); break; }
      if (!T.fissile[mat - 1]) continue;
      // ceNeutronMaterial%sampleFission
      double x[8]; matMacro(ce.xs, u, E, mat, x);
      double xs = x[5] * rngGet(rng);
      const int k0 = ce.xs.matOff[mat - 1], k1 = ce.xs.matOff[mat];
      int nuc0 = -1;
      for (int k = k0; k < k1; ++k) {
        const int nn = ce.xs.matNuc[k] - 1;
        const NucPoint p = nucPoint(ce.xs, u, E, nn);
        const double nuf = (p.rows == 8) ? nucRow(p, 6) : 0.0;
        xs = xs - nuf * ce.xs.matDens[k] * 1.0 * 1.0;
        if (xs < 0.0) { nuc0 = nn; break; }
      }
      if (nuc0 < 0) { atomicMax(&cd->error, SB_ERR_SAMPLING); break; }
      const CeNucRec& N = ce.nuc[nuc0];
      const Tape tp{ce.tape, N.base};
      double mu, phi, E_out; int kerr = 0;
      sbk::tapeSampleFission(tp, N, E, rng, mu, phi, E_out, &kerr);
      if (kerr) { atomicMax(&cd->error, SB_ERR_CE_DATA); break; }
      double d[3] = {1.0, 0.0, 0.0};
      ceRotate(d, mu, phi);
      if (E_out > ce.eHi) E_out = ce.eHi;
      out.rx[i] = r[0]; out.ry[i] = r[1]; out.rz[i] = r[2];
      out.ux[i] = d[0]; out.uy[i] = d[1]; out.uz[i] = d[2];
      out.w[i] = 1.0; out.G[i] = 0; out.E[i] = E_out; out.brood[i] = 0; out.seq[i] = 0;
      ok = true;
    }
    if (!ok) atomicMax(&cd->error, SB_ERR_SOURCE);
  }
}

// ---- the kernel -------------------------------------------------------------------------------------------------------------
// SYNC: the warps of a CTA run the event phases (refill | flight | collision 1 | collision 2) in lockstep, separated by CTA
// barriers.  The loop body is far larger than the instruction caches (L0 6 KB, L1.5 32 KB per SM); without the barriers
// every warp is at a different place of it and the SM stalls on instruction fetch (ncu r01f: 85 % of the stall samples are
// "no instruction", I-cache hit rate 43 %); in lockstep one fetch serves all the warps of the CTA.
template <int THREADS, int BPS, bool SYNC>
__global__ void __launch_bounds__(THREADS, BPS) k_histories_ce(const CeArgs a) {
  const char* base = a.blob;
  __shared__ CeCtx s_ctx;
  if (threadIdx.x == 0) { s_ctx.M = a.M; s_ctx.T = bind(a.M, base); s_ctx.ce = a.ce; s_ctx.bins = a.bins; s_ctx.phase = a.phase; s_ctx.needMacro = a.needMacro; s_ctx.impScores = a.impScores; }
  __syncthreads();
  const CeCtx& ctx = s_ctx;
  const Tables& T = s_ctx.T;
  const Model& M = s_ctx.M;
  const sbce::CeDev& X = s_ctx.ce.xs;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned ltMask = (1u << lane) - 1u;
  const double collisionXS = M.collisionXS;

  sbt::Coords c; sbt::DistCache cache; cache.lvl = 0;
  bool alive = false, exhausted = false;
  const int gLane = blockIdx.x * blockDim.x + threadIdx.x, nLanes = gridDim.x * blockDim.x;
  int nStk = 0;                                                 // fixed source: entries in this history's private buffer
  int hi = -1, nSite = 0, hSeg = 0, mode = 0, u = 1;            // mode: 0 = transport call begins, 1 = delta, 2 = surface; u = union interval of E
  double E = 1.0, w = 0.0, w0 = 0.0, trackXS = 1.0, majXS = 1.0, sigTot = 0.0;
  uint64_t rng = 0;
  double sProd = 0.0, sAbs = 0.0, sScat = 0.0, sLeak = 0.0;     // sLeak: leaked weight of the history (with its secondaries in a fixed-source run)
  unsigned nSeg = 0, nColl = 0, nScore = 0, nTerms = 0;
  c.nesting = 1; c.mat = SB_UNDEF_MAT; c.uid = -3;

  for (;;) {
    // ---------------- refill dead lanes ---------------------------------------------------------------
    {
      unsigned need = __ballot_sync(FULL, !alive);
      if (need != 0u && !exhausted) {
        int cnt = __popc(need);
        int b = 0;
        if (lane == 0) b = atomicAdd(&a.cd->nextHistory, cnt);
        b = __shfl_sync(FULL, b, 0);
        if (b + cnt >= a.n) exhausted = true;
        int my = b + __popc(need & ltMask);
        if (!alive && my < a.n) {
          hi = my;
          c.r[0][0] = a.in.rx[hi]; c.r[0][1] = a.in.ry[hi]; c.r[0][2] = a.in.rz[hi];
          c.u[0][0] = a.in.ux[hi]; c.u[0][1] = a.in.uy[hi]; c.u[0][2] = a.in.uz[hi];
          w = a.in.w[hi]; w0 = w; E = a.in.E[hi];
          rng = sbh::rngSeed(a.seedTab, a.rng0, (unsigned)(a.histOffset + hi + 1));
          if (!sbt::placeCoord(M, T, c)) atomicMax(&a.cd->error, SB_ERR_NEST);
          nSite = 0; hSeg = 0; sProd = 0.0; sAbs = 0.0; sScat = 0.0; sLeak = 0.0; mode = 0;
          alive = true;
          u = sbce::unionSearch(X, E);
          if (u == 0) { atomicMax(&a.cd->error, SB_ERR_CE_ENERGY); u = 1; }
          majXS = fmax(majorantAt(X, u, E) + 0.0, collisionXS);
        }
        need = __ballot_sync(FULL, !alive);
      }
      const bool warpDone = (need == FULL && exhausted);
      if (SYNC) { if (__syncthreads_and(warpDone ? 1 : 0)) break; }
      else if (warpDone) break;
    }

    // ---------------- event: one flight segment -------------------------------------------------------
    bool realColl = false, died = false, scoreVirt = false;
    double leak = 0.0;
    if (alive) {
      if (mode == 0) {                                      // transportOperator%transport begins
        if (a.tracking == SB_TRACK_DT) mode = 1;
        else if (a.tracking == SB_TRACK_ST) mode = 2;
        else {                                              // transportOperatorHT_class.f90:49-81
          double majorant_inv = 1.0 / majXS;
          double sigmaT = (c.mat == SB_VOID_MAT) ? 0.0 : ceMatTotal(X, u, E, c.mat, nTerms) + 0.0;
          double ratio = sigmaT * majorant_inv;
          mode = (ratio > (1.0 - a.htCutoff)) ? 1 : 2;
        }
        cache.lvl = 0;
      }
      if (mode == 1) {                                      // deltaTracking, one tentative flight
        trackXS = majXS;
        double majorant_inv = 1.0 / trackXS;
        double distance = -sbk::kLog(rngGet(rng)) * majorant_inv;
        sbt::geomTeleportCoords(M, T, c, distance);
        ++nSeg; ++hSeg;
        if (c.mat == SB_OUTSIDE_MAT) { leak = w; sLeak = sLeak + w; died = true; }
        else if (c.mat >= SB_OVERLAP_MAT && c.mat != SB_VOID_MAT) { atomicMax(&a.cd->error, c.mat == SB_UNDEF_MAT ? SB_ERR_UNDEF_MAT : SB_ERR_OVERLAP_MAT); died = true; }
        else {
          bool virt = true;
          sigTot = 0.0;
          if (c.mat != SB_VOID_MAT) {
            sigTot = ceMatTotal(X, u, E, c.mat, nTerms);
            if (rngGet(rng) < (sigTot + 0.0) * majorant_inv) { realColl = true; virt = false; }
          }
          scoreVirt = virt;
        }
      } else {                                              // surfaceTracking, one segment
        const double tol = 1.0E-12;
        int m = c.mat;
        sigTot = (m == SB_VOID_MAT) ? 0.0 : ceMatTotal(X, u, E, m, nTerms);
        double sigmaTrack = (m == SB_VOID_MAT) ? collisionXS : fmax(sigTot + 0.0, collisionXS);
        trackXS = sigmaTrack;
        double dist, invSigmaTrack, sigmaT;
        if (sigmaTrack < tol) { dist = INF; invSigmaTrack = INF; sigmaT = 0.0; }
        else {
          invSigmaTrack = 1.0 / sigmaTrack;
          dist = -sbk::kLog(rngGet(rng)) * invSigmaTrack;
          sigmaT = sigTot + 0.0;
        }
        int event;
        const double rPre[3] = {c.r[0][0], c.r[0][1], c.r[0][2]};      // p%savePrePath
        sbt::geomMove(M, T, c, dist, event, a.stCache ? &cache : nullptr);
        ++nSeg; ++hSeg;
        if (a.nTrackClerks) scorePathCE(ctx, base, rPre, m, E, u, w, dist, nScore);      // tally%reportPath(p, dist)
        m = c.mat;
        if (m == SB_OUTSIDE_MAT) { leak = w; sLeak = sLeak + w; died = true; }
        else if (m >= SB_OVERLAP_MAT && m != SB_VOID_MAT) { atomicMax(&a.cd->error, m == SB_UNDEF_MAT ? SB_ERR_UNDEF_MAT : SB_ERR_OVERLAP_MAT); died = true; }
        else if (event == sbt::COLL_EV) {
          if (rngGet(rng) < sigmaT * invSigmaTrack) realColl = true;
          else scoreVirt = true;
        }
      }
    }
    if (SYNC) __syncthreads();
    // ---------------- tallyAdmin%reportInColl of the virtual collisions ------------------------------------
    if (scoreVirt) scoreInCollCE(ctx, base, c.r[0], c.mat, E, u, w, trackXS, sigTot, true, sProd, sAbs, nScore);

    if (SYNC) __syncthreads();
    // ---------------- event: collision, part 1: nuclide, channel, number of fission sites ---------------------
    int MT = 0, nNew = 0, nuc0 = 0;
    const int mat = c.mat;
    double mic[8];
    if (realColl) {
      (void)rngGet(rng);                                    // alpha-absorption test always draws (probAlpha = 0)
      // ceNeutronMaterial%sampleNuclide
      double rem = (sigTot * 1.0) * rngGet(rng);
      const int k0 = __ldg(X.matOff + mat - 1), k1 = __ldg(X.matOff + mat);
      nuc0 = -1;
#pragma unroll NUC_UNROLL
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
      nucMicro(p, mic);
      const double rr = rngGet(rng);
      {                                                     // neutronMicroXSs%invert
        int C = 1;
        double xs = mic[0] * rr - mic[1];
        if (xs > 0.0) C += 1;
        xs = xs - mic[2];
        if (xs > 0.0) C += 1;
        xs = xs - mic[3];
        if (xs > 0.0) C += 1;
        MT = C;                                             // 1 elastic, 2 inelastic, 3 capture (N_disap), 4 fission
      }
      ++nColl;
    }
    if (SYNC) __syncthreads();
    if (realColl) {
      // tallyAdmin%reportInColl(p, virtual = .false.) comes after sampleCollision (collisionProcessor_inter.f90:131)
      scoreInCollCE(ctx, base, c.r[0], mat, E, u, w, trackXS, sigTot, false, sProd, sAbs, nScore);
      if (ctx.ce.nuc[nuc0].fissile) {                         // neutronCEstd implicit (:217-300)
        double rand1 = rngGet(rng);
        nNew = (int)(fabs((w * mic[5]) / (w0 * mic[0] * a.k_eff)) + rand1);
        if (nNew < 0) nNew = 0;
      }
    }
    // ---------------- warp-aggregated allocation of fission-bank slots --------------------------------
    int slot = -1;
    {
      unsigned spawn = __ballot_sync(FULL, nNew > 0 && !a.stk.on);
      if (spawn) {
        int inc = nNew;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += t; }
        int total = __shfl_sync(FULL, inc, 31);
        int b = 0;
        if (lane == 0) b = atomicAdd(&a.cd->nSites, total);
        b = __shfl_sync(FULL, b, 0);
        slot = b + inc - nNew;
        if (b + total > a.cap) { atomicMax(&a.cd->error, SB_ERR_BANK_OVERFLOW); slot = -1; }
      }
    }
    if (SYNC) __syncthreads();
    // ---------------- collision, part 2: fission sites, then the channel ------------------------------------------
    if (realColl) {
      const CeNucRec& N = ctx.ce.nuc[nuc0];
      const Tape tp{ctx.ce.tape, N.base};
      int kerr = 0;
      const double wSite = fsign(w0, w);
#pragma unroll 1
      for (int i = 0; i < nNew; ++i) {
        double mu, phi, E_out;
        sbk::tapeSampleFission(tp, N, E, rng, mu, phi, E_out, &kerr);
        double d[3] = {c.u[0][0], c.u[0][1], c.u[0][2]};
        ceRotate(d, mu, phi);
        if (E_out > ctx.ce.maxE) E_out = ctx.ce.maxE;
        if (a.stk.on) {                                     // fixed source: a secondary of this history
          if (nStk >= a.stk.cap) atomicMax(&a.cd->error, SB_ERR_BANK_OVERFLOW);
          else { a.stk.push(nStk, gLane, nLanes, c.r[0], d, wSite * 1.0, E_out, 0); ++nStk; }
        } else if (slot >= 0) {
          int s = slot + i;
          a.out.rx[s] = c.r[0][0]; a.out.ry[s] = c.r[0][1]; a.out.rz[s] = c.r[0][2];
          a.out.ux[s] = d[0]; a.out.uy[s] = d[1]; a.out.uz[s] = d[2];
          a.out.w[s] = wSite * 1.0; a.out.G[s] = 0; a.out.E[s] = E_out; a.out.brood[s] = hi; a.out.seq[s] = nSite + i;
        }
      }
      if (kerr) atomicMax(&a.cd->error, SB_ERR_CE_DATA);
      if (!a.stk.on) nSite += nNew;
    }
    if (SYNC) __syncthreads();
    if (realColl) {
      const CeNucRec& N = ctx.ce.nuc[nuc0];
      const Tape tp{ctx.ce.tape, N.base};
      int kerr = 0;
      const double wPre = w;
      int MTout = 0;
      if (MT == 1) {                                        // elastic (:330-375)
        const double A = N.awr, kT = N.kT;
        const bool isFixed = (E > kT * ctx.ce.threshE) && (A > ctx.ce.threshA);
        if (isFixed) {                                      // scatterFromFixed
          double mu = sbk::tapeSampleMu(tp, N.elAng, N.andPos, E, rng, &kerr);
          double phi = rngGet(rng) * sbk::TWO_PI;
          double E_out = E;
          asymptoticScatter(E_out, mu, A);
          double d[3] = {c.u[0][0], c.u[0][1], c.u[0][2]};
          ceRotate(d, mu, phi);
          sbt::coordsRotate(T, c, d);
          E = E_out;
        } else {                                            // scatterFromMoving (:482-588), constant-XS free gas
          const double dir_pre[3] = {c.u[0][0], c.u[0][1], c.u[0][2]};
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
          sbt::coordsRotate(T, c, dir_post);                // p%point(dir_post)
        }
        MTout = 2;
      } else if (MT == 2) {                                 // inelastic (:377-405)
        const NucPoint p = nucPoint(X, u, E, nuc0);        // aceNeutronNuclide%invertInelastic
        double XS = nucRow(p, 3);
        XS = XS * rngGet(rng);
        int which = -1;
#pragma unroll 1
        for (int i = 0; i < N.nMT; ++i) {
          const CeMtRec& m = ctx.ce.mt[N.mtFirst + i];
          const int idxT = p.idx - m.firstIdx + 1;
          if (idxT < 1) continue;
          const double topXS = tp(m.xsPos + idxT), bottomXS = tp(m.xsPos + idxT - 1);
          XS = XS - topXS * p.f - (1.0 - p.f) * bottomXS;
          if (XS <= 0.0) { which = i; break; }
        }
        if (which < 0) { atomicMax(&a.cd->error, SB_ERR_SAMPLING); which = 0; }
        const CeMtRec& m = ctx.ce.mt[N.mtFirst + which];
        MTout = m.MT;
        // neutronScatter%sampleOut
        double mu = sbk::tapeSampleMu(tp, m.angPos, N.andPos, E, rng, &kerr);
        double E_o = sbk::tapeSampleEnergy(tp, m.lawPos, N.dlwPos, E, rng, &kerr);
        E_o = fmax(E_o, sbk::MIN_E);
        double phi = rngGet(rng) * sbk::TWO_PI;
        if (m.cmFrame) {                                    // scatterFromFixed
          double E_out = E;
          asymptoticInelasticScatter(E_out, mu, E_o, N.awr);
          E = E_out;
        } else E = E_o;                                     // scatterInLAB
        double d[3] = {c.u[0][0], c.u[0][1], c.u[0][2]};
        ceRotate(d, mu, phi);
        sbt::coordsRotate(T, c, d);
        double rel = (double)m.TY;
        if (m.relPos) rel = sbk::tapeTableAtNI(tp, m.relPos, E, &kerr, nullptr);
        w = w * rel;                                        // p%w * reac%release(p%E), at the outgoing energy as the reference
      } else died = true;                                   // capture / fission
      if (E < ctx.ce.minE) died = true;                       // cutoffs
      if (kerr) atomicMax(&a.cd->error, SB_ERR_CE_DATA);
      // keffImplicitClerk%reportOutColl: (n,xn) multiplicities by MT (keffImplicitClerk_class.f90:245-270)
      if (a.impScores && MT == 2) {
        double score = 0.0;
        if (MTout == 16 || MTout == 11 || MTout == 24 || MTout == 30 || MTout == 41 || (MTout >= 875 && MTout <= 891)) score = 1.0 * wPre;
        else if (MTout == 17 || MTout == 25 || MTout == 42) score = 2.0 * wPre;
        else if (MTout == 37) score = 3.0 * wPre;
        if (score > 0.0) sScat += score;
      }
      if (!died) {
        mode = 0;                                           // the next flight is a new transport call
        u = sbce::unionSearch(X, E);
        if (u == 0) { atomicMax(&a.cd->error, SB_ERR_CE_ENERGY); died = true; u = 1; }
        else majXS = fmax(majorantAt(X, u, E) + 0.0, collisionXS);
      }
    }
    if (died && nStk > 0) {                                  // bufferLoop: release the last particle detained and carry on
      int Gdummy;
      --nStk;
      a.stk.pop(nStk, gLane, nLanes, c.r[0], c.u[0], w, E, Gdummy);
      w0 = w;
      if (!sbt::placeCoord(M, T, c)) atomicMax(&a.cd->error, SB_ERR_NEST);
      mode = 0; cache.lvl = 0;
      died = false;
      u = sbce::unionSearch(X, E);
      if (u == 0) { atomicMax(&a.cd->error, SB_ERR_CE_ENERGY); u = 1; died = true; }
      else majXS = fmax(majorantAt(X, u, E) + 0.0, collisionXS);
    }
    if (died) {
      a.nsites[hi] = nSite;
      a.hProd[hi] = sProd; a.hAbs[hi] = sAbs; a.hLeak[hi] = sLeak; a.hScat[hi] = sScat;
      if (hSeg > 256) atomicMax(&a.cd->maxSeg, hSeg);
      alive = false;
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
