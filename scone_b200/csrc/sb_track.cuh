// Surface tracking and hybrid tracking on the device: the history kernel for decks whose transportOperator is
// transportOperatorST or transportOperatorHT. Same event rounds, bank handling, collision physics and tallies
// as k_histories (sb_hist.cuh); what differs is the flight event: the particle carries its full coordList
// (one (r, dir, uniIdx, uniRootID, localID) per nesting level, kept in L1-resident local memory) and moves from
// surface crossing to surface crossing.
//
//   transportOperatorST_class.f90:48-166, transportOperatorHT_class.f90:49-284   tracking loops, HT selector
//   geometryStd_class.f90:214-352     move_noCache / move_withCache (events COLL, BOUNDARY, CROSS)
//   geometryStd_class.f90:633-717     closestDist / closestDist_cache (FP_REL_TOL tie-break between levels)
//   geometryStd_class.f90:565-619     diveToMat
//   coord_class.f90:341-408           moveGlobal / moveLocal / rotate
//   rootUniverse_class.f90:145-181, pinUniverse_class.f90:175-253, latUniverse_class.f90:312-379,
//   cellUniverse_class.f90:284-437    distance / cross per universe ; simpleCell_class.f90:112-141 distance
#pragma once
#include "sb_hist.cuh"

namespace sbt {
using namespace sbd;
using sbh::rngGet;

constexpr int COLL_EV = 1, BOUNDARY_EV = 2, CROSS_EV = 3;
constexpr int PIN_MOVING_IN = -1, PIN_MOVING_OUT = -2, LAT_OUTLINE_SURF = -7;

#ifndef SB_COORD_NEST
#define SB_COORD_NEST 12              // capacity of the coordList (hardcoded max nesting of the reference, universalVariables.f90 MAX_NEST... = 12 here)
#endif
constexpr int COORD_NEST = SB_COORD_NEST;
struct Coords {                       // coordList (coord_class.f90:33-115), without the rotation matrices (read from the universe)
  double r[COORD_NEST][3], u[COORD_NEST][3];
  int uni[COORD_NEST], root[COORD_NEST], local[COORD_NEST];
  int nesting, mat, uid;
};
struct DistCache { int lvl; double dist[COORD_NEST]; int surf[COORD_NEST]; };

// The geometry procedures below are called from several places of the event loop; they are kept out of line (one copy each)
// so that the loop body stays within the instruction cache (the fully inlined CE kernel was 585 KB of SASS and stalled on
// instruction fetch for most of its cycles).
__device__ __noinline__ int uniFindCellNI(const Tables& T, int ui, const double r[3], const double u[3]) { return uniFindCell(T, ui, r, u); }
__device__ __noinline__ double surfDistanceNI(int type, const double* p, const double r[3], const double u[3]) { return surfDistance(type, p, r, u); }

// geometryStd%diveToMat from level `start` (1-based) ; levels below are (re)entered
__device__ __noinline__ bool diveToMat(const Tables& T, Coords& c, int start) {
  for (int i = start; i <= COORD_NEST; ++i) {
    int2 f = T.graph[c.root[i - 1] + c.local[i - 1] - 2];
    if (f.x >= 0) { c.mat = f.x; c.uid = f.y; return true; }
    if (i == COORD_NEST) break;
    double off[3]; uniCellOffset(T, c.uni[i - 1], c.local[i - 1], off);
    int ui = -f.x - 1;
    bool glob = T.uniIpar[ui * SB_UNI_NIPAR + 1] != 0;
    double rin[3], uin[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { rin[k] = glob ? c.r[0][k] : c.r[i - 1][k] - off[k]; uin[k] = c.u[i - 1][k]; }
    c.nesting += 1;
    uniEnter(T, ui, rin, uin, c.r[i], c.u[i]);
    c.uni[i] = ui; c.root[i] = f.y;
    c.local[i] = uniFindCellNI(T, ui, c.r[i], c.u[i]);
  }
  c.mat = SB_UNDEF_MAT; c.uid = -3;
  return false;
}
// geometryStd%placeCoord: from level-1 position and direction
__device__ __noinline__ bool placeCoord(const Model& M, const Tables& T, Coords& c) {
  c.nesting = 1; c.mat = SB_UNDEF_MAT; c.uid = -3;
  int ui = M.rootIdx - 1;
  double rin[3] = {c.r[0][0], c.r[0][1], c.r[0][2]}, uin[3] = {c.u[0][0], c.u[0][1], c.u[0][2]};
  uniEnter(T, ui, rin, uin, c.r[0], c.u[0]);
  c.uni[0] = ui; c.root[0] = 1;
  c.local[0] = uniFindCellNI(T, ui, c.r[0], c.u[0]);
  return diveToMat(T, c, 1);
}

// universe%distance at one level -> d, surface memento
__device__ __noinline__ void uniDistance(const Tables& T, int ui, const double r[3], const double u[3], int localID, double& d, int& sIdx) {
  const int type = T.uniType[ui];
  const int* ip = T.uniIpar + ui * SB_UNI_NIPAR;
  const double* dp = T.uniDpar + ui * SB_UNI_NDPAR;
  if (type == SB_UNI_ROOT) {                                  // rootUniverse_class.f90:145-160
    int s = ip[2] - 1;
    sIdx = ip[2];
    d = surfDistanceNI(T.surfType[s], T.surfPar + s * SB_SURF_NPAR, r, u);
  } else if (type == SB_UNI_PIN) {                            // pinUniverse_class.f90:175-214
    const int N = ip[2]; const double* r_sq = T.auxD + ip[3]; const double* tol = r_sq + N;
    const double rs = r[0] * r[0] + r[1] * r[1];
    const double k = r[0] * u[0] + r[1] * u[1];
    const double a = 1.0 - u[2] * u[2];
    double d_out = INF, d_in = INF;
    if (localID <= N) d_out = cylDistance(rs - r_sq[localID - 1], k, a, tol[localID - 1]);
    if (localID != 1) d_in = cylDistance(rs - r_sq[localID - 2], k, a, tol[localID - 2]);
    if (d_in < d_out) { sIdx = PIN_MOVING_IN; d = d_in; } else { sIdx = PIN_MOVING_OUT; d = d_out; }
  } else if (type == SB_UNI_LAT) {                            // latUniverse_class.f90:312-365
    if (localID == ip[5]) {
      double p[SB_SURF_NPAR] = {0.0, 0.0, 0.0, dp[21], dp[22], dp[23], SURF_TOL, 0.0};
      sIdx = LAT_OUTLINE_SURF;
      d = surfDistanceNI(SB_SURF_BOX, p, r, u);
      return;
    }
    int ijk[3]; lat_get_ijk(ijk, localID, ip + 2);
    d = INF; int ax = 1;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double r_bar = r[i] - dp[15 + i];
      r_bar = r_bar - (ijk[i] - 0.5) * dp[12 + i];
      double bound = fsign(dp[12 + i] * 0.5, u[i]);
      double test_d = (bound - r_bar) / u[i];
      if (test_d < d) { d = test_d; ax = i + 1; }
    }
    d = fmax(0.0, d);
    d = fmin(INF, d);
    sIdx = ax * 2;
    if (u[ax - 1] < 0.0) sIdx -= 1;
    sIdx = -sIdx;
  } else {                                                    // cellUniverse_class.f90:284-310 + simpleCell distance
    const int N = ip[2];
    d = INF; sIdx = 0;
    if (localID > N) return;                                  // undefined / overlapping local cell: the material check reports it
    int cidx = T.auxI[ip[3] + localID - 1] - 1;
    for (int k = T.cellOff[cidx]; k < T.cellOff[cidx + 1]; ++k) {
      int sidx = T.cellSurf[k];
      int s = (sidx < 0 ? -sidx : sidx) - 1;
      double t = surfDistanceNI(T.surfType[s], T.surfPar + s * SB_SURF_NPAR, r, u);
      if (t < d) { d = t; sIdx = s + 1; }
    }
  }
}
// universe%cross at one level
__device__ inline void uniCross(const Tables& T, int ui, double r[3], const double u[3], int& localID, int sIdx) {
  const int type = T.uniType[ui];
  if (type == SB_UNI_PIN) {                                   // pinUniverse_class.f90:216-237
    if (sIdx == PIN_MOVING_IN) localID -= 1; else localID += 1;
    return;
  }
  if (type == SB_UNI_CELL) {                                  // cellUniverse_class.f90:312-345: nudge, then search again
#pragma unroll
    for (int k = 0; k < 3; ++k) r[k] = r[k] + u[k] * NUDGE;
  }
  localID = uniFindCellNI(T, ui, r, u);
}

// geometryStd%move_noCache / move_withCache (no fields)
__device__ __noinline__ void geomMove(const Model& M, const Tables& T, Coords& c, double& maxDist, int& event, DistCache* cache) {
  double dist = INF; int surfIdx = 0, level = 0;
  for (int l = 1; l <= c.nesting; ++l) {                      // closestDist[_cache]
    double td; int ti;
    if (cache) {
      if (cache->lvl < l) { uniDistance(T, c.uni[l - 1], c.r[l - 1], c.u[l - 1], c.local[l - 1], cache->dist[l - 1], cache->surf[l - 1]); cache->lvl += 1; }
      td = cache->dist[l - 1]; ti = cache->surf[l - 1];
    } else uniDistance(T, c.uni[l - 1], c.r[l - 1], c.u[l - 1], c.local[l - 1], td, ti);
    if ((dist - td) >= dist * FP_REL_TOL) { dist = td; surfIdx = ti; level = l; }
  }
  if (maxDist < dist) {                                       // collision inside the cell: moveLocal at every level
    for (int i = 0; i < c.nesting; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k) c.r[i][k] = c.r[i][k] + maxDist * c.u[i][k];
    event = COLL_EV;
    if (cache) cache->lvl = 0;
  } else if (surfIdx == M.borderIdx && level == 1) {          // domain boundary: moveGlobal, explicitBC, placeCoord
#pragma unroll
    for (int k = 0; k < 3; ++k) c.r[0][k] = c.r[0][k] + dist * c.u[0][k];
    event = BOUNDARY_EV;
    maxDist = dist;
    if (cache) cache->lvl = 0;
    int s = M.borderIdx - 1;
    int b6[6]; for (int i = 0; i < 6; ++i) b6[i] = M.bc[i];
    surfExplicitBC(T.surfType[s], T.surfPar + s * SB_SURF_NPAR, b6, c.r[0], c.u[0]);
    placeCoord(M, T, c);
  } else {                                                    // crossing at `level`: moveLocal down to it, cross, dive
    c.nesting = level;
    for (int i = 0; i < level; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k) c.r[i][k] = c.r[i][k] + dist * c.u[i][k];
    event = CROSS_EV;
    maxDist = dist;
    if (cache) {
      for (int l = 0; l < level - 1; ++l) cache->dist[l] = cache->dist[l] - dist;
      cache->lvl = level - 1;
    }
    uniCross(T, c.uni[level - 1], c.r[level - 1], c.u[level - 1], c.local[level - 1], surfIdx);
    diveToMat(T, c, level);
  }
}
// geometryStd%teleport on a coordList
__device__ __noinline__ void geomTeleportCoords(const Model& M, const Tables& T, Coords& c, double dist) {
#pragma unroll
  for (int k = 0; k < 3; ++k) c.r[0][k] = c.r[0][k] + dist * c.u[0][k];
  placeCoord(M, T, c);
  if (c.mat == SB_OUTSIDE_MAT) {
    int s = M.borderIdx - 1;
    int b6[6]; for (int i = 0; i < 6; ++i) b6[i] = M.bc[i];
    surfTransformBC(T.surfType[s], T.surfPar + s * SB_SURF_NPAR, b6, c.r[0], c.u[0]);
    placeCoord(M, T, c);
  }
}
// coordList%rotate (coord_class.f90:386-408)
__device__ __noinline__ void coordsRotate(const Tables& T, Coords& c, const double d[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) c.u[0][k] = d[k];
  for (int i = 1; i < c.nesting; ++i) {
    const int ui = c.uni[i];
    if (T.uniIpar[ui * SB_UNI_NIPAR]) {
      const double* m = T.uniDpar + ui * SB_UNI_NDPAR + 3;
#pragma unroll
      for (int k = 0; k < 3; ++k) c.u[i][k] = m[3 * k] * c.u[i - 1][0] + m[3 * k + 1] * c.u[i - 1][1] + m[3 * k + 2] * c.u[i - 1][2];
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) c.u[i][k] = c.u[i - 1][k];
    }
  }
}

// private buffer of a history in a fixed-source calculation (fixedSourcePhysicsPackage_class.f90:198-246): last in, first out.
// Layout [field][entry][lane] so that the lanes of a warp touch consecutive words.
struct SecStack {
  double* d; int* G; int cap; int on;
  __device__ __forceinline__ size_t at(int field, int entry, int lane, int nLanes) const { return ((size_t)field * cap + entry) * nLanes + lane; }
  __device__ __forceinline__ void push(int entry, int lane, int nLanes, const double r[3], const double u[3], double w, double E, int g) const {
    d[at(0, entry, lane, nLanes)] = r[0]; d[at(1, entry, lane, nLanes)] = r[1]; d[at(2, entry, lane, nLanes)] = r[2];
    d[at(3, entry, lane, nLanes)] = u[0]; d[at(4, entry, lane, nLanes)] = u[1]; d[at(5, entry, lane, nLanes)] = u[2];
    d[at(6, entry, lane, nLanes)] = w; d[at(7, entry, lane, nLanes)] = E; G[(size_t)entry * nLanes + lane] = g;
  }
  __device__ __forceinline__ void pop(int entry, int lane, int nLanes, double r[3], double u[3], double& w, double& E, int& g) const {
    r[0] = d[at(0, entry, lane, nLanes)]; r[1] = d[at(1, entry, lane, nLanes)]; r[2] = d[at(2, entry, lane, nLanes)];
    u[0] = d[at(3, entry, lane, nLanes)]; u[1] = d[at(4, entry, lane, nLanes)]; u[2] = d[at(5, entry, lane, nLanes)];
    w = d[at(6, entry, lane, nLanes)]; E = d[at(7, entry, lane, nLanes)]; g = G[(size_t)entry * nLanes + lane];
  }
};

struct TrackArgs {
  Model M; const char* blob; int useSmem;
  const ulonglong2* seedTab;
  int n; sbh::Bank in; sbh::Bank out; int cap;
  int* nsites; double *hProd, *hAbs, *hLeak, *hScat;
  double* bins; int phase; int impScores;
  uint64_t rng0; int histOffset; double k_eff;
  sbh::CycleDev* cd;
  int tracking; double htCutoff; int stCache;
  SecStack stk;
  int nTrackClerks;                       // trackClerks of this phase (path-length scores along surface-tracking segments)
};

// collisionClerk / keffImplicitClerk scoring of one collision (virtual or real), generic tables
__device__ inline void scoreInColl(const TrackArgs& a, const Tables& T, const char* base, const double r[3], int mat, int G,
                                   double w, double trackXS, bool virt, double& sProd, double& sAbs, unsigned& nScore) {
  const bool isVoid = (mat == SB_VOID_MAT);
  const double* x = isVoid ? T.xs : mgRow(a.M, T, mat, G);
  const bool fissile = isVoid ? false : (T.fissile[mat - 1] != 0);
  const double flux = w / trackXS;
  const int nC = a.M.nClerk[a.phase];
  const DClerk* cl = (const DClerk*)(base + a.M.oClerk[a.phase]);
  for (int c = 0; c < nC; ++c) {
    const DClerk& k = cl[c];
    if (k.kind != SB_CLERK_COLLISION) continue;
    if (!k.handleVirtual && (virt || isVoid)) continue;
    int bin = clerkBin(k, base, r, mat);
    if (bin == 0) continue;
    double f = k.handleVirtual ? flux : w / (x[XS_TOTAL] + 0.0);
    int addr = k.addr + k.nResp * (bin - 1) - 1;
    for (int i = 0; i < k.nResp; ++i) {
      double resp = (k.respMT[i] == 0) ? 1.0 : (isVoid ? 0.0 : mgResponse(x, fissile, k.respMT[i]));
      double s = resp * f;
      if (s != 0.0) { binAdd(a.bins + addr + i, s); ++nScore; }
    }
  }
  if (a.impScores && !isVoid) {
    double nuf = fissile ? x[XS_NUFISSION] : 0.0, fis = fissile ? x[XS_FISSION] : 0.0;
    sProd += nuf * flux;
    sAbs += (x[XS_CAPTURE] + fis) * flux;
    nScore += 2;
  }
}

// tallyAdmin%reportPath -> trackClerk%reportPath (trackClerk_class.f90:185-232): bin from the pre-path state, response in the pre-path material
__device__ __noinline__ void scorePath(const TrackArgs& a, const Tables& T, const char* base, const double rPre[3], int matPre, int G, double w, double L, unsigned& nScore) {
  const bool isVoid = (matPre == SB_VOID_MAT);
  const double* x = isVoid ? T.xs : mgRow(a.M, T, matPre, G);
  const bool fissile = isVoid ? false : (T.fissile[matPre - 1] != 0);
  const int nC = a.M.nClerk[a.phase];
  const DClerk* cl = (const DClerk*)(base + a.M.oClerk[a.phase]);
  for (int c = 0; c < nC; ++c) {
    const DClerk& k = cl[c];
    if (k.kind != SB_CLERK_TRACK) continue;
    int bin = clerkBin(k, base, rPre, matPre);
    if (bin == 0) continue;
    int addr = k.addr + k.nResp * (bin - 1) - 1;
    for (int i = 0; i < k.nResp; ++i) {
      double resp = (k.respMT[i] == 0) ? 1.0 : (isVoid ? 0.0 : mgResponse(x, fissile, k.respMT[i]));
      double s = resp * w * L;
      if (s != 0.0) { binAdd(a.bins + addr + i, s); ++nScore; }
    }
  }
}

extern __shared__ __align__(16) char g_trackSmem[];

// SYNC: the warps of a CTA run the event phases in lockstep (CTA barriers between them): the loop body is larger than the
// instruction caches, and in lockstep one instruction fetch serves every warp of the CTA (measured on the CE kernel: 2.3x).
template <int THREADS, int BPS, bool SYNC>
__global__ void __launch_bounds__(THREADS, BPS) k_histories_track(const TrackArgs a) {
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ Model s_M; __shared__ Tables s_T;      // shared, not parameters: the out-of-line geometry functions take them by reference
  const char* base = a.blob;
  if (a.useSmem) { sbh::stageHot(g_trackSmem, a.blob, a.M.blobBytes, &s_bar); base = g_trackSmem; }
  if (threadIdx.x == 0) { s_M = a.M; s_T = bind(a.M, base); }
  __syncthreads();
  const Tables& T = s_T;
  const Model& M = s_M;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned ltMask = (1u << lane) - 1u;
  const double collisionXS = M.collisionXS;

  Coords c; DistCache cache; cache.lvl = 0;
  bool alive = false, exhausted = false;
  const int gLane = blockIdx.x * blockDim.x + threadIdx.x, nLanes = gridDim.x * blockDim.x;
  int nStk = 0;                                                 // fixed source: entries in this history's private buffer
  int hi = -1, G = 1, nSite = 0, hSeg = 0, mode = 0;          // mode: 0 = transport call begins, 1 = delta, 2 = surface
  double w = 0.0, w0 = 0.0, trackXS = 1.0;
  uint64_t rng = 0;
  double sProd = 0.0, sAbs = 0.0, sScat = 0.0, sLeak = 0.0;     // sLeak: leaked weight of the history (with its secondaries in a fixed-source run)
  unsigned nSeg = 0, nColl = 0, nScore = 0;
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
          w = a.in.w[hi]; w0 = w; G = a.in.G[hi];
          rng = sbh::rngSeed(a.seedTab, a.rng0, (unsigned)(a.histOffset + hi + 1));
          if (!placeCoord(M, T, c)) atomicMax(&a.cd->error, SB_ERR_NEST);       // geom%placeCoord (eigenPhysicsPackage_class.f90:224)
          nSite = 0; hSeg = 0; sProd = 0.0; sAbs = 0.0; sScat = 0.0; sLeak = 0.0; mode = 0;
          alive = true;
        }
        need = __ballot_sync(FULL, !alive);
      }
      const bool warpDone = (need == FULL && exhausted);
      if (SYNC) { if (__syncthreads_and(warpDone ? 1 : 0)) break; }
      else if (warpDone) break;
    }

    // ---------------- event: one flight segment -------------------------------------------------------
    bool realColl = false, died = false;
    double leak = 0.0;
    if (alive) {
      if (mode == 0) {                                      // transportOperator%transport begins
        if (a.tracking == SB_TRACK_ST) mode = 2;
        else if (a.tracking == SB_TRACK_DT) mode = 1;
        else {                                              // transportOperatorHT_class.f90:49-81
          double majorant_inv = 1.0 / mgMajorant(M, T, G);
          double sigmaT = (c.mat == SB_VOID_MAT) ? 0.0 : mgRow(M, T, c.mat, G)[XS_TOTAL] + 0.0;
          double ratio = sigmaT * majorant_inv;
          mode = (ratio > (1.0 - a.htCutoff)) ? 1 : 2;
        }
        cache.lvl = 0;
      }
      if (mode == 1) {                                      // deltaTracking, one tentative flight
        trackXS = mgMajorant(M, T, G);
        double majorant_inv = 1.0 / trackXS;
        double distance = negLogHot(rngGet(rng)) * majorant_inv;
        geomTeleportCoords(M, T, c, distance);
        ++nSeg; ++hSeg;
        if (c.mat == SB_OUTSIDE_MAT) { leak = w; sLeak = sLeak + w; died = true; }
        else if (c.mat >= SB_OVERLAP_MAT && c.mat != SB_VOID_MAT) { atomicMax(&a.cd->error, c.mat == SB_UNDEF_MAT ? SB_ERR_UNDEF_MAT : SB_ERR_OVERLAP_MAT); died = true; }
        else {
          bool virt = true;
          if (c.mat != SB_VOID_MAT) {
            double sigmaT = mgRow(M, T, c.mat, G)[XS_TOTAL] + 0.0;
            if (rngGet(rng) < sigmaT * majorant_inv) { realColl = true; virt = false; }
          }
          scoreInColl(a, T, base, c.r[0], c.mat, G, w, trackXS, virt, sProd, sAbs, nScore);
        }
      } else {                                              // surfaceTracking, one segment
        const double tol = 1.0E-12;
        int m = c.mat;
        double sigmaTrack = (m == SB_VOID_MAT) ? collisionXS : fmax(mgRow(M, T, m, G)[XS_TOTAL] + 0.0, collisionXS);
        trackXS = sigmaTrack;
        double dist, invSigmaTrack, sigmaT;
        if (sigmaTrack < tol) { dist = INF; invSigmaTrack = INF; sigmaT = 0.0; }
        else {
          invSigmaTrack = 1.0 / sigmaTrack;
          dist = negLogHot(rngGet(rng)) * invSigmaTrack;
          sigmaT = (m == SB_VOID_MAT) ? 0.0 : mgRow(M, T, m, G)[XS_TOTAL] + 0.0;      // getTotalMatXS of a void region is 0
        }
        int event;
        const double rPre[3] = {c.r[0][0], c.r[0][1], c.r[0][2]};      // p%savePrePath (transportOperatorST_class.f90:107)
        geomMove(M, T, c, dist, event, a.stCache ? &cache : nullptr);
        ++nSeg; ++hSeg;
        if (a.nTrackClerks) scorePath(a, T, base, rPre, m, G, w, dist, nScore);      // tally%reportPath(p, dist), :125
        m = c.mat;
        if (m == SB_OUTSIDE_MAT) { leak = w; sLeak = sLeak + w; died = true; }
        else if (m >= SB_OVERLAP_MAT && m != SB_VOID_MAT) { atomicMax(&a.cd->error, m == SB_UNDEF_MAT ? SB_ERR_UNDEF_MAT : SB_ERR_OVERLAP_MAT); died = true; }
        else if (event == COLL_EV) {
          bool virt = true;
          if (rngGet(rng) < sigmaT * invSigmaTrack) { realColl = true; virt = false; }
          scoreInColl(a, T, base, c.r[0], m, G, w, trackXS, virt, sProd, sAbs, nScore);
        }
      }
    }

    if (SYNC) __syncthreads();
    // ---------------- event: collision, part 1 --------------------------------------------------------
    int MT = 0, nNew = 0;
    const int mat = c.mat;
    if (realColl) {
      const double* x = mgRow(M, T, mat, G);
      (void)rngGet(rng);
      double rr = rngGet(rng);
      {
        int C = 1;
        double xs = x[XS_TOTAL] * rr - 0.0;
        if (xs > 0.0) C += 1;
        xs = xs - x[XS_IESCATTER];
        if (xs > 0.0) C += 1;
        xs = xs - x[XS_CAPTURE];
        if (xs > 0.0) C += 1;
        MT = C;
      }
      ++nColl;
      if (T.fissile[mat - 1] != 0) {
        double rand1 = rngGet(rng);
        nNew = (int)(fabs((w * x[XS_NUFISSION]) / (w0 * x[XS_TOTAL] * a.k_eff)) + rand1);
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
    // ---------------- collision, part 2 ---------------------------------------------------------------
    if (realColl) {
      const double wSite = fsign(w0, w);
      const int nIter = nNew + (MT == 2 ? 1 : 0);
      const int row = (mat - 1) * M.nG + (G - 1);
      for (int i = 0; i < nIter; ++i) {
        const bool isScat = (i == nNew);
        const double* cdf = isScat ? T.P0 + (size_t)row * M.nG : T.chi + (size_t)(mat - 1) * M.nG;
        double mu = 0.0, phi = 0.0, rem;
        if (isScat) rem = rngGet(rng) * T.xs[row * 6 + XS_IESCATTER];
        else { mu = 2.0 * rngGet(rng) - 1.0; phi = TWO_PI * rngGet(rng); rem = rngGet(rng); }
        int Gout = 0;
        for (int g = 1; g <= M.nG; ++g) { rem = rem - cdf[g - 1]; if (rem < 0.0) { Gout = g; break; } }
        if (Gout == 0) { atomicMax(&a.cd->error, SB_ERR_SAMPLING); Gout = G; }
        if (isScat) {
          if (M.isP1) mu = sampleLegendreP1(T.P1[(size_t)row * M.nG + (Gout - 1)], rng);
          else mu = 2.0 * rngGet(rng) - 1.0;
          phi = TWO_PI * rngGet(rng);
        }
        double d[3] = {c.u[0][0], c.u[0][1], c.u[0][2]};
        rotateVector(d, mu, phi);
        if (isScat) {
          double w_mul = T.prod[(size_t)row * M.nG + (Gout - 1)];
          double wPre = w;
          G = Gout;
          w = w * w_mul;
          coordsRotate(T, c, d);
          double sc = fmax(w - wPre, 0.0);
          if (sc > 0.0) sScat += sc;
          mode = 0;                                          // the next flight is a new transport call
        } else if (a.stk.on) {                               // fixed source: a secondary of this history
          if (nStk >= a.stk.cap) atomicMax(&a.cd->error, SB_ERR_BANK_OVERFLOW);
          else { a.stk.push(nStk, gLane, nLanes, c.r[0], d, wSite, 0.0, Gout); ++nStk; }
        } else if (slot >= 0) {
          int s = slot + i;
          a.out.rx[s] = c.r[0][0]; a.out.ry[s] = c.r[0][1]; a.out.rz[s] = c.r[0][2];
          a.out.ux[s] = d[0]; a.out.uy[s] = d[1]; a.out.uz[s] = d[2];
          a.out.w[s] = wSite; a.out.G[s] = Gout; a.out.brood[s] = hi; a.out.seq[s] = nSite + i;
        }
      }
      if (!a.stk.on) nSite += nNew;
      if (MT == 3 || MT == 4) died = true;
      if (MT == 1) mode = 0;
    }
    if (died && nStk > 0) {                                  // bufferLoop: release the last particle detained and carry on (:232-236)
      double Edummy;
      --nStk;
      a.stk.pop(nStk, gLane, nLanes, c.r[0], c.u[0], w, Edummy, G);
      w0 = w;
      if (!placeCoord(M, T, c)) atomicMax(&a.cd->error, SB_ERR_NEST);
      mode = 0; cache.lvl = 0;
      died = false;
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
  }
  if (lane == 0) {
    atomicAdd(&a.cd->nSeg, (unsigned long long)nSeg); atomicAdd(&a.cd->nColl, (unsigned long long)nColl);
    atomicAdd(&a.cd->nScore, (unsigned long long)nScore);
  }
}

}  // namespace sbt
