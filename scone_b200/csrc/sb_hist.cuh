// The history kernel of the MG delta-tracking path and its compact ("hot") device model.
//
//   eigenPhysicsPackage_class.f90:213-252    history loop (one lane = one history at a time)
//   transportOperatorDT_class.f90:47-130     deltaTracking
//   geometryStd_class.f90:119-147,492-514,565-619   placeCoord / teleport / diveToMat
//   collisionProcessor_inter.f90:114-195 + neutronMGstd_class.f90:85-297   collide
//   collisionClerk_class.f90:192-244, keffImplicitClerk_class.f90:180-236  scoring
//
// Layout: every table the loop touches (universe records, geometry graph, pin radii, cross sections,
// scattering matrices, clerk records) is packed by the host into one 16-byte aligned "hot blob" that each
// CTA stages into shared memory with one TMA bulk copy; all loads in the loop are LDS. The generic model
// blob (sb_device.cuh) stays in global memory and is only touched by the out-of-line cold paths
// (rotated universes, general CSG cells, non-box borders, boundary transformations).
//
// Arithmetic is the reference's, operation for operation (no FMA contraction). Two things are done
// differently without changing any result bit:
//   * lattice index floor((r-corner)/pitch) is evaluated as floor((r-corner)*(1/pitch)) and only falls back
//     to the division when the product is within 1e-7 of an integer (the two quotients differ by < 4e-16
//     relative, so the floor can differ only there);
//   * the axial direction of 2-D lattices (pitch = 2*INF, corner = -INF) is skipped: for |z| < 1000 the
//     reference arithmetic gives ijk = 1, r_bar = 0 and offset 0 exactly.
#pragma once
#include "sb_device.cuh"

namespace sbh {
using namespace sbd;

enum { HU_ROOTBOX = 1, HU_PIN = 2, HU_LAT = 3, HU_COLD = 4 };
enum { HF_ROT = 1, HF_GLOBAL = 2, HF_LAT2D = 4, HF_OFFALL = 8, HF_OFFMAP = 16, HF_ORG0 = 32 };

// one universe, 192 bytes.  ROOTBOX: corner = box origin, pitch = halfwidth, abar[0] = surface tolerance
struct __align__(16) HUni {
  int type, flags, n0, n1, n2, outID, aux, pad;
  double org[3];
  double pitch[3], corner[3], abar[3], inv[3], hp[3];
  double pad2[2];
};
static_assert(sizeof(HUni) == 192, "HUni layout");

struct HotLayout {
  int bytes;
  int oUni, oGraph, oAuxD, oAuxI, oXs, oP0, oProd, oP1, oChi, oFissile, oMajT, oMajInv;
  int oClerk[2], nClerk[2];
  int oScoreMask[2];                     // per phase: [nMat*nG + 1] bytes, 1 if any clerk can score a non-zero in (mat, G); last = void
  int nG, nMat, isP1, rootIdx, borderS, borderIsBox;
};

struct Bank {            // particleDungeon as structure of arrays
  double *rx, *ry, *rz, *ux, *uy, *uz, *w;
  double* E;               // continuous-energy runs: particle energy [MeV] (MG runs carry G)
  int *G, *brood, *seq;
};

struct CycleDev {        // small device-resident record of the running cycle
  int nStart, nSites, nextHistory, error;
  int selBin, selRank, nCand, nNew;
  int maxSeg, pad0;                      // longest history of the cycle in flight segments (critical path)
  unsigned long long thrState;
  double thrReal;
  double startWgt, endWgt, impProd, impAbs, scatProd, anaLeak, kAnalog, kImplicit, normFactor;
  unsigned long long nSeg, nColl, nScore;
  // cumulative k of the attachment clerks: [phase] CSUM, CSUM2, batches
  double kCsum[2], kCsum2[2]; int kBatches[2];
  double kCum, kCumStd;
};

struct HistArgs {
  HotLayout L; const char* hot;          // hot blob in global memory (source of the TMA copy / direct use)
  const char* blob;                      // generic model blob: Model header + tables (cold paths)
  const ulonglong2* seedTab;             // [3][1024] affine maps of the LCG for stride*idx*1024^level
  int n; Bank in; Bank out; int cap;
  int* nsites; double *hProd, *hAbs, *hLeak, *hScat;
  double* bins; int phase; int impScores;      // impScores: keffImplicitClerk scores wanted (active phase, or a user clerk)
  uint64_t rng0; int histOffset; double k_eff;
  CycleDev* cd; int refillMin;
};

// ------------------------------------------------------------------------------------------------
// RNG: same stream as sb_rng.h; the int64 -> double conversion is done with two exact magic-number
// subtractions (hi*2^32 + lo rounded once = round-to-nearest of the 63-bit integer, as I2F does)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double rngGet(uint64_t& s) {
  s = (RNG_G * s + 1ULL) & RNG_MASK;
  // the same with every term scaled by 2^-63 (exact): (2^21 + hi*2^-31) - (2^21 + 2^-11) + (2^-11 + lo*2^-63)
  double hi = __hiloint2double(0x41400000, (int)(s >> 32));
  double lo = __hiloint2double(0x3f400000, (int)(s & 0xffffffffu));
  return (hi - 2097152.00048828125) + lo;                                // exact difference; the sum rounds once = RN(s) * 2^-63
}
__device__ __forceinline__ uint64_t rngSeed(const ulonglong2* tab, uint64_t s, unsigned n) {
#pragma unroll
  for (int lvl = 0; lvl < 3; ++lvl) {
    unsigned i = (n >> (10 * lvl)) & 1023u;
    if (i) { ulonglong2 t = __ldg(tab + lvl * 1024 + i); s = (t.x * s + t.y) & RNG_MASK; }
  }
  return s;
}

// ------------------------------------------------------------------------------------------------
// cold paths: generic code over the global model blob, never inlined into the loop
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ const Model& blobModel(const char* blob) { return *(const Model*)blob; }
__device__ __noinline__ int coldFindCell(const char* blob, int ui, double r0, double r1, double r2, double u0, double u1, double u2) {
  const Model& M = blobModel(blob);
  const Tables T = bind(M, blob);
  return uniFindCellCold(T, ui, r0, r1, r2, u0, u1, u2);
}
// universe%enter rotation (universe_inter.f90:400-424)
__device__ __noinline__ void coldRotate(const char* blob, int ui, double* r, double* u) {
  const Model& M = blobModel(blob);
  const double* m = (const double*)(blob + M.oUniDpar) + ui * SB_UNI_NDPAR + 3;
  double a[3] = {r[0], r[1], r[2]}, b[3] = {u[0], u[1], u[2]};
  for (int i = 0; i < 3; ++i) {
    r[i] = m[3 * i] * a[0] + m[3 * i + 1] * a[1] + m[3 * i + 2] * a[2];
    u[i] = m[3 * i] * b[0] + m[3 * i + 1] * b[1] + m[3 * i + 2] * b[2];
  }
}
// geometryStd%teleport, boundary part: transformBC of the border surface
__device__ __noinline__ void coldTransformBC(const char* blob, double* r, double* u) {
  const Model& M = blobModel(blob);
  const int s = M.borderIdx - 1;
  const int* st = (const int*)(blob + M.oSurfType);
  const double* sp = (const double*)(blob + M.oSurfPar);
  int b6[6]; for (int i = 0; i < 6; ++i) b6[i] = M.bc[i];
  surfTransformBC(st[s], sp + s * SB_SURF_NPAR, b6, r, u);
}
__device__ __noinline__ int coldGridSearchUnstruct(const double* bounds, int N, double v) {
  return gridSearch(SB_GRID_UNSTRUCT, 0.0, 0.0, N, bounds, v);
}

// ------------------------------------------------------------------------------------------------
// floor((d)/pitch) with the reciprocal fast path described at the top of the file
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double floorDiv(double d, double pitch, double inv) {
  double t = d * inv;
  double f = floor(t);
  double frac = t - f;
  if (!(fabs(t) < 1.0e6) || frac < 1.0e-7 || frac > 1.0 - 1.0e-7) f = floor(d / pitch);
  return f;
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
extern __shared__ __align__(16) char g_hotSmem[];

__device__ __forceinline__ void stageHot(char* smem, const char* gsrc, int bytes, uint64_t* bar) {
  unsigned barAddr = (unsigned)__cvta_generic_to_shared(bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barAddr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barAddr), "r"(bytes) : "memory");
    int off = 0;
    while (off < bytes) {                  // TMA 1-D bulk copies (UBLKCP), 32 KiB pieces
      int piece = min(bytes - off, 32768);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst + off), "l"(gsrc + off), "r"(piece), "r"(barAddr) : "memory");
      off += piece;
    }
  }
  unsigned done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(barAddr) : "memory");
  }
}

template <bool SMEM, int BPS>
__global__ void __launch_bounds__(256, BPS) k_histories(const HistArgs a) {
  __shared__ __align__(8) uint64_t s_bar;
  const char* hb;
  if (SMEM) { stageHot(g_hotSmem, a.hot, a.L.bytes, &s_bar); hb = g_hotSmem; }
  else hb = a.hot;
  const HUni* const uni = (const HUni*)(hb + a.L.oUni);
  const int2* const graph = (const int2*)(hb + a.L.oGraph);
  const double* const auxD = (const double*)(hb + a.L.oAuxD);
  const int* const auxI = (const int*)(hb + a.L.oAuxI);
  const double* const xsT = (const double*)(hb + a.L.oXs);
  const double* const P0 = (const double*)(hb + a.L.oP0);
  const double* const prodT = (const double*)(hb + a.L.oProd);
  const double* const P1 = (const double*)(hb + a.L.oP1);
  const double* const chiT = (const double*)(hb + a.L.oChi);
  const int* const fissileT = (const int*)(hb + a.L.oFissile);
  const double* const majT = (const double*)(hb + a.L.oMajT);
  const double* const majInvT = (const double*)(hb + a.L.oMajInv);
  const DClerk* const clerks = (const DClerk*)(hb + a.L.oClerk[0]);     // phase offset applied by the host (oClerk[0] = this launch)
  const int nClerk = a.L.nClerk[0];
  const unsigned char* const scoreMask = (const unsigned char*)(hb + a.L.oScoreMask[0]);
  const int nG = a.L.nG;
  const bool active = a.impScores != 0;

  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned ltMask = (1u << lane) - 1u;

  bool alive = false, exhausted = false;
  int hi = -1, G = 1, mat = 0, nSite = 0, hSeg = 0;
  double r0 = 0.0, r1 = 0.0, r2 = 0.0, u0 = 1.0, u1 = 0.0, u2 = 0.0;
  double w = 0.0, w0 = 0.0, flux = 0.0, majInv = 1.0;
  uint64_t rng = 0;
  double sProd = 0.0, sAbs = 0.0, sScat = 0.0, sLeak = 0.0;     // sLeak: leaked weight of the history (with its secondaries in a fixed-source run)
  unsigned nSeg = 0, nColl = 0, nScore = 0;        // per lane, over all its histories

  for (;;) {
    // ---------------- refill dead lanes (warp-level compaction of the bank) ----------------------------
    {
      unsigned need = __ballot_sync(FULL, !alive);
      if (need != 0u && !exhausted) {
        int cnt = __popc(need);
        if (cnt >= a.refillMin || need == FULL) {
          int b = 0;
          if (lane == 0) b = atomicAdd(&a.cd->nextHistory, cnt);
          b = __shfl_sync(FULL, b, 0);
          if (b + cnt >= a.n) exhausted = true;
          int my = b + __popc(need & ltMask);
          if (!alive && my < a.n) {
            hi = my;
            r0 = a.in.rx[hi]; r1 = a.in.ry[hi]; r2 = a.in.rz[hi];
            u0 = a.in.ux[hi]; u1 = a.in.uy[hi]; u2 = a.in.uz[hi];
            w = a.in.w[hi]; w0 = w; G = a.in.G[hi];
            rng = rngSeed(a.seedTab, a.rng0, (unsigned)(a.histOffset + hi + 1));
            // geom%placeCoord of the source site is not needed by delta tracking: the first thing the
            // flight does is teleport + placeCoord (transportOperatorDT_class.f90:57-75)
            majInv = majInvT[G - 1]; flux = w / majT[G - 1];
            nSite = 0; hSeg = 0; sProd = 0.0; sAbs = 0.0; sScat = 0.0; sLeak = 0.0;
            alive = true;
            need &= ~(1u << lane);
          }
          need = __ballot_sync(FULL, !alive);
        }
      }
      if (need == FULL && exhausted) break;
    }

    // ---------------- event: tentative flight = move + cell search + virtual/real decision -------------
    bool realColl = false, died = false;
    double leak = 0.0;
    if (alive) {
      {
        double distance = -sbm::log(rngGet(rng)) * majInv;
        r0 = r0 + distance * u0; r1 = r1 + distance * u1; r2 = r2 + distance * u2;
        ++nSeg; ++hSeg;
      }
      int uid = 0;
#pragma unroll 1
      for (int pass = 0;; ++pass) {
        // ---- placeCoord + diveToMat ----
        double p0 = r0, p1 = r1, p2 = r2, v0 = u0, v1 = u1, v2 = u2;
        int ui = a.L.rootIdx - 1, rootID = 1;
        mat = SB_UNDEF_MAT; uid = -3;
#pragma unroll 1
        for (int lvl = 1; lvl <= MAX_NEST; ++lvl) {
          const HUni& U = uni[ui];
          const int type = U.type, flags = U.flags;
          if (flags & HF_ROT) {
            double tr[3] = {p0, p1, p2}, tu[3] = {v0, v1, v2};
            coldRotate(a.blob, ui, tr, tu);
            p0 = tr[0]; p1 = tr[1]; p2 = tr[2]; v0 = tu[0]; v1 = tu[1]; v2 = tu[2];
          }
          if (!(flags & HF_ORG0)) { p0 = p0 - U.org[0]; p1 = p1 - U.org[1]; p2 = p2 - U.org[2]; }
          int localID;
          double o0 = 0.0, o1 = 0.0, o2 = 0.0;
          if (type == HU_LAT) {                                   // latUniverse_class.f90:270-310
            double f0 = floorDiv(p0 - U.corner[0], U.pitch[0], U.inv[0]) + 1.0;
            double f1 = floorDiv(p1 - U.corner[1], U.pitch[1], U.inv[1]) + 1.0;
            double rb0 = p0 - U.corner[0] - f0 * U.pitch[0] + U.hp[0];
            double rb1 = p1 - U.corner[1] - f1 * U.pitch[1] + U.hp[1];
            if (fabs(rb0) > U.abar[0] && rb0 * v0 > 0.0) f0 += (v0 < 0.0) ? -1.0 : 1.0;
            if (fabs(rb1) > U.abar[1] && rb1 * v1 > 0.0) f1 += (v1 < 0.0) ? -1.0 : 1.0;
            double f2 = 1.0;
            const bool flat = (flags & HF_LAT2D) && fabs(p2) < 1000.0;
            if (!flat) {
              f2 = floor((p2 - U.corner[2]) / U.pitch[2]) + 1.0;
              double rb2 = p2 - U.corner[2] - f2 * U.pitch[2] + U.hp[2];
              if (fabs(rb2) > U.abar[2] && rb2 * v2 > 0.0) f2 += (v2 < 0.0) ? -1.0 : 1.0;
            }
            int i0 = (int)f0, i1 = (int)f1, i2 = (int)f2;
            if (i0 <= 0 || i0 > U.n0 || i1 <= 0 || i1 > U.n1 || i2 <= 0 || i2 > U.n2) localID = U.outID;
            else {
              localID = i0 + U.n0 * (i1 - 1 + U.n1 * (i2 - 1));
              bool doOff = (flags & HF_OFFALL) || ((flags & HF_OFFMAP) && auxI[U.aux + localID - 1] == 1);
              if (doOff) {                                        // cellOffset (latUniverse_class.f90:381-401)
                o0 = (f0 - 0.5) * U.pitch[0] + U.corner[0];
                o1 = (f1 - 0.5) * U.pitch[1] + U.corner[1];
                if (!flat) o2 = (f2 - 0.5) * U.pitch[2] + U.corner[2];
              }
            }
          } else if (type == HU_PIN) {                            // pinUniverse_class.f90:150-172
            double rs = p0 * p0 + p1 * p1;
            double mul = (p0 * v0 + p1 * v1 >= 0.0) ? -1.0 : 1.0;
            const int N = U.n0; const double* r_sq = auxD + U.aux; const double* tol = r_sq + N;
#pragma unroll 1
            for (localID = 1; localID <= N; ++localID) if (rs < r_sq[localID - 1] + mul * tol[localID - 1]) break;
          } else {
            localID = 0;
            if (type == HU_ROOTBOX) {                             // box evaluate + halfspace (box_class.f90:134-146)
              double c = fmax(fmax(fabs(p0 - U.corner[0]) - U.pitch[0], fabs(p1 - U.corner[1]) - U.pitch[1]), fabs(p2 - U.corner[2]) - U.pitch[2]);
              if (fabs(c) >= U.abar[0]) localID = (c > 0.0) ? 2 : 1;
            }
            if (localID == 0) localID = coldFindCell(a.blob, ui, p0, p1, p2, v0, v1, v2);
          }
          int2 f = graph[rootID + localID - 2];
          if (f.x >= 0) { mat = f.x; uid = f.y; break; }
          if (lvl == MAX_NEST) { atomicMax(&a.cd->error, SB_ERR_NEST); break; }
          ui = -f.x - 1; rootID = f.y;
          if (uni[ui].flags & HF_GLOBAL) { p0 = r0; p1 = r1; p2 = r2; }
          else { p0 = p0 - o0; p1 = p1 - o1; p2 = p2 - o2; }
        }
        // ---- geometryStd%teleport: outside -> transformBC, place again (once) ----
        if (mat != SB_OUTSIDE_MAT || pass == 1 || !a.L.borderIsBox) break;
        double tr[3] = {r0, r1, r2}, tu[3] = {u0, u1, u2};
        coldTransformBC(a.blob, tr, tu);
        r0 = tr[0]; r1 = tr[1]; r2 = tr[2]; u0 = tu[0]; u1 = tu[1]; u2 = tu[2];
      }
      (void)uid;
      if (mat == SB_OUTSIDE_MAT) { leak = w; sLeak = sLeak + w; died = true; }                     // LEAK_FATE
      else if (mat >= SB_OVERLAP_MAT && mat != SB_VOID_MAT) {
        atomicMax(&a.cd->error, mat == SB_UNDEF_MAT ? SB_ERR_UNDEF_MAT : SB_ERR_OVERLAP_MAT); died = true;
      } else {
        const bool isVoid = (mat == SB_VOID_MAT);
        const double* x = isVoid ? xsT : xsT + ((mat - 1) * nG + (G - 1)) * 6;
        const bool fissile = isVoid ? false : (fissileT[mat - 1] != 0);
        bool virt = true;
        if (!isVoid) {
          double sigmaT = x[XS_TOTAL] + 0.0;
          if (rngGet(rng) < sigmaT * majInv) { realColl = true; virt = false; }
        }
        // ---- tallyAdmin%reportInColl: collisionClerks, then keffImplicitClerk (active cycles) ----
        // (skipped when no clerk of this phase can score a non-zero value in this material and group)
        const int nC = scoreMask[isVoid ? a.L.nMat * nG : (mat - 1) * nG + (G - 1)] ? nClerk : 0;
#pragma unroll 1
        for (int c = 0; c < nC; ++c) {
          const DClerk& k = clerks[c];
          if (k.kind != SB_CLERK_COLLISION) continue;
          if (!k.handleVirtual && (virt || isVoid)) continue;
          const double f = k.handleVirtual ? flux : w / (x[XS_TOTAL] + 0.0);
          bool any = false;
#pragma unroll 1
          for (int i = 0; i < k.nResp; ++i) {
            double resp = (k.respMT[i] == 0) ? 1.0 : (isVoid ? 0.0 : mgResponse(x, fissile, k.respMT[i]));
            if (resp * f != 0.0) any = true;
          }
          if (!any) continue;
          int bin = 1;                                            // multiMap (multiMap_class.f90:153-173)
#pragma unroll 1
          for (int i = 0; i < k.nMaps && bin > 0; ++i) {
            int b = 0;
            if (k.mapType[i] == SB_MAP_SPACE) {
              const int ax = k.mapAxis[i];
              const double v = (ax == 0) ? r0 : ((ax == 1) ? r1 : r2);
              if (k.mapGrid[i] == SB_GRID_LIN) {                  // grid_class.f90:154-176
                double fl = floorDiv(v - k.mapFirst[i], k.mapStep[i], k.mapInv[i]);
                b = (int)fl + 1;
                if (b < 1 || b >= k.mapN[i] + 1) b = 0;
              } else if (k.mapGrid[i] == SB_GRID_UNSTRUCT) b = coldGridSearchUnstruct((const double*)(hb + k.mapOff[i]), k.mapN[i], v);
            } else if (k.mapType[i] == SB_MAP_MATERIAL) {
              const int* mb = (const int*)(hb + k.mapOff[i]);     // mapGrid holds the table length (n_mat)
              b = (mat >= 1 && mat <= k.mapGrid[i]) ? mb[mat - 1] : k.mapDef[i];
            }                                                     // energyMap: MG particles are not scored
            bin = (b == 0) ? 0 : bin + (b - 1) * k.mapMul[i];
          }
          if (bin == 0) continue;
          const int addr = k.addr + k.nResp * (bin - 1) - 1;      // 0-based slot of response 1
#pragma unroll 1
          for (int i = 0; i < k.nResp; ++i) {
            double resp = (k.respMT[i] == 0) ? 1.0 : (isVoid ? 0.0 : mgResponse(x, fissile, k.respMT[i]));
            double s = resp * f;
            if (s != 0.0) { binAdd(a.bins + addr + i, s); ++nScore; }
          }
        }
        if (active && !isVoid) {
          double nuf = fissile ? x[XS_NUFISSION] : 0.0, fis = fissile ? x[XS_FISSION] : 0.0;
          sProd += nuf * flux;
          sAbs += (x[XS_CAPTURE] + fis) * flux;
          nScore += 2;
        }
      }
    }

    // ---------------- event: collision, part 1 (channel + number of fission sites) --------------------
    int MT = 0, nNew = 0;
    if (realColl) {
      const double* x = xsT + ((mat - 1) * nG + (G - 1)) * 6;
      (void)rngGet(rng);                                  // alpha-absorption test always draws (probAlpha = 0)
      double rr = rngGet(rng);
      {                                                   // neutronMacroXSs%invert (neutronXsPackages_class.f90:211-250)
        int C = 1;
        double xs = x[XS_TOTAL] * rr - 0.0;
        if (xs > 0.0) C += 1;
        xs = xs - x[XS_IESCATTER];
        if (xs > 0.0) C += 1;
        xs = xs - x[XS_CAPTURE];
        if (xs > 0.0) C += 1;
        MT = C;                                           // 1 elastic, 2 inelastic, 3 capture, 4 fission
      }
      ++nColl;
      if (fissileT[mat - 1] != 0) {                       // neutronMGstd implicit (:131-199)
        double rand1 = rngGet(rng);
        nNew = (int)(fabs((w * x[XS_NUFISSION]) / (w0 * x[XS_TOTAL] * a.k_eff)) + rand1);
        if (nNew < 0) nNew = 0;
      }
    }

    // ---------------- warp-aggregated allocation of fission-bank slots --------------------------------
    int slot = -1;
    {
      unsigned spawn = __ballot_sync(FULL, nNew > 0);
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

    // ---------------- collision, part 2: fission sites, then the scattered neutron --------------------
    // one loop, one rotateVector: iterations 0..nNew-1 emit sites (fissionMG%sampleOut: mu, phi, then chi),
    // the last iteration is the scattering itself (multiScatterMG%sampleOut: G_out, then mu, phi)
    if (realColl) {
      const double wSite = fsign(w0, w);
      const int nIter = nNew + (MT == 2 ? 1 : 0);
      const int row = (mat - 1) * nG + (G - 1);
#pragma unroll 1
      for (int i = 0; i < nIter; ++i) {
        const bool isScat = (i == nNew);
        const double* cdf = isScat ? P0 + row * nG : chiT + (mat - 1) * nG;
        double mu = 0.0, phi = 0.0, rem;
        if (isScat) rem = rngGet(rng) * xsT[row * 6 + XS_IESCATTER];
        else { mu = 2.0 * rngGet(rng) - 1.0; phi = TWO_PI * rngGet(rng); rem = rngGet(rng); }
        int Gout = 0;
#pragma unroll 1
        for (int g = 1; g <= nG; ++g) { rem = rem - cdf[g - 1]; if (rem < 0.0) { Gout = g; break; } }
        if (Gout == 0) { atomicMax(&a.cd->error, SB_ERR_SAMPLING); Gout = G; }
        if (isScat) {
          if (a.L.isP1) mu = sampleLegendreP1(P1[row * nG + (Gout - 1)], rng);
          else mu = 2.0 * rngGet(rng) - 1.0;
          phi = TWO_PI * rngGet(rng);
        }
        double d[3] = {u0, u1, u2};
        rotateVector(d, mu, phi);
        if (isScat) {                                       // neutronMGstd inelastic (:221-252)
          double w_mul = prodT[row * nG + (Gout - 1)];
          double wPre = w;
          G = Gout;
          majInv = majInvT[G - 1];
          w = w * w_mul;
          flux = w / majT[G - 1];
          u0 = d[0]; u1 = d[1]; u2 = d[2];
          double sc = fmax(w - wPre, 0.0);                   // keffImplicitClerk%reportOutColl
          if (sc > 0.0) sScat += sc;
        } else if (slot >= 0) {
          int s = slot + i;
          a.out.rx[s] = r0; a.out.ry[s] = r1; a.out.rz[s] = r2;
          a.out.ux[s] = d[0]; a.out.uy[s] = d[1]; a.out.uz[s] = d[2];
          a.out.w[s] = wSite; a.out.G[s] = Gout; a.out.brood[s] = hi; a.out.seq[s] = nSite + i;
        }
      }
      nSite += nNew;
      if (MT == 3 || MT == 4) died = true;                   // capture / fission: history ends (ABS_FATE)
      // MT == 1 (elastic) cannot be selected for MG data (elasticScatter = 0): "Do nothing"
    }

    if (died) {
      a.nsites[hi] = nSite;
      a.hProd[hi] = sProd; a.hAbs[hi] = sAbs; a.hLeak[hi] = sLeak; a.hScat[hi] = sScat;
      if (hSeg > 256) atomicMax(&a.cd->maxSeg, hSeg);
      alive = false;
    }
  }

  // ---------------- per-warp event counters (integers: order-independent) -----------------------------
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    nSeg += __shfl_down_sync(FULL, nSeg, d); nColl += __shfl_down_sync(FULL, nColl, d); nScore += __shfl_down_sync(FULL, nScore, d);
  }
  if (lane == 0) {
    atomicAdd(&a.cd->nSeg, (unsigned long long)nSeg); atomicAdd(&a.cd->nColl, (unsigned long long)nColl);
    atomicAdd(&a.cd->nScore, (unsigned long long)nScore);
  }
}

}  // namespace sbh
