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
#include <type_traits>

namespace sbh {
using namespace sbd;

enum { HU_ROOTBOX = 1, HU_PIN = 2, HU_LAT = 3, HU_COLD = 4 };
enum { HF_ROT = 1, HF_GLOBAL = 2, HF_LAT2D = 4, HF_OFFALL = 8, HF_OFFMAP = 16, HF_ORG0 = 32 };

// one universe, 208 bytes, laid out for 16-byte shared-memory loads.  Per axis a: ci[a] = {corner, 1/pitch},
// ph[a] = {pitch, pitch/2}, ab[a] = a_bar.  ROOTBOX: ci[a].x = box origin, ph[a].x = halfwidth, ab[0] = surface tolerance
struct __align__(16) HUni {
  int type, flags, n0, n1, n2, outID, aux, pad;
  double org[3], pad0;
  double2 ci[3];
  double2 ph[3];
  double ab[3], pad1;
  double pad2[2];                      // 208 bytes = 52 words: consecutive universes start 20 banks apart (192 bytes would put every second one on the same banks)
};
static_assert(sizeof(HUni) == 208, "HUni layout");

struct HotLayout {
  int bytes;
  int oUni, oGraph, oAuxD, oAuxI, oXs, oP0, oProd, oP1, oP0First, oFissile, oMajT, oMajInv;
  int oClerk[2], nClerk[2];
  int oScoreMask[2];                     // per phase: [nMat*nG + 1] bytes, 1 if any clerk can score a non-zero in (mat, G); last = void
  int nG, nMat, isP1, rootIdx, borderS, borderIsBox;      // borderIsBox: 1 box-like border (transformBC applies), 2 it is the root universe's box (hot record)
  int bc[6]; double borderTol;           // boundary conditions and SURF_TOL of that box
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
  unsigned long long nXsTerms;           // continuous energy: nuclide terms of the flights' total-cross-section lookups
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
  unsigned laneMask;                     // lanes of a warp that take histories from the bank (all: 0xffffffff)
  int maxSegMin;                         // histories longer than this report their length (cd->maxSeg)
  long long* prof;                       // SB_PROFILE_ROUNDS builds: per-warp round timings
  int cellCache;                         // placement resumes below the lattice cell of the previous site when the new one is safely inside it,
                                         // for histories older than this many flights (a huge value switches the cache off)
  int loneMode;                          // 1: a history left alone in its warp gets its random numbers from the warp's draw window
  int assist;                            // > 0: once the bank is exhausted, a warp with no more than this many histories left hands them to k_lone
  struct LoneRec* loneQ; int* loneCount; int* loneNext; int loneCap;     // the queue between the two kernels
  int* loneReady; int* loneDone; int loneTag, loneWarps;                 // per slot: tag of the launch that filled it; warps of k_histories that have ended, of loneWarps
};

// ------------------------------------------------------------------------------------------------
// RNG: same stream as sb_rng.h; the int64 -> double conversion is done with two exact magic-number
// subtractions (hi*2^32 + lo rounded once = round-to-nearest of the 63-bit integer, as I2F does)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double rngReal(uint64_t s) {
  // the same with every term scaled by 2^-63 (exact): (2^21 + hi*2^-31) - (2^21 + 2^-11) + (2^-11 + lo*2^-63)
  double hi = __hiloint2double(0x41400000, (int)(s >> 32));
  double lo = __hiloint2double(0x3f400000, (int)(s & 0xffffffffu));
  return (hi - 2097152.00048828125) + lo;                                // exact difference; the sum rounds once = RN(s) * 2^-63
}
__device__ __forceinline__ double rngGet(uint64_t& s) {
  s = (RNG_G * s + 1ULL) & RNG_MASK;
  return rngReal(s);
}
// state after K draws as ONE affine map of the current state (K compile-time): the draws of an event whose order is
// fixed are then independent of one another instead of a chain of multiplications
__host__ __device__ constexpr uint64_t rngJumpA(int k) { uint64_t a = 1; for (int i = 0; i < k; ++i) a = (a * RNG_G) & RNG_MASK; return a; }
__host__ __device__ constexpr uint64_t rngJumpC(int k) { uint64_t c = 0; for (int i = 0; i < k; ++i) c = (c * RNG_G + 1ULL) & RNG_MASK; return c; }
template <int K> __device__ __forceinline__ uint64_t rngJump(uint64_t s) {
  constexpr uint64_t A = rngJumpA(K), C = rngJumpC(K);
  return (A * s + C) & RNG_MASK;
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
// a rotated universe on the way down (universe_inter.f90:400-424): the whole placement is redone by the generic search
// over the model blob, which carries the rotated direction from level to level; returns the material (nesting overflow: -1)
__device__ __noinline__ int coldPlace(const char* blob, double r0, double r1, double r2, double u0, double u1, double u2) {
  const Model& M = blobModel(blob);
  const Tables T = bind(M, blob);
  const double r[3] = {r0, r1, r2}, u[3] = {u0, u1, u2};
  int mat, uid;
  return geomPlace(M, T, r, u, mat, uid) ? mat : -1;
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
// true if floor(d * inv) may differ from floor(d / pitch): the product is within 1e-7 of an integer, huge or NaN
__device__ __forceinline__ bool floorUnsafe(double t, double fl) {
  return !(fabs((t - fl) - 0.5) < 0.5 - 1.0e-7) || !(fabs(t) < 1.0e6);
}

// ------------------------------------------------------------------------------------------------
// draw window of a history that is alone in its warp (the tail of every cycle: a few long histories, one per warp).
// The 31 idle lanes compute the NEXT 32 numbers of the history's stream at once - state, uniform, -log, sin / cos of
// the azimuth 2 pi xi, sin of the polar angle of mu = 2 xi - 1 - and the history picks what the reference's draw order
// asks for: the random-number arithmetic, the logarithm and the trigonometry leave the history's dependent chain.
// Every value is what the inline code computes from the same state, so histories do not change.
// ------------------------------------------------------------------------------------------------
constexpr int WIN_SMALL = 64, WIN_BIG = 128; // numbers in a draw window: two per lane; four in the kernels that run speculative batches
constexpr int WIN_ROUND = 12;              // draws a round may take from the window before the generic fallback: flight 2 + channel 3 + one site 3 + scattering 3 (+1)
template <int N> struct DrawWinT {
  unsigned long long st[N];
  double xi[N], nlog[N], sn[N], cs[N], A[N];
};

// ------------------------------------------------------------------------------------------------
// The end of a cycle: two kernels. At the populations of an eigenvalue cycle the time of the history kernel is the dependent
// chain of its longest histories - thermal neutrons that scatter hundreds of times in the moderator and the reflector. Once the
// bank is exhausted, a warp of k_histories that is down to a few histories (HistArgs::assist) writes them to a queue in global
// memory and ends; k_lone, launched behind k_histories with programmatic dependent launch and NOT waiting for it, takes the SMs
// the CTAs of k_histories leave and runs the queued histories one per warp, while the other warps of k_histories are still
// working. k_lone is a kernel of its own so that its loop gets its own register allocation (198 registers, 8 warps per SM)
// instead of spilling in the loop every history runs.
//
// speculative batches of a history that has a warp to itself (k_lone). Most collisions of these histories happen in the
// material of the collision before. The lane
// therefore runs up to SPEC_MAX rounds AHEAD under the assumption that the material does not change: flight, acceptance test,
// channel, outgoing group, rotated direction, weight, implicit k-eff scores - everything but the geometry search and the
// clerks, which do not feed back into the history when the material is known. It leaves one record per round in shared
// memory. Then the 32 lanes of the warp search the geometry for the recorded collision points at the same time (one point per
// lane, one instruction stream) and the history keeps the rounds in front of the first point that lies in another material
// (or beyond the boundary): their clerk scores are made by the lanes that checked them, the history's own variables are
// taken from the record of the first round that is not kept, and the next batch assumes the material found there. A history
// that changes material often (fuel lattice), meets the boundary or void, or banks fission sites runs its rounds one after the
// other with the search in its place. Every number is drawn from the same stream position and every operation is the one the
// loop of k_histories makes, so histories do not change (tests/test_gpu_eigen.py: banks bit-identical to the oracle and to
// k_histories alone, SB_ASSIST=0).
// ------------------------------------------------------------------------------------------------
constexpr int SPEC_MAX = 16;
enum { SR_REAL = 1, SR_DIED = 2, SR_SAMPLING = 4 };
struct __align__(16) SpecRec {              // the history in front of round k, and the collision point of round k
  double b0, b1, b2, u0, u1, u2, w;         // position and direction before the flight, weight
  double sProd, sAbs, sScat;                // the history's scores so far
  double r0, r1, r2;                        // the tentative collision point of the round
  int G, p, nColl, flags;                   // group, position in the draw window, real collisions so far; SR_* of the round
  double pad;
};
static_assert(sizeof(SpecRec) == 128, "SpecRec layout");
struct JumpTab { ulonglong2 j[32]; };        // the affine map of lane + 1 draws, per lane (draw windows)

// ------------------------------------------------------------------------------------------------
// cell cache of the placement. A history that scatters in a moderator collides again a fraction of a pitch away: most
// tentative collision sites lie in the lattice cell of the previous one. Under a root box, up to two nested 2-D lattices
// (no rotation, no origin shift) the search is then known in advance: the lane keeps, from its last full search, the
// offsets (cellOffset of each lattice level), the universe filling the innermost lattice cell and a SAFE BOX in root
// coordinates - the intersection of the root box and of the lattice cells found, shrunk by GC_MARGIN, far more than the
// surface tolerances and the rounding of the index arithmetic. A point strictly inside the safe box gets the same lattice
// indices from latUniverse%findCell at every level (no face adjustment can trigger), so its local coordinates are the
// same two subtractions r - offset_A - offset_B the full search makes, and the search resumes at the cached universe.
// Anything else (near a face, another cell, 3-D lattices, rotated / shifted universes) takes the full search.
// ------------------------------------------------------------------------------------------------
constexpr double GC_MARGIN = 1.0e-8;

// ------------------------------------------------------------------------------------------------
// what the event code of a history reads: the tables of the hot blob (shared memory when the kernel stages it), the per-thread
// cell caches and scatter scores. Built once per thread by the kernel; the lone-history loop, which is a function of its own
// (own register allocation), builds the same from its arguments.
// ------------------------------------------------------------------------------------------------
struct GcRows { double2* base; int stride; __device__ __forceinline__ double2* operator[](int k) const { return base + k * stride; } };
struct HotCtx {
  const char* hb;
  const HUni* uni; const int2* graph; const double* auxD; const int* auxI;
  const double *xsT, *P0, *prodT, *P1, *majT, *majInvT;
  const int *p0First, *fissileT;
  const DClerk* clerks; const unsigned char* scoreMask;
  int nClerk, nG; bool active;
  GcRows gcB, gcO; int4 *gcI, *gcP; double* scat;
};
__device__ __forceinline__ HotCtx makeHotCtx(const HistArgs& a, const char* hb, char* cacheBase, int threads) {
  HotCtx H;
  H.hb = hb;
  H.uni = (const HUni*)(hb + a.L.oUni); H.graph = (const int2*)(hb + a.L.oGraph);
  H.auxD = (const double*)(hb + a.L.oAuxD); H.auxI = (const int*)(hb + a.L.oAuxI);
  H.xsT = (const double*)(hb + a.L.oXs); H.P0 = (const double*)(hb + a.L.oP0); H.prodT = (const double*)(hb + a.L.oProd);
  H.P1 = (const double*)(hb + a.L.oP1);
  H.p0First = (const int*)(hb + a.L.oP0First);             // per (material, group in): first non-zero term of the P0 row
  H.fissileT = (const int*)(hb + a.L.oFissile);
  H.majT = (const double*)(hb + a.L.oMajT); H.majInvT = (const double*)(hb + a.L.oMajInv);
  H.clerks = (const DClerk*)(hb + a.L.oClerk[0]);          // phase offset applied by the host (oClerk[0] = this launch)
  H.nClerk = a.L.nClerk[0];
  H.scoreMask = (const unsigned char*)(hb + a.L.oScoreMask[0]);
  H.nG = a.L.nG; H.active = a.impScores != 0;
  // [cell caches: safe box {lo, hi} per axis | offsets of the outer (A) and inner (B) lattice level {x, y}, origin of the cached universe |
  //  universe to resume at, its rootID, its level | if that universe is a plain pin: position of its radii, their number, has-origin flag (else -1)]
  H.gcB.base = (double2*)cacheBase; H.gcB.stride = threads;
  H.gcO.base = H.gcB.base + 3 * threads; H.gcO.stride = threads;
  H.gcI = (int4*)(H.gcO.base + 3 * threads);
  H.gcP = H.gcI + threads;
  H.scat = (double*)(H.gcP + threads);                     // keffImplicitClerk%reportOutColl score of the history (non-zero only with multiplicities)
  return H;
}

// geometryStd%placeCoord + diveToMat of the point (r0, r1, r2) flying along (u0, u1, u2), and the one teleport of delta tracking when it is
// outside: sets mat; true if the boundary transformed the point (and perhaps the direction). direct: the level loop starts at the level the
// cell cache resumes at (a history alone in its warp); otherwise lanes that resume below join the others at their level
__device__ __forceinline__ bool placePoint(const HistArgs& a, const HotCtx& H, const bool direct, const int hSeg,
                                           double& r0, double& r1, double& r2, double& u0, double& u1, double& u2, int& mat) {
  const HUni* const uni = H.uni; const int2* const graph = H.graph; const double* const auxD = H.auxD; const int* const auxI = H.auxI;
  const GcRows s_gcB = H.gcB, s_gcO = H.gcO; int4* const s_gcI = H.gcI; int4* const s_gcP = H.gcP;

  bool tele = false;
#pragma unroll 1
  for (int pass = 0;; ++pass) {
    // ---- placeCoord + diveToMat ----
    double p0 = r0, p1 = r1, p2 = r2;
    double o0 = 0.0, o1 = 0.0, o2 = 0.0;
    int ui = a.L.rootIdx - 1, rootID = 1, lvl0 = 1, pinMat = -1;
    // cacheable prefix of this search: 0 nothing yet, 1 root box passed, 2 / 3 one / two lattice levels passed, -1 closed
    // (young histories are fast neutrons with long flights: the cache only starts after a.cellCache flights)
    int gcN = (hSeg > a.cellCache) ? 0 : -1, gcA = 0, gcB = 0; double ga0 = 0.0, ga1 = 0.0;   // gcA / gcB: the lattices passed; (ga0, ga1): cellOffset of the first when a second follows
    if (gcN == 0) {
      const double2 b0 = s_gcB[0][threadIdx.x], b1 = s_gcB[1][threadIdx.x], b2 = s_gcB[2][threadIdx.x];
      if (r0 > b0.x && r0 < b0.y && r1 > b1.x && r1 < b1.y && r2 > b2.x && r2 < b2.y) {      // inside the safe box of the cached cell
        const double2 oa = s_gcO[0][threadIdx.x], ob = s_gcO[1][threadIdx.x], og = s_gcO[2][threadIdx.x];
        const int4 gi = s_gcI[threadIdx.x], gp = s_gcP[threadIdx.x];
        p0 = r0 - oa.x; p1 = r1 - oa.y; o0 = ob.x; o1 = ob.y;
        ui = gi.x; rootID = gi.y; lvl0 = gi.z; gcN = -1;
        if (gp.x >= 0) {                                      // the cached universe is a plain pin: pinUniverse%findCell right here
          double q0 = p0 - o0, q1 = p1 - o1;
          if (gp.z) { q0 = q0 - og.x; q1 = q1 - og.y; }       // universe%enter: its origin
          const double rs = q0 * q0 + q1 * q1;
          const double mul = (q0 * u0 + q1 * u1 >= 0.0) ? -1.0 : 1.0;
          const int N = gp.y; const double* r_sq = auxD + gp.x; const double* tol = r_sq + N;
          int localID;
#pragma unroll 1
          for (localID = 1; localID <= N; ++localID) if (rs < r_sq[localID - 1] + mul * tol[localID - 1]) break;
          const int2 f = graph[rootID + localID - 2];
          if (f.x >= 0) pinMat = f.x;                          // (a universe in a pin ring: the level loop below goes on from the cached universe)
        }
      }
    }
    mat = SB_UNDEF_MAT;
    if (pinMat >= 0) { mat = pinMat; break; }
#pragma unroll 1
    for (int lvl = direct ? lvl0 : 1; lvl <= MAX_NEST; ++lvl) {      // (a history alone in its warp starts where it resumes)
      if (lvl < lvl0) continue;                               // a lane that resumes below joins the others at its level
      const HUni& U = uni[ui];
      const int4 h0 = *(const int4*)&U.type;                  // type, flags, n0, n1
      const int type = h0.x, flags = h0.y;
      if (gcN > 0 && !(type == HU_LAT && gcN < 3 && (flags & (HF_LAT2D | HF_ORG0 | HF_ROT | HF_GLOBAL)) == (HF_LAT2D | HF_ORG0))) {
        // the cacheable prefix ends above this universe: remember where to resume and the safe box
        // (o0, o1) is still the cellOffset of the last lattice passed
        if (gcN > 1) {
          const HUni& R = uni[a.L.rootIdx - 1]; const HUni& LA = uni[gcA];
          const double ax = (gcN == 3) ? ga0 : o0, ay = (gcN == 3) ? ga1 : o1;              // centre of the cell of lattice A, root coordinates
          double l0 = fmax(R.ci[0].x - R.ph[0].x, ax - LA.ph[0].y), h0x = fmin(R.ci[0].x + R.ph[0].x, ax + LA.ph[0].y);
          double l1 = fmax(R.ci[1].x - R.ph[1].x, ay - LA.ph[1].y), h1x = fmin(R.ci[1].x + R.ph[1].x, ay + LA.ph[1].y);
          if (gcN == 3) {
            const HUni& LB = uni[gcB]; const double bx = ga0 + o0, by = ga1 + o1;           // centre of the cell of lattice B
            l0 = fmax(l0, bx - LB.ph[0].y); h0x = fmin(h0x, bx + LB.ph[0].y); l1 = fmax(l1, by - LB.ph[1].y); h1x = fmin(h1x, by + LB.ph[1].y);
          }
          s_gcB[0][threadIdx.x] = make_double2(l0 + GC_MARGIN, h0x - GC_MARGIN); s_gcB[1][threadIdx.x] = make_double2(l1 + GC_MARGIN, h1x - GC_MARGIN);
          s_gcB[2][threadIdx.x] = make_double2(fmax(R.ci[2].x - R.ph[2].x, -999.0) + GC_MARGIN, fmin(R.ci[2].x + R.ph[2].x, 999.0) - GC_MARGIN);   // |z| < 1000: the 2-D lattices
          s_gcO[0][threadIdx.x] = (gcN == 3) ? make_double2(ga0, ga1) : make_double2(0.0, 0.0);
          s_gcO[1][threadIdx.x] = make_double2(o0, o1);
          s_gcI[threadIdx.x] = make_int4(ui, rootID, lvl, 0);
          const bool plainPin = type == HU_PIN && !(flags & (HF_ROT | HF_GLOBAL));
          s_gcP[threadIdx.x] = make_int4(plainPin ? U.aux : -1, h0.z, (flags & HF_ORG0) ? 0 : 1, 0);
          s_gcO[2][threadIdx.x] = make_double2(U.org[0], U.org[1]);
        }
        gcN = -1;
      }
      if (gcN == 2) { ga0 = o0; ga1 = o1; }                   // a second lattice follows the first: keep the first's cellOffset
      if (lvl > 1) {                                          // local coordinates of the universe below the cell just found
        if (flags & HF_GLOBAL) { p0 = r0; p1 = r1; p2 = r2; }
        else { p0 = p0 - o0; p1 = p1 - o1; p2 = p2 - o2; }
      }
      if (flags & HF_ROT) {                                   // rare: hand the whole placement to the generic search
        mat = coldPlace(a.blob, r0, r1, r2, u0, u1, u2);
        if (mat == -1) { atomicMax(&a.cd->error, SB_ERR_NEST); mat = SB_UNDEF_MAT; }
        break;
      }
      if (!(flags & HF_ORG0)) { p0 = p0 - U.org[0]; p1 = p1 - U.org[1]; p2 = p2 - U.org[2]; }
      int localID;
      o0 = 0.0; o1 = 0.0; o2 = 0.0;
      if (type == HU_LAT) {                                   // latUniverse_class.f90:270-310
        const double2 c0 = U.ci[0], c1 = U.ci[1], q0 = U.ph[0], q1 = U.ph[1];
        const double d0 = p0 - c0.x, d1 = p1 - c1.x;
        const double t0 = d0 * c0.y, t1 = d1 * c1.y;
        double f0 = floor(t0), f1 = floor(t1);
        if (floorUnsafe(t0, f0) || floorUnsafe(t1, f1)) { f0 = floor(d0 / q0.x); f1 = floor(d1 / q1.x); }
        f0 = f0 + 1.0; f1 = f1 + 1.0;
        const double rb0 = d0 - f0 * q0.x + q0.y;
        const double rb1 = d1 - f1 * q1.x + q1.y;
        if (fabs(rb0) > U.ab[0] || fabs(rb1) > U.ab[1]) {     // within the surface tolerance of a cell face: the direction decides
          if (fabs(rb0) > U.ab[0] && rb0 * u0 > 0.0) f0 += (u0 < 0.0) ? -1.0 : 1.0;
          if (fabs(rb1) > U.ab[1] && rb1 * u1 > 0.0) f1 += (u1 < 0.0) ? -1.0 : 1.0;
        }
        const int4 h1 = *(const int4*)&U.n2;                  // n2, outID, aux, pad
        double f2 = 1.0;
        const bool flat = (flags & HF_LAT2D) && fabs(p2) < 1000.0;
        if (!flat) {
          const double2 c2 = U.ci[2], q2 = U.ph[2];
          f2 = floor((p2 - c2.x) / q2.x) + 1.0;
          double rb2 = p2 - c2.x - f2 * q2.x + q2.y;
          if (fabs(rb2) > U.ab[2] && rb2 * u2 > 0.0) f2 += (u2 < 0.0) ? -1.0 : 1.0;
        }
        const int i0 = (int)f0, i1 = (int)f1, i2 = (int)f2;
        bool doOff = false;
        if ((unsigned)(i0 - 1) >= (unsigned)h0.z || (unsigned)(i1 - 1) >= (unsigned)h0.w || (unsigned)(i2 - 1) >= (unsigned)h1.x) localID = h1.y;
        else {
          localID = i0 + h0.z * (i1 - 1 + h0.w * (i2 - 1));
          doOff = (flags & HF_OFFALL) || ((flags & HF_OFFMAP) && auxI[h1.z + localID - 1] == 1);
          if (doOff) {                                        // cellOffset (latUniverse_class.f90:381-401)
            o0 = (f0 - 0.5) * q0.x + c0.x;
            o1 = (f1 - 0.5) * q1.x + c1.x;
            if (!flat) o2 = (f2 - 0.5) * U.ph[2].x + U.ci[2].x;
          }
        }
        if (gcN > 0) {                                        // cell cache: this lattice cell in root coordinates
          if (flat && doOff) {                                // (the cellOffset is the centre of the cell)
            if (gcN == 1) gcA = ui; else { gcB = ui; }
            ++gcN;
          } else gcN = -1;
        }
      } else if (type == HU_PIN) {                            // pinUniverse_class.f90:150-172
        double rs = p0 * p0 + p1 * p1;
        double mul = (p0 * u0 + p1 * u1 >= 0.0) ? -1.0 : 1.0;
        const int N = h0.z; const double* r_sq = auxD + U.aux; const double* tol = r_sq + N;
#pragma unroll 1
        for (localID = 1; localID <= N; ++localID) if (rs < r_sq[localID - 1] + mul * tol[localID - 1]) break;
      } else {
        localID = 0;
        if (type == HU_ROOTBOX) {                             // box evaluate + halfspace (box_class.f90:134-146)
          double c = fmax(fmax(fabs(p0 - U.ci[0].x) - U.ph[0].x, fabs(p1 - U.ci[1].x) - U.ph[1].x), fabs(p2 - U.ci[2].x) - U.ph[2].x);
          if (fabs(c) >= U.ab[0]) localID = (c > 0.0) ? 2 : 1;
          if (gcN == 0 && lvl == 1 && localID == 1 && (flags & (HF_ORG0 | HF_ROT)) == HF_ORG0) {
            gcN = 1;                                          // cell cache: the search passed a plain root box
          }
        }
        if (gcN == 0) gcN = -1;
        if (localID == 0) localID = coldFindCell(a.blob, ui, p0, p1, p2, u0, u1, u2);
      }
      const int2 f = graph[rootID + localID - 2];
      if (f.x >= 0) { mat = f.x; break; }
      if (lvl == MAX_NEST) { atomicMax(&a.cd->error, SB_ERR_NEST); break; }
      ui = -f.x - 1; rootID = f.y;
    }
    // ---- geometryStd%teleport: outside -> transformBC, place again (once) ----
    if (mat != SB_OUTSIDE_MAT || pass == 1 || !a.L.borderIsBox) break;
    tele = true;
    if (a.L.borderIsBox == 2) {
      // the border is the root universe's box (box_class.f90:432-487 transformBC): an axis on which the point is within
      // a_bar = halfwidth (1 - SURF_TOL) of the origin makes no reflection - |d| / a_bar <= 1 exactly when |d| <= a_bar - so
      // the quotient is only formed for the axes that are outside
      const HUni& R = uni[a.L.rootIdx - 1];
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) {
        double& rc = (ax == 0) ? r0 : ((ax == 1) ? r1 : r2); double& uc = (ax == 0) ? u0 : ((ax == 1) ? u1 : u2);
        const double org = R.ci[ax].x, hw = R.ph[ax].x;
        const double a_bar = hw * (1.0 - a.L.borderTol);
        if (fabs(rc - org) <= a_bar) continue;
        const int Ri = (int)ceil(fabs(rc - org) / a_bar) / 2;
#pragma unroll 1
        for (int t = 1; t <= Ri; ++t) {
          const double d0 = rc - org;
          const int b = (d0 < 0.0) ? a.L.bc[2 * ax] : a.L.bc[2 * ax + 1];
          if (b == 1) { const double a0 = fsign(hw, d0) + org; const double d = rc - a0; rc = rc - 2.0 * d; uc = -uc; }
          else if (b == 2) { const double d = fsign(hw, d0); rc = rc - 2.0 * d; }
        }
      }
    } else {
      double tr[3] = {r0, r1, r2}, tu[3] = {u0, u1, u2};
      coldTransformBC(a.blob, tr, tu);
      r0 = tr[0]; r1 = tr[1]; r2 = tr[2]; u0 = tu[0]; u1 = tu[1]; u2 = tu[2];
    }
  }
  return tele;
}

// scores of a (tentative) collision in material mat: collision clerks (tallyAdmin%reportInColl) and, with keff, the implicit k-eff estimators
__device__ __forceinline__ void scoreColl(const HistArgs& a, const HotCtx& H, const bool isVoid, const bool virt, const bool keff,
                                          const double r0, const double r1, const double r2, const int mat, const int G, const double w, const double majInv,
                                          double& sProd, double& sAbs, unsigned& nScore) {
  const double* const xsT = H.xsT; const int* const fissileT = H.fissileT; const double* const majT = H.majT; const DClerk* const clerks = H.clerks;
  const unsigned char* const scoreMask = H.scoreMask; const char* const hb = H.hb; const int nG = H.nG, nClerk = H.nClerk; const bool active = H.active;
  const double* x = isVoid ? xsT : xsT + ((mat - 1) * nG + (G - 1)) * 6;
  const bool fissile = isVoid ? false : (fissileT[mat - 1] != 0);
  // ---- tallyAdmin%reportInColl: collisionClerks, then keffImplicitClerk (active cycles) ----
  // (skipped when no clerk of this phase can score a non-zero value in this material and group)
  const int nC = scoreMask[isVoid ? a.L.nMat * nG : (mat - 1) * nG + (G - 1)] ? nClerk : 0;
  // flux score w / Sigma_maj; majInv = 1 / Sigma_maj is that quotient for a particle of weight one
  const double flux = (nC || active) ? ((w == 1.0) ? majInv : w / majT[G - 1]) : 0.0;
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
  if (keff && active && !isVoid) {
    double nuf = fissile ? x[XS_NUFISSION] : 0.0, fis = fissile ? x[XS_FISSION] : 0.0;
    sProd += nuf * flux;
    sAbs += (x[XS_CAPTURE] + fis) * flux;
    nScore += 2;
  }
}

// ------------------------------------------------------------------------------------------------
// the loop of a history that has a warp to itself (speculative batches, see the top of the file): the body of k_lone.
//   mode 1: rounds run ahead in the material of the last collision, the geometry searched afterwards by the 32 lanes;
//   mode 2: the same rounds one after the other with the search in its place (histories that change material often, the
//           boundary, void, fission sites).
// Lane `owner` has the history; the other lanes use the same variables as scratch for the points they check.
// ------------------------------------------------------------------------------------------------
// the next WIN numbers of a history's stream behind state sb, WIN / 32 per lane, with everything the loops take from them: state, uniform,
// -log, sin / cos of the azimuth 2 pi xi, sin of the polar angle of mu = 2 xi - 1. All entries of the lane in one straight line (sb_math.h:
// log_main / sincos_main, the same values without branches) so that the logarithms, sines / cosines and square roots overlap; the special
// arguments after. Out of line: the loops that use the window stay compact (instruction caches).
template <int WIN>
__device__ __noinline__ void buildDrawWindow(DrawWinT<WIN>* Wp, const uint64_t sb, const JumpTab* jt) {
  DrawWinT<WIN>& W = *Wp;
  const int lane_ = (int)(threadIdx.x & 31);
  const ulonglong2 jm = jt->j[lane_];
  uint64_t st = (jm.x * sb + jm.y) & RNG_MASK;                             // lane + 1 draws ahead, then 32 more
  bool rare[WIN / 32];
#pragma unroll
  for (int k = 0; k < WIN / 32; ++k) {
    const int e = lane_ + 32 * k;
    const double xi = rngReal(st);
    double sn, cs;
    sbm::sincos_main(TWO_PI * xi, &sn, &cs);
    const double mu = 2.0 * xi - 1.0, a2 = fmax(0.0, 1.0 - mu * mu);
    bool rl;
    const double lg = sbm::log_main(xi, &rl);
    rare[k] = rl || !fastRange(a2);
    W.st[e] = st; W.xi[e] = xi; W.nlog[e] = -lg; W.sn[e] = sn; W.cs[e] = cs;
    W.A[e] = sqrtFast(a2);
    st = rngJump<32>(st);
  }
#pragma unroll
  for (int k = 0; k < WIN / 32; ++k)
    if (rare[k]) { const int e = lane_ + 32 * k; const double xi = W.xi[e]; W.nlog[e] = -sbm::log(xi); W.A[e] = sinPolar(2.0 * xi - 1.0); }
  __syncwarp();
}

struct __align__(16) LoneRec {             // a history handed from k_histories to k_lone
  double r0, r1, r2, u0, u1, u2, w, w0, sProd, sAbs, sScat;
  unsigned long long rng;
  int G, mat, hi, nSite, hSeg, leaked, pad0, pad1;
};
static_assert(sizeof(LoneRec) == 128, "LoneRec layout");
template <int WIN, bool LEG>
__device__ __forceinline__ void loneHistory(const HistArgs& a, const HotCtx& H, DrawWinT<WIN>& W, SpecRec* const rec, const JumpTab* const jt,
                                            const LoneRec& in, const int owner, unsigned& nSeg, unsigned& nColl, unsigned& nScore, long long* prL) {
  const unsigned FULL = 0xffffffffu;
#ifdef SB_PROFILE_ROUNDS
  long long pT = clock64(); const long long pT0 = pT; prL[10] += 1;
#define PL_MARK(j) { long long t_ = clock64(); prL[j] += t_ - pT; pT = t_; }
#else
#define PL_MARK(j)
#endif
  const int lane = (int)(threadIdx.x & 31);
  const double* const xsT = H.xsT; const double* const P0 = H.P0; const double* const prodT = H.prodT;
  const double* const majT = H.majT; const double* const majInvT = H.majInvT;
  const int* const p0First = H.p0First; const int* const fissileT = H.fissileT;
  const int nG = H.nG; const bool active = H.active; constexpr bool legendre = LEG;    // (P1 data run the copy with multiScatterP1MG)
  double r0 = in.r0, r1 = in.r1, r2 = in.r2, u0 = in.u0, u1 = in.u1, u2 = in.u2, w = in.w, sProd = in.sProd, sAbs = in.sAbs;
  const double w0 = in.w0;
  uint64_t rng = in.rng;
  int G = in.G, mat = in.mat, nSite = in.nSite, hSeg = in.hSeg;
  const int hi = in.hi;
  double majInv = majInvT[G - 1];
  bool alive = lane == owner, leaked = false;
  int winPos = WIN + 1;                                      // (the window is built behind rng)
  int specK = 2;                                             // rounds run ahead at once (> 0) or geometry searches in their place (0)
  if (lane == owner) H.scat[threadIdx.x] = in.sScat;
  for (;;) {
    // ---- the draw window: the next WIN numbers of the history's stream, four per lane ----
    if (__shfl_sync(FULL, winPos, owner) > WIN - 32) {
      buildDrawWindow<WIN>(&W, __shfl_sync(FULL, rng, owner), jt);
      if (lane == owner) winPos = 0;
    }
    PL_MARK(1)
    int mode = 2, K = SPEC_MAX;
    if (lane == owner && specK > 0 && mat >= 1 && mat <= a.L.nMat) { mode = 1; K = specK; }
    mode = __shfl_sync(FULL, mode, owner); K = __shfl_sync(FULL, K, owner);
    const int mat0 = __shfl_sync(FULL, mat, owner);
    const int hSeg0 = __shfl_sync(FULL, hSeg, owner);
    int nR = 0;
    // two copies of the rounds, one per mode: the one that runs ahead is compact enough for the instruction cache of its scheduler
    auto rounds = [&](auto seqTag) {
      constexpr bool seq = decltype(seqTag)::value;
      int p = winPos, nC = 0, m = mat0, sameRun = 0;               // sameRun: rounds in a row in one material
      double sScat = H.scat[threadIdx.x];
      bool ended = false, recorded = false;                        // ended: absorbed, leaked or lost; recorded: rec[nR] holds the history in front of round nR
      uint64_t sd = 0;                                             // the stream where it has left the window (many fission sites)
#pragma unroll 1
      for (int k = 0; k < K && !ended && p + 9 <= WIN; ++k) {      // a round in the window: flight 1, acceptance 1, channel 3, scattering 3 (P1: 4)
        SpecRec& R = rec[k];
        if (!seq) {
          R.b0 = r0; R.b1 = r1; R.b2 = r2; R.u0 = u0; R.u1 = u1; R.u2 = u2; R.w = w;
          R.sProd = sProd; R.sAbs = sAbs; R.sScat = sScat; R.G = G; R.p = p; R.nColl = nC;
        }
        const double distance = W.nlog[p] * majInv;
        r0 = r0 + distance * u0; r1 = r1 + distance * u1; r2 = r2 + distance * u2;
        if (seq) {                                                 // the search in its place
          hSeg = hSeg0 + k + 1;
          placePoint(a, H, true, hSeg, r0, r1, r2, u0, u1, u2, mat);
          sameRun = (mat == m) ? sameRun + 1 : 0;
          m = mat;
          nR = k + 1;
          if (m == SB_OUTSIDE_MAT) { leaked = true; ended = true; p += 1; break; }                   // LEAK_FATE
          if (m >= SB_OVERLAP_MAT && m != SB_VOID_MAT) {
            atomicMax(&a.cd->error, m == SB_UNDEF_MAT ? SB_ERR_UNDEF_MAT : SB_ERR_OVERLAP_MAT); ended = true; p += 1; break;
          }
          if (m == SB_VOID_MAT) { scoreColl(a, H, true, true, true, r0, r1, r2, mat, G, w, majInv, sProd, sAbs, nScore); p += 1; continue; }                      // no acceptance test in void
        } else { R.r0 = r0; R.r1 = r1; R.r2 = r2; }
        const bool fissile = fissileT[m - 1] != 0;
        const int row = (m - 1) * nG + (G - 1);
        const double* x = xsT + row * 6;
        const double tot = x[XS_TOTAL];
        const bool real = W.xi[p + 1] < (tot + 0.0) * majInv;
        int C = 1, nNew = 0, p3 = p + 4;
        if (real) {
          double xs = tot * W.xi[p + 3] - 0.0;                     // neutronMacroXSs%invert
          if (xs > 0.0) C += 1;
          xs = xs - x[XS_IESCATTER];
          if (xs > 0.0) C += 1;
          xs = xs - x[XS_CAPTURE];
          if (xs > 0.0) C += 1;
          if (fissile) {
            nNew = (int)(fabs((w * x[XS_NUFISSION]) / (w0 * tot * a.k_eff)) + W.xi[p + 4]);
            if (nNew < 0) nNew = 0;
            p3 = p + 5;
            if (nNew > 0 && !seq) { recorded = true; break; }      // fission sites: not in a round run ahead
          }
        }
        if (seq) scoreColl(a, H, false, !real, true, r0, r1, r2, mat, G, w, majInv, sProd, sAbs, nScore);                        // clerks and implicit k-eff scores
        else if (active) {                                         // keffImplicitClerk%reportInColl, as in scoreColl
          const double flux = (w == 1.0) ? majInv : divCold(w, majT[G - 1]);
          const double nuf = fissile ? x[XS_NUFISSION] : 0.0, fis = fissile ? x[XS_FISSION] : 0.0;
          sProd += nuf * flux;
          sAbs += (x[XS_CAPTURE] + fis) * flux;
        }
        if (!real) { R.flags = 0; p += 2; nR = k + 1; continue; }
        int fl = SR_REAL;
        nC += 1;
        if (nNew > 0) {                                            // (in its place) the sites, unfinished as in the loop above
          const int sb = atomicAdd(&a.cd->nSites, nNew);
          const bool room = sb + nNew <= a.cap;
          if (!room) atomicMax(&a.cd->error, SB_ERR_BANK_OVERFLOW);
          const double wSite = fsign(w0, w);
          sd = W.st[p3 - 1];
#pragma unroll 1
          for (int i = 0; i < nNew; ++i) {
            if (room) {
              const int sI = sb + i;
              a.out.rx[sI] = r0; a.out.ry[sI] = r1; a.out.rz[sI] = r2;
              a.out.ux[sI] = u0; a.out.uy[sI] = u1; a.out.uz[sI] = u2;
              a.out.w[sI] = wSite; a.out.G[sI] = m; a.out.brood[sI] = hi; a.out.seq[sI] = nSite + i;
              a.out.E[sI] = __longlong_as_double((long long)sd);
            }
            sd = rngJump<3>(sd);
          }
          nSite += nNew;
          p3 += 3 * nNew;
        }
        if (C == 2) {                                              // multiScatterMG%sampleOut, rotateVector, neutronMGstd inelastic
          double rem, mu = 0.0, A = 0.0, sn = 0.0, cs = 1.0;
          const bool inWin = p3 + (legendre ? 4 : 3) <= WIN;       // the numbers of the scattering are in the window
          int used = 3;
          if (legendre) {                                          // multiScatterP1MG: G_out first, then mu from the P1 coefficient, then phi
            if (inWin) rem = W.xi[p3]; else rem = rngGet(sd);
          } else if (inWin) { rem = W.xi[p3]; mu = 2.0 * W.xi[p3 + 1] - 1.0; A = W.A[p3 + 1]; sn = W.sn[p3 + 2]; cs = W.cs[p3 + 2]; }
          else {                                                   // (many sites) the three numbers straight from the stream
            const uint64_t s1 = rngJump<1>(sd), s2 = rngJump<2>(sd), s3 = rngJump<3>(sd);
            rem = rngReal(s1); mu = 2.0 * rngReal(s2) - 1.0; sbm::sincos(TWO_PI * rngReal(s3), &sn, &cs); A = sinPolar(mu);
            sd = s3;
          }
          rem = rem * x[XS_IESCATTER];
          const double* cdf = P0 + row * nG;
          int Gout = 0;
#pragma unroll 1
          for (int g0 = p0First[row]; g0 < nG && Gout == 0; g0 += 4) {
            const double c0 = cdf[g0], c1 = (g0 + 1 < nG) ? cdf[g0 + 1] : 0.0, c2 = (g0 + 2 < nG) ? cdf[g0 + 2] : 0.0, c3 = (g0 + 3 < nG) ? cdf[g0 + 3] : 0.0;
            const double e1 = rem - c0, e2 = e1 - c1, e3 = e2 - c2, e4 = e3 - c3;
            if (e1 < 0.0) Gout = g0 + 1;
            else if (e2 < 0.0) Gout = g0 + 2;
            else if (e3 < 0.0) Gout = g0 + 3;
            else if (e4 < 0.0) Gout = g0 + 4;
            rem = e4;
          }
          if (Gout == 0) { fl |= SR_SAMPLING; Gout = G; if (seq) atomicMax(&a.cd->error, SB_ERR_SAMPLING); }
          if (legendre) {                                          // sampleLegendre_P1 (legendrePoly_func.f90:35-93) on the window's numbers
            const double P1v = H.P1[row * nG + (Gout - 1)];
            if (inWin) {
              const double P1_loc = fabs(P1v);
              double threshold; int Low, Top;                      // 1 UNIFORM, 2 LIN, 3 DELTA
              if (P1_loc < 1.0) { threshold = P1_loc; Top = 2; Low = 1; }
              else { threshold = 0.5 * (P1_loc - 1.0); Top = 3; Low = 2; }
              int q = p3 + 1;
              const int exec = (W.xi[q] < threshold) ? Top : Low; ++q;
              double xx;
              if (exec == 1) { xx = 2.0 * W.xi[q] - 1.0; ++q; }
              else if (exec == 2) { xx = 2.0 * sqrt(W.xi[q]) - 1.0; ++q; }
              else xx = 1.0;
              if (P1v < 0.0) xx = -xx;
              mu = xx; sn = W.sn[q]; cs = W.cs[q]; ++q;
              used = q - p3;
            } else {
              mu = sampleLegendreP1(P1v, sd);
              sincosHot(TWO_PI * rngGet(sd), &sn, &cs);
            }
            A = sinPolar(mu);
          }
          if (!inWin) used = WIN + 1 - p3;                         // (the stream has left the window: p > WIN below)
          rotateVectorHot(u0, u1, u2, mu, sn, cs, A);
          const double w_mul = prodT[row * nG + (Gout - 1)];
          const double wPre = w;
          if (Gout != G || w_mul != 1.0) { G = Gout; majInv = majInvT[G - 1]; w = w * w_mul; }
          const double sc = fmax(w - wPre, 0.0);
          if (sc > 0.0) sScat += sc;
          p = p3 + used;
        } else {
          p = p3;
          if (C == 3 || C == 4) { fl |= SR_DIED; ended = true; }  // capture / fission: the history ends (ABS_FATE)
        }
        R.flags = fl;
        nR = k + 1;
      }
      if (seq) {                                                   // the rounds are made
        H.scat[threadIdx.x] = sScat;
        nColl += (unsigned)nC;
        if (p <= WIN) { if (p > 0) rng = W.st[p - 1]; } else rng = sd;          // (p > WIN: the window is rebuilt behind the stream)
        winPos = p;
        hSeg = hSeg0 + nR;
        if (ended) alive = false;
        specK = (sameRun >= 4) ? 4 : 0;                            // the history has stayed in one material for a while: run ahead from now on
      } else if (!recorded) {                                      // the history behind the last round run ahead
        SpecRec& R = rec[nR];
        R.b0 = r0; R.b1 = r1; R.b2 = r2; R.u0 = u0; R.u1 = u1; R.u2 = u2; R.w = w;
        R.sProd = sProd; R.sAbs = sAbs; R.sScat = sScat; R.G = G; R.p = p; R.nColl = nC; R.flags = 0;
      }
    };
    if (lane == owner) { if (mode == 2) rounds(std::true_type{}); else rounds(std::false_type{}); }
    nR = __shfl_sync(FULL, nR, owner);
    __syncwarp();
#ifdef SB_PROFILE_ROUNDS
    if (mode == 1) { PL_MARK(2) prL[7] += 1; prL[9] += nR; } else { PL_MARK(4) prL[8] += 1; prL[6] += nR; }
#endif
    if (mode == 1) {
      // ---- the geometry searches of the recorded collision points, one per lane ----
      bool ok = true; int found = -1;
      if (lane < nR) {
        const SpecRec& R = rec[lane];
        r0 = R.r0; r1 = R.r1; r2 = R.r2; u0 = R.u0; u1 = R.u1; u2 = R.u2;
        hSeg = hSeg0 + lane + 1;                                   // (the cell cache's age gate)
        const bool tele = placePoint(a, H, true, hSeg, r0, r1, r2, u0, u1, u2, mat);
        found = tele ? -1 : mat;
        ok = found == mat0;
      }
      const unsigned bad = __ballot_sync(FULL, !ok);
      const int keep = bad ? __ffs(bad) - 1 : nR;                  // rounds in front of the first point in another material
      const int matNext = __shfl_sync(FULL, found, keep & 31);     // (keep < nR) the material of that point
      // ---- clerk scores of the rounds that are kept, by the lanes that checked them ----
      if (lane < keep) {
        const SpecRec& R = rec[lane];
        G = R.G; w = R.w; majInv = majInvT[G - 1];
        scoreColl(a, H, false, !(R.flags & SR_REAL), false, r0, r1, r2, mat, G, w, majInv, sProd, sAbs, nScore);
        if (R.flags & SR_SAMPLING) atomicMax(&a.cd->error, SB_ERR_SAMPLING);
      }
      __syncwarp();
      // ---- the history goes on in front of the first round that is not kept ----
      if (lane == owner) {
        const SpecRec& R = rec[keep];
        r0 = R.b0; r1 = R.b1; r2 = R.b2; u0 = R.u0; u1 = R.u1; u2 = R.u2; w = R.w;
        sProd = R.sProd; sAbs = R.sAbs; H.scat[threadIdx.x] = R.sScat; G = R.G; majInv = majInvT[G - 1];
        winPos = R.p; nColl += (unsigned)R.nColl; if (active) nScore += 2u * (unsigned)keep;
        if (winPos > 0) rng = W.st[winPos - 1];
        hSeg = hSeg0 + keep;
        mat = (keep < nR && matNext >= 1) ? matNext : ((keep < nR) ? 0 : mat0);   // (beyond the boundary / outside / void: the search in its place takes the round)
        if (keep > 0 && (rec[keep - 1].flags & SR_DIED)) alive = false;         // absorbed in the last round kept: the history ends (ABS_FATE)
        if (keep == nR) { if (nR == K) specK = min(2 * K, SPEC_MAX); else if (nR == 0) specK = 0; }   // every point in the material (no round run: fission sites)
        else if (keep <= 1) specK = 0;                             // the history changes material often: searches in their place
        else specK = max(2, keep);
      }
      __syncwarp();
#ifdef SB_PROFILE_ROUNDS
      PL_MARK(3) prL[5] += keep;
#endif
    }
    if (!__shfl_sync(FULL, (int)alive, owner)) break;
  }
  if (lane == owner) {                                       // the scores of the history
    a.nsites[hi] = nSite;
    a.hProd[hi] = sProd; a.hAbs[hi] = sAbs; a.hLeak[hi] = leaked ? 0.0 + w : 0.0; a.hScat[hi] = H.scat[threadIdx.x];
    nSeg += (unsigned)hSeg;
    if (hSeg > a.maxSegMin) atomicMax(&a.cd->maxSeg, hSeg);
  }
  __syncwarp();
#ifdef SB_PROFILE_ROUNDS
  prL[0] += clock64() - pT0;
#endif
#undef PL_MARK
}

// dynamic shared memory of k_histories beside the hot blob, per thread count
__host__ __device__ constexpr int histScratchBytes(int threads) {
  return (int)sizeof(DrawWinT<WIN_SMALL>) * (threads / 32) + threads * (3 * 16 + 3 * 16 + 16 + 16 + 8) + (int)sizeof(JumpTab);
}
// k_lone: per warp a draw window of WIN_BIG numbers and the records of a batch; per thread the cell cache and the scatter score
__host__ __device__ constexpr int loneScratchBytes(int threads) {
  return (int)(sizeof(DrawWinT<WIN_BIG>) + sizeof(SpecRec) * (SPEC_MAX + 1)) * (threads / 32) + threads * (3 * 16 + 3 * 16 + 16 + 16 + 8) + (int)sizeof(JumpTab);
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

// LW: the loop has the draw window of a history that is alone in its warp. The launches that hand their last histories to k_lone
// (HistArgs::assist > 0) never have such a history and run the copy without it (fewer tests per draw, a smaller loop).
// LEG: the loop has multiScatterP1MG; P0 data run the copy without it (k_lone<THREADS, LEG> likewise).
template <bool SMEM, int BPS, int THREADS = 256, bool LW = true, bool LEG = true>
__global__ void __launch_bounds__(THREADS, BPS) k_histories(const __grid_constant__ HistArgs a) {
  __shared__ __align__(8) uint64_t s_bar;
  constexpr int NW = THREADS / 32;
  constexpr int WIN = WIN_SMALL;
  using DrawWin = DrawWinT<WIN>;
  // dynamic shared memory: [hot blob (SMEM)] [draw windows] [cell caches] [scatter scores] [jump table]
  char* const sx = g_hotSmem + (SMEM ? a.L.bytes : 0);
  DrawWin* const s_win = (DrawWin*)sx;                                          // one per warp
  double2 (* const s_gcB)[THREADS] = (double2 (*)[THREADS])(sx + sizeof(DrawWin) * NW);   // cell cache: safe box {lo, hi} per axis
  double2 (* const s_gcO)[THREADS] = s_gcB + 3;                                 // offsets of the outer (A) and inner (B) lattice level, {x, y}; origin of the cached universe
  int4* const s_gcI = (int4*)(s_gcO + 3);                                       // universe to resume at, its rootID, its level
  int4* const s_gcP = s_gcI + THREADS;                                          // if that universe is a plain pin: position of its radii, their number, has-origin flag (else -1)
  double* const s_scat = (double*)(s_gcP + THREADS);                            // keffImplicitClerk%reportOutColl score of the history (non-zero only with multiplicities)
  JumpTab* const s_jump = (JumpTab*)(s_scat + THREADS);
  if (threadIdx.x < 32) s_jump->j[threadIdx.x] = __ldg(a.seedTab + 3 * 1024 + threadIdx.x);   // (ordered before the loop by the barrier below)
  if (a.assist > 0) asm volatile("griddepcontrol.launch_dependents;");          // k_lone may take the SMs this kernel's CTAs leave
  const char* hb;
  if (SMEM) { stageHot(g_hotSmem, a.hot, a.L.bytes, &s_bar); hb = g_hotSmem; }
  else { hb = a.hot; __syncthreads(); }                                          // (the jump table above)
  const HUni* const uni = (const HUni*)(hb + a.L.oUni);
  const int2* const graph = (const int2*)(hb + a.L.oGraph);
  const double* const auxD = (const double*)(hb + a.L.oAuxD);
  const int* const auxI = (const int*)(hb + a.L.oAuxI);
  const double* const xsT = (const double*)(hb + a.L.oXs);
  const double* const P0 = (const double*)(hb + a.L.oP0);
  const double* const prodT = (const double*)(hb + a.L.oProd);
  const double* const P1 = (const double*)(hb + a.L.oP1);
  const int* const p0First = (const int*)(hb + a.L.oP0First);           // per (material, group in): first non-zero term of the P0 row
  const int* const fissileT = (const int*)(hb + a.L.oFissile);
  const double* const majT = (const double*)(hb + a.L.oMajT);
  const double* const majInvT = (const double*)(hb + a.L.oMajInv);
  const DClerk* const clerks = (const DClerk*)(hb + a.L.oClerk[0]);     // phase offset applied by the host (oClerk[0] = this launch)
  const int nClerk = a.L.nClerk[0];
  const unsigned char* const scoreMask = (const unsigned char*)(hb + a.L.oScoreMask[0]);
  const int nG = a.L.nG;
  const bool active = a.impScores != 0;
  const HotCtx H = makeHotCtx(a, hb, (char*)&s_gcB[0][0], THREADS);

  const unsigned FULL = 0xffffffffu;
#define lane ((int)(threadIdx.x & 31))
#define ltMask ((1u << lane) - 1u)
  DrawWin& W = s_win[threadIdx.x >> 5];

  bool alive = false, exhausted = false;
  int hi = -1, G = 1, mat = 0, nSite = 0, hSeg = 0;
  int winPos = 9999;                                 // < WIN only while this lane's draws come from the window
  double r0 = 0.0, r1 = 0.0, r2 = 0.0, u0 = 1.0, u1 = 0.0, u2 = 0.0;
  double w = 0.0, w0 = 0.0, majInv = 1.0;
  uint64_t rng = 0;
  double sProd = 0.0, sAbs = 0.0;                  // implicit k-eff scores of the history
  bool leaked = false;                             // (a delta-tracking history of an eigenvalue cycle leaks at most once: its weight at that point)
  unsigned nSeg = 0, nColl = 0, nScore = 0;        // per lane, over all its histories

  // leave the window for the rest of the round (code that draws through rng_get directly)
  auto leaveWindow = [&]() {
    if (LW && winPos <= WIN) { if (winPos > 0) rng = W.st[winPos - 1]; winPos = WIN + 1; }
  };
  // fission sites: the stream position is needed as a state; the window is left (its next rebuild starts behind the sites)
  auto leaveWindowKeep = leaveWindow;

  auto place = [&]() -> bool { return placePoint(a, H, LW && winPos <= WIN, hSeg, r0, r1, r2, u0, u1, u2, mat); };
  auto score = [&](const bool isVoid, const bool virt, const bool keff) { scoreColl(a, H, isVoid, virt, keff, r0, r1, r2, mat, G, w, majInv, sProd, sAbs, nScore); };

#ifdef SB_PROFILE_ROUNDS
  // per LANE: cycles per region of the rounds in which the lane's history was alone in the warp with its window [0..7],
  // alone without [8..15], with 1 - 3 others [16..23]; rounds of each kind [24..26]
  long long prR[36]; for (int i = 0; i < 36; ++i) prR[i] = 0;
  long long prS = clock64(); int prB = -1;
#define PR_MARK(j) { long long t = clock64(); if (alive && prB >= 0) prR[8 * prB + (j)] += t - prS; prS = t; }
#else
#define PR_MARK(j)
#endif
  for (;;) {
    PR_MARK(7)
    // ---------------- refill dead lanes (warp-level compaction of the bank) ----------------------------
    unsigned need = __ballot_sync(FULL, !alive);
    if ((need & a.laneMask) != 0u && !exhausted) {
      const unsigned take = need & a.laneMask;             // lanes that may be refilled
      int cnt = __popc(take);
      if (cnt >= a.refillMin || need == FULL) {
        int b = 0;
        if (lane == 0) b = atomicAdd(&a.cd->nextHistory, cnt);
        b = __shfl_sync(FULL, b, 0);
        if (b + cnt >= a.n) exhausted = true;
        int my = b + __popc(take & ltMask);
        if (!alive && ((take >> lane) & 1u) && my < a.n) {
          hi = my;
          r0 = a.in.rx[hi]; r1 = a.in.ry[hi]; r2 = a.in.rz[hi];
          u0 = a.in.ux[hi]; u1 = a.in.uy[hi]; u2 = a.in.uz[hi];
          w = a.in.w[hi]; w0 = w; G = a.in.G[hi];
          rng = rngSeed(a.seedTab, a.rng0, (unsigned)(a.histOffset + hi + 1));
          // geom%placeCoord of the source site is not needed by delta tracking: the first thing the
          // flight does is teleport + placeCoord (transportOperatorDT_class.f90:57-75)
          majInv = majInvT[G - 1];
          nSite = 0; hSeg = 0; sProd = 0.0; sAbs = 0.0; leaked = false; s_scat[threadIdx.x] = 0.0;
          s_gcB[0][threadIdx.x] = make_double2(INF, -INF);           // no cached cell yet
          alive = true;
        }
        need = __ballot_sync(FULL, !alive);
      }
    }
    if (need == FULL && exhausted) break;
    // ---------------- the bank is exhausted and few histories are left in the warp: they go on in k_lone, one warp each ----------
    if (a.assist > 0 && exhausted && __popc(~need) <= a.assist) {
      const unsigned live = ~need;
      int b = 0;
      if (lane == 0) b = atomicAdd(a.loneCount, __popc(live));
      b = __shfl_sync(FULL, b, 0);
      if (alive) {
        const int q = b + __popc(live & ltMask);
        if (q < a.loneCap) {
          leaveWindow();                                   // (rng = the position of the stream)
          LoneRec R;
          R.r0 = r0; R.r1 = r1; R.r2 = r2; R.u0 = u0; R.u1 = u1; R.u2 = u2; R.w = w; R.w0 = w0; R.sProd = sProd; R.sAbs = sAbs;
          R.sScat = s_scat[threadIdx.x]; R.rng = rng; R.G = G; R.mat = mat; R.hi = hi; R.nSite = nSite; R.hSeg = hSeg; R.leaked = 0; R.pad0 = 0; R.pad1 = 0;
          a.loneQ[q] = R;
          __threadfence();                                 // k_lone runs beside this kernel: the record, then its flag
          *(volatile int*)&a.loneReady[q] = a.loneTag;
        } else atomicMax(&a.cd->error, SB_ERR_BANK_OVERFLOW);     // (the queue has one slot per resident lane: cannot happen)
        alive = false;
      }
      break;
    }

    // ---------------- a history alone in its warp: (re)build its draw window with all 32 lanes ---------
    if (LW && exhausted && a.loneMode && __popc(~need) == 1) {
      const int owner = __ffs(~need) - 1;
      if (__shfl_sync(FULL, winPos, owner) > WIN - WIN_ROUND) {
        buildDrawWindow<WIN>(&W, __shfl_sync(FULL, rng, owner), s_jump);
        if (lane == owner) winPos = 0;
      }
    }
#ifdef SB_PROFILE_ROUNDS
    { const int na = __popc(~need); prB = (na == 1) ? (winPos <= WIN ? 0 : 1) : (na <= 4 ? 2 : -1); if (alive && prB >= 0) prR[24 + prB] += 1; }
#endif
    PR_MARK(0)
    // ---------------- event: tentative flight = move + cell search + virtual/real decision -------------
    bool realColl = false, died = false;
    if (alive) {
      {
        double nlog;
        if (LW && winPos < WIN - 1) { nlog = W.nlog[winPos]; winPos += 1; }
        else {
          leaveWindow();
          rng = rngJump<1>(rng);
          nlog = negLogHot(rngReal(rng));
        }
        const double distance = nlog * majInv;
        r0 = r0 + distance * u0; r1 = r1 + distance * u1; r2 = r2 + distance * u2;
        ++hSeg;
      }
      PR_MARK(1)
      const bool tele = place(); (void)tele;
      PR_MARK(2)
      if (mat == SB_OUTSIDE_MAT) { leaked = true; died = true; }                          // LEAK_FATE
      else if (mat >= SB_OVERLAP_MAT && mat != SB_VOID_MAT) {
        atomicMax(&a.cd->error, mat == SB_UNDEF_MAT ? SB_ERR_UNDEF_MAT : SB_ERR_OVERLAP_MAT); died = true;
      } else {
        const bool isVoid = (mat == SB_VOID_MAT);
        bool virt = true;
        if (!isVoid) {
          const double* x = xsT + ((mat - 1) * nG + (G - 1)) * 6;
          double sigmaT = x[XS_TOTAL] + 0.0;
          double xiAcc;                                                        // the acceptance test draws (not in void)
          if (LW && winPos < WIN) { xiAcc = W.xi[winPos]; winPos += 1; } else { leaveWindow(); rng = rngJump<1>(rng); xiAcc = rngReal(rng); }
          if (xiAcc < sigmaT * majInv) { realColl = true; virt = false; }
        }
        score(isVoid, virt, true);
      }
    }

    PR_MARK(3)
    // ---------------- event: collision, part 1 (channel + number of fission sites) --------------------
    int MT = 0, nNew = 0;
    if (realColl) {
      const double* x = xsT + ((mat - 1) * nG + (G - 1)) * 6;
      const bool fissile = fissileT[mat - 1] != 0;
      double rr, rand1 = 0.0;                             // the alpha-absorption test always draws first (probAlpha = 0)
      if (LW && winPos <= WIN - 3) { rr = W.xi[winPos + 1]; rand1 = W.xi[winPos + 2]; winPos += fissile ? 3 : 2; }
      else {
        leaveWindow();
        const uint64_t s2 = rngJump<2>(rng), s3 = rngJump<3>(rng);
        rr = rngReal(s2); rand1 = rngReal(s3);
        rng = fissile ? s3 : s2;
      }
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
      if (fissile) {                                      // neutronMGstd implicit (:131-199)
        nNew = (int)(fabs((w * x[XS_NUFISSION]) / (w0 * x[XS_TOTAL] * a.k_eff)) + rand1);
        if (nNew < 0) nNew = 0;
      }
    }

    PR_MARK(4)
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

    PR_MARK(5)
    // ---------------- collision, part 2: fission sites, then the scattered neutron --------------------
    // The sites are written UNFINISHED: position, the parent's direction, the material, and the state of the history's
    // stream in front of the site's three numbers. fissionMG%sampleOut (mu, phi, chi walk) and rotateVector are applied
    // to all sites of the cycle at once by k_finish_sites, with the same arithmetic on the same numbers - they do not
    // feed back into the history, so they leave its dependent chain (and the chains of the other lanes of the warp).
    if (realColl) {
      if (nNew > 0) {
        const double wSite = fsign(w0, w);
        leaveWindowKeep();                                // rng = the stream position in front of the sites
#pragma unroll 1
        for (int i = 0; i < nNew; ++i) {
          if (slot >= 0) {
            const int s = slot + i;
            a.out.rx[s] = r0; a.out.ry[s] = r1; a.out.rz[s] = r2;
            a.out.ux[s] = u0; a.out.uy[s] = u1; a.out.uz[s] = u2;
            a.out.w[s] = wSite; a.out.G[s] = mat; a.out.brood[s] = hi; a.out.seq[s] = nSite + i;
            a.out.E[s] = __longlong_as_double((long long)rng);
          }
          rng = rngJump<3>(rng);
        }
        nSite += nNew;
      }
      if (MT == 2) {                                 // multiScatterMG%sampleOut: G_out, then mu, phi
        const int row = (mat - 1) * nG + (G - 1);
        const double* cdf = P0 + row * nG;
        double mu, A, sn, cs, rem;
        const bool legendre = LEG && a.L.isP1 != 0;         // (LEG = false: a copy of the loop for P0 data without the P1 code)
        if (LW && winPos <= WIN - 3 && !legendre) {              // three numbers in the order of the reaction, all from the window
          rem = W.xi[winPos]; mu = 2.0 * W.xi[winPos + 1] - 1.0; A = W.A[winPos + 1]; sn = W.sn[winPos + 2]; cs = W.cs[winPos + 2];
          winPos += 3;
        } else if (!legendre) {
          leaveWindow();
          const uint64_t s1 = rngJump<1>(rng), s2 = rngJump<2>(rng), s3 = rngJump<3>(rng);
          rng = s3;
          rem = rngReal(s1);
          mu = 2.0 * rngReal(s2) - 1.0;
          sbm::sincos_main(TWO_PI * rngReal(s3), &sn, &cs);       // (as the draw windows: the same values for 0 <= x < 2 pi)
          A = sinPolar(mu);
        } else { leaveWindow(); rem = rngGet(rng); mu = 0.0; A = 0.0; sn = 0.0; cs = 1.0; }
        rem = rem * xsT[row * 6 + XS_IESCATTER];
        int Gout = 0;
#pragma unroll 1
        for (int g0 = p0First[row]; g0 < nG && Gout == 0; g0 += 4) {   // rem = rem - cdf(g) in the reference's order, four terms per trip;
          // leading zero terms are skipped (x - 0 = x) and a zero term past nG leaves a non-negative remainder non-negative
          const double c0 = cdf[g0], c1 = (g0 + 1 < nG) ? cdf[g0 + 1] : 0.0, c2 = (g0 + 2 < nG) ? cdf[g0 + 2] : 0.0, c3 = (g0 + 3 < nG) ? cdf[g0 + 3] : 0.0;
          const double e1 = rem - c0, e2 = e1 - c1, e3 = e2 - c2, e4 = e3 - c3;
          if (e1 < 0.0) Gout = g0 + 1;
          else if (e2 < 0.0) Gout = g0 + 2;
          else if (e3 < 0.0) Gout = g0 + 3;
          else if (e4 < 0.0) Gout = g0 + 4;
          rem = e4;
        }
        if (Gout == 0) { atomicMax(&a.cd->error, SB_ERR_SAMPLING); Gout = G; }
        if (legendre) {
          mu = sampleLegendreP1(P1[row * nG + (Gout - 1)], rng);
          sbm::sincos(TWO_PI * rngGet(rng), &sn, &cs);
          A = sinPolar(mu);
        }
        PR_MARK(6)
        double d[3] = {u0, u1, u2};
        rotateVectorHot(d[0], d[1], d[2], mu, sn, cs, A);
        {                                                   // neutronMGstd inelastic (:221-252)
          double w_mul = prodT[row * nG + (Gout - 1)];
          double wPre = w;
          if (Gout != G || w_mul != 1.0) {
            G = Gout;
            majInv = majInvT[G - 1];
            w = w * w_mul;
          }
          u0 = d[0]; u1 = d[1]; u2 = d[2];
          double sc = fmax(w - wPre, 0.0);                   // keffImplicitClerk%reportOutColl
          if (sc > 0.0) s_scat[threadIdx.x] += sc;
        }
      }
      if (MT == 3 || MT == 4) died = true;                   // capture / fission: history ends (ABS_FATE)
      // MT == 1 (elastic) cannot be selected for MG data (elasticScatter = 0): "Do nothing"
    }
    PR_MARK(7)
    if (LW && winPos <= WIN && winPos > 0) rng = W.st[winPos - 1];  // the stream position after this round's draws from the window
    if (died) {
      a.nsites[hi] = nSite;
      a.hProd[hi] = sProd; a.hAbs[hi] = sAbs; a.hLeak[hi] = leaked ? 0.0 + w : 0.0; a.hScat[hi] = s_scat[threadIdx.x];
      nSeg += hSeg;
      if (hSeg > a.maxSegMin) atomicMax(&a.cd->maxSeg, hSeg);
      alive = false; winPos = 9999;
    }
  }

#ifdef SB_PROFILE_ROUNDS
  if (a.prof) { long long* o = a.prof + 36 * ((long long)blockIdx.x * THREADS + threadIdx.x); for (int i = 0; i < 36; ++i) o[i] = prR[i]; }
#endif
  // ---------------- per-warp event counters (integers: order-independent) -----------------------------
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    nSeg += __shfl_down_sync(FULL, nSeg, d); nColl += __shfl_down_sync(FULL, nColl, d); nScore += __shfl_down_sync(FULL, nScore, d);
  }
  if (lane == 0) {
    atomicAdd(&a.cd->nSeg, (unsigned long long)nSeg); atomicAdd(&a.cd->nColl, (unsigned long long)nColl);
    atomicAdd(&a.cd->nScore, (unsigned long long)nScore);
    if (a.assist > 0 && lane == 0) { __threadfence(); atomicAdd(a.loneDone, 1); }      // (behind the records this warp has queued)
}
}


// ------------------------------------------------------------------------------------------------
// k_lone: the histories k_histories handed over, one warp each (loneHistory). Launched behind k_histories on the same stream
// with programmatic dependent launch: the tables are staged while the last warps of k_histories finish.
// ------------------------------------------------------------------------------------------------
template <int THREADS, bool LEG>
__global__ void __launch_bounds__(THREADS, 2) k_lone(const __grid_constant__ HistArgs a) {
  __shared__ __align__(8) uint64_t s_bar;
  constexpr int NW = THREADS / 32;
  using DrawWin = DrawWinT<WIN_BIG>;
  char* const sx = g_hotSmem + a.L.bytes;
  DrawWin* const s_win = (DrawWin*)sx;
  SpecRec* const s_rec = (SpecRec*)(s_win + NW);
  char* const cacheBase = (char*)(s_rec + NW * (SPEC_MAX + 1));
  JumpTab* const s_jump = (JumpTab*)(cacheBase + THREADS * (3 * 16 + 3 * 16 + 16 + 16 + 8));
  asm volatile("griddepcontrol.launch_dependents;");
  if (threadIdx.x < 32) s_jump->j[threadIdx.x] = __ldg(a.seedTab + 3 * 1024 + threadIdx.x);
  stageHot(g_hotSmem, a.hot, a.L.bytes, &s_bar);           // (the hot blob is not written by the kernel in front)
  const HotCtx H = makeHotCtx(a, g_hotSmem, cacheBase, THREADS);
  H.gcB[0][threadIdx.x] = make_double2(INF, -INF);         // no cached cell yet
  // no griddepcontrol.wait: the queue is consumed while k_histories fills it; this kernel ends when every warp of k_histories has ended
  const unsigned FULL = 0xffffffffu;
  unsigned nSeg = 0, nColl = 0, nScore = 0;
  long long prL[12]; for (int i = 0; i < 12; ++i) prL[i] = 0;
  for (;;) {
    int q = 0;
    if (lane == 0) q = atomicAdd(a.loneNext, 1);
    q = __shfl_sync(FULL, q, 0);
    int st = 0;                                             // 1: the record of ticket q is there; 2: k_histories has ended without filling it
    if (lane == 0) {
      for (;;) {
        if (q < a.loneCap && *(volatile int*)&a.loneReady[q] == a.loneTag) { st = 1; break; }
        if (*(volatile int*)a.loneDone == a.loneWarps) {
          __threadfence();
          st = (q < a.loneCap && *(volatile int*)&a.loneReady[q] == a.loneTag) ? 1 : 2;
          break;
        }
        __nanosleep(256);
      }
    }
    st = __shfl_sync(FULL, st, 0);
    if (st == 2) break;
    __threadfence();
    LoneRec in;
    {                                                       // (every lane reads the record past L1: it was written by a kernel that is still running)
      const int4* src = (const int4*)(a.loneQ + q); int4* dst = (int4*)&in;
#pragma unroll
      for (int i = 0; i < (int)(sizeof(LoneRec) / 16); ++i) dst[i] = __ldcg(src + i);
    }
    loneHistory<WIN_BIG, LEG>(a, H, s_win[threadIdx.x >> 5], s_rec + (threadIdx.x >> 5) * (SPEC_MAX + 1), s_jump, in, 0, nSeg, nColl, nScore, prL);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    nSeg += __shfl_down_sync(FULL, nSeg, d); nColl += __shfl_down_sync(FULL, nColl, d); nScore += __shfl_down_sync(FULL, nScore, d);
  }
  if (lane == 0 && (nSeg | nColl | nScore)) {
    atomicAdd(&a.cd->nSeg, (unsigned long long)nSeg); atomicAdd(&a.cd->nColl, (unsigned long long)nColl);
    atomicAdd(&a.cd->nScore, (unsigned long long)nScore);
  }
#ifdef SB_PROFILE_ROUNDS
  if (a.prof && lane == 0) { long long* o = a.prof + 36LL * 148 * 384 + 12 * ((long long)blockIdx.x * NW + (threadIdx.x >> 5)); for (int i = 0; i < 12; ++i) o[i] = prL[i]; }
#endif
}

#undef lane
#undef ltMask
}  // namespace sbh
