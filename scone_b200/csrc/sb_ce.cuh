// Continuous-energy cross-section lookup on the device (the XS-lookup event kernel of the CE path).
//
//   aceNeutronNuclide%search / totalXS / microXSs   aceDatabase/aceNeutronNuclide_class.f90:342-453
//   aceNeutronDatabase%updateTotalMatXS / updateMacroXSs / updateMajorantXS / initMajorant
//                                                   aceDatabase/aceNeutronDatabase_class.f90:346-394,509-644,1330-1621
//   binarySearch                                    SharedModules/genericProcedures.f90:132-166
//
// The reference does one binary search per nuclide per lookup (~12 dependent loads for a 3667-point grid).
// Here the searches are replaced by ONE search on the unionised grid of all nuclides plus a table
// idxTab[union interval][nuclide] that holds, for every nuclide, the index binarySearch would return for any energy
// inside that union interval (no nuclide grid point lies strictly inside a union interval, so the index is exact;
// tests compare it bit for bit).  The union search itself is hashed on the IEEE bits of E (exponent + 9 mantissa
// bits: monotone in E, no log), then walks at most the few points of one bucket.
// Interpolation and summation are the reference's, in the reference's order (nuclides in material order), so the
// macroscopic cross sections are bit-identical to the CPU restatement, not just within 1e-12.
//
// Layout: nuclide grids concatenated (f64); main data as the reference holds it, mainData(rows, N) with the rows of
// one energy point contiguous (rows = 4 or 8), so the two points an interpolation needs are one contiguous
// 64 B / 128 B segment; idxTab row-major by union interval (one coalesced row per lookup); materials in CSR.
// Algorithmic bytes per lookup: total only: 36 B per nuclide + 20 B; full macro set:
// 4 + 16 + 2*8*rows B per nuclide + 12 + 64 B.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/scone_b200.h"

namespace sbce {

constexpr int HASH_MBITS_MIN = 9, HASH_MBITS_MAX = 22;      // mantissa bits in the bucket key of the union grid: about one point per bucket

struct CeDev {                                    // device pointers + sizes, passed by value
  int nNuc, nMat, nUnion, nBuckets, uShift;
  long long keyMin;
  const double* grid; const double* data; const long long* gridOff; const long long* dataOff; const int* rows; const int* gridSize;
  const int* matOff; const int* matNuc; const double* matDens;
  const double* uGrid; const double* uMaj; const int* idxTab; const int* bucketStart;
  const double* pairTot; const long long* pairOff;      // per nuclide and grid interval: { E_low, E_top, total_low, total_top }, 32-byte aligned
  const int* activeMat; int nActive;                    // materials the majorant covers (nuclearDatabase%activate: those present in the geometry)
  double eMin, eMax;
  // per-nuclide hashed index (always built; the only index when idxTab == nullptr, i.e. for libraries whose
  // [union interval][nuclide] table would not fit): nuclide n keys its grid on the IEEE bits of E shifted by nbShift[n]
  // (about one grid point per bucket), nbStart[nbOff[n] + b] = number of its grid points whose key is below bucket b
  const int* nbStart; const long long* nbOff; const int* nbShift; const long long* nbKeyMin; const int* nbCount;
  const struct NucMeta* nbMeta;                         // the same per nuclide in one 32-byte record (one load per nuclide and lookup)
};
struct __align__(32) NucMeta { long long keyMin, nbOff, pairOff; int shift, last; };      // last = index of the last pair record (N - 2)

__host__ __device__ inline long long hashKey(double E, int shift) {
  long long b;
#if defined(__CUDA_ARCH__)
  b = __double_as_longlong(E);
#else
  memcpy(&b, &E, 8);
#endif
  return b >> shift;
}

// number of union points <= E (1 .. nUnion); 0 if E is outside [eMin, eMax].  Row (count - 1) of idxTab holds the
// nuclide indices; min(count, nUnion - 1) is the floor index on the union grid itself (binarySearch returns N-1 at the top edge)
__device__ __forceinline__ int unionSearch(const CeDev& c, double E) {
  if (!(E >= c.eMin) || !(E <= c.eMax)) return 0;
  long long b = hashKey(E, c.uShift) - c.keyMin;
  b = b < 0 ? 0 : (b >= c.nBuckets ? c.nBuckets - 1 : b);
  int u = __ldg(c.bucketStart + b);
  const int hi = __ldg(c.bucketStart + b + 1);
  while (u < hi && __ldg(c.uGrid + u) <= E) ++u;            // u = number of union points <= E
  return u < 1 ? 1 : u;
}

// one 32-byte sector per nuclide: 256-bit load (LDG.E.256) of { E_low, E_top, total_low, total_top }
__device__ __forceinline__ void ldPair(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
// nuclide%search without the union table: the bucket of e in the nuclide's own hashed index, then a short walk.
// lo = number of grid points whose key is below the key of e (all of them are < e); the points sharing e's bucket may lie on
// either side of it. Returns what binarySearch(eGrid, e) returns (genericProcedures.f90:132-166): the number of points <= e,
// at most N - 1.
__device__ __forceinline__ void nucBucket(const CeDev& c, int nuc0, double e, int& lo, int& hi) {
  long long b = (__double_as_longlong(e) >> __ldg(c.nbShift + nuc0)) - __ldg(c.nbKeyMin + nuc0);
  const int nB = __ldg(c.nbCount + nuc0);
  b = b < 0 ? 0 : (b >= nB ? nB - 1 : b);
  const int* q = c.nbStart + __ldg(c.nbOff + nuc0) + b;
  lo = __ldg(q); hi = __ldg(q + 1);
}
__device__ __forceinline__ int nucSearch(const CeDev& c, int nuc0, double e) {
  int p, hi; nucBucket(c, nuc0, e, p, hi);
  const double* g = c.grid + __ldg(c.gridOff + nuc0);
  while (p < hi && __ldg(g + p) <= e) ++p;
  const int N = __ldg(c.gridSize + nuc0);
  return p < 1 ? 1 : (p > N - 1 ? N - 1 : p);
}
// the index of nuclide nuc0 for an energy in union interval u: from the union table when the library has one
__device__ __forceinline__ int nucIndex(const CeDev& c, int u, double e, int nuc0) {
  if (c.idxTab) return __ldg(c.idxTab + (size_t)(u - 1) * c.nNuc + nuc0);
  return nucSearch(c, nuc0, e);
}
// the same search walking the 32-byte pair records of the total cross section (the record found is the one the interpolation
// needs: no separate read of the grid): record i covers [grid(i+1), grid(i+2)) in 1-based grid terms
__device__ __forceinline__ void nucPairSearch(const CeDev& c, int nuc0, double e, double& E_low, double& E_top, double& s_low, double& s_top) {
  int lo, hi; nucBucket(c, nuc0, e, lo, hi);
  const int last = __ldg(c.gridSize + nuc0) - 2;                 // last record
  int i = lo - 1; i = i < 0 ? 0 : (i > last ? last : i);
  const double* rec = c.pairTot + 4 * __ldg(c.pairOff + nuc0);
  ldPair(rec + 4 * (size_t)i, E_low, E_top, s_low, s_top);
  while (i < last && e >= E_top) { ++i; ldPair(rec + 4 * (size_t)i, E_low, E_top, s_low, s_top); }
}
// Sigma_t(material m, E) with E in union interval u: updateTotalMatXS (aceNeutronDatabase_class.f90:509-571)
__device__ __forceinline__ double matTotal(const CeDev& c, int u, double e, int m) {
  const int k0 = __ldg(c.matOff + m - 1), k1 = __ldg(c.matOff + m);
  double tot = 0.0;
  for (int k = k0; k < k1; ++k) {
    const int nuc = __ldg(c.matNuc + k) - 1;
    double E_low, E_top, s_low, s_top;
    if (c.idxTab) {
      const int idx = __ldg(c.idxTab + (size_t)(u - 1) * c.nNuc + nuc);    // what binarySearch(eGrid, E) returns for this nuclide
      ldPair(c.pairTot + 4 * (__ldg(c.pairOff + nuc) + (idx - 1)), E_low, E_top, s_low, s_top);
    } else nucPairSearch(c, nuc, e, E_low, E_top, s_low, s_top);
    const double f = (e - E_low) / (E_top - E_low);            // nuclide%search
    tot = tot + __ldg(c.matDens + k) * (s_top * f + (1.0 - f) * s_low);      // nuclide%totalXS
  }
  return tot * 1.0;
}
// Sigma_t for a library without the union table, as a kernel of its own (few registers: many lookups in flight per SM, which is
// what a chain of two dependent random memory accesses per nuclide needs): per nuclide one 32-byte record of the hashed index
// (L1 resident), the bucket entry, then the pair records from that point on. The terms are added in material order (matTotal).
// perm: lane j of the launch handles lookup perm[j] (the lookups binned by material and energy, k_ce_sort_*); nullptr: lookup j
__global__ void __launch_bounds__(256, 6) k_ce_total_hashed(const CeDev c, long long n, const double* __restrict__ E, const int* __restrict__ mat,
                                                            double* __restrict__ total, int* __restrict__ err, const int* __restrict__ perm) {
  for (long long tj = blockIdx.x * (long long)blockDim.x + threadIdx.x; tj < n; tj += (long long)gridDim.x * blockDim.x) {
    const long long t = perm ? perm[tj] : tj;
    const double e = E[t];
    if (!(e >= c.eMin) || !(e <= c.eMax)) { atomicMax(err, 1); continue; }             // "Failed to find energy"
    const int m = mat[t];
    if (m < 1 || m > c.nMat) { atomicMax(err, 2); continue; }
    const long long eb = __double_as_longlong(e);
    const int k0 = __ldg(c.matOff + m - 1), k1 = __ldg(c.matOff + m);
    double tot = 0.0;
#pragma unroll 2
    for (int k = k0; k < k1; ++k) {
      const int4* q = (const int4*)(c.nbMeta + (__ldg(c.matNuc + k) - 1));
      const int4 a = __ldg(q), b = __ldg(q + 1);                                        // { keyMin, nbOff | pairOff, shift, last }
      const long long keyMin = ((long long)(unsigned)a.x) | ((long long)a.y << 32), nbOff = ((long long)(unsigned)a.z) | ((long long)a.w << 32);
      const long long pairOff = ((long long)(unsigned)b.x) | ((long long)b.y << 32);
      int i = __ldg(c.nbStart + nbOff + ((eb >> b.z) - keyMin)) - 1;                    // inside the key range: eMin / eMax lie inside every grid
      i = i < 0 ? 0 : (i > b.w ? b.w : i);
      const double* rec = c.pairTot + 4 * pairOff;
      double E_low, E_top, s_low, s_top;
      ldPair(rec + 4 * (size_t)i, E_low, E_top, s_low, s_top);
      while (i < b.w && e >= E_top) { ++i; ldPair(rec + 4 * (size_t)i, E_low, E_top, s_low, s_top); }
      const double f = (e - E_low) / (E_top - E_low);                                   // nuclide%search
      tot = tot + __ldg(c.matDens + k) * (s_top * f + (1.0 - f) * s_low);               // nuclide%totalXS
    }
    total[t] = tot * 1.0;
  }
}
// The lookups of a batch binned by (material, energy): bin = (mat - 1) * nEb + (IEEE key of E with 10 mantissa bits) - so that the
// lanes of a warp ask for the same nuclides at neighbouring energies and their gathers fall into the same sectors.  A counting
// sort: histogram, exclusive scan (the engine's scan kernels), scatter of the lookup numbers (order inside a bin is arbitrary; every
// lookup is computed exactly as in an unsorted launch, only the lane that computes it changes).
constexpr int SORT_MBITS = 10;
__device__ __forceinline__ int sortBin(const CeDev& c, double e, int m, int nEb, long long keyLo) {
  long long b = (__double_as_longlong(e) >> (52 - SORT_MBITS)) - keyLo;
  b = b < 0 ? 0 : (b >= nEb ? nEb - 1 : b);
  const int mm = (m < 1 || m > c.nMat) ? 1 : m;
  return (mm - 1) * nEb + (int)b;
}
__global__ void k_ce_sort_hist(const CeDev c, long long n, const double* __restrict__ E, const int* __restrict__ mat, int nEb, long long keyLo, int* __restrict__ hist) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    atomicAdd(hist + sortBin(c, E[i], mat ? mat[i] : 1, nEb, keyLo), 1);
}
__global__ void k_ce_sort_scatter(const CeDev c, long long n, const double* __restrict__ E, const int* __restrict__ mat, int nEb, long long keyLo, int* __restrict__ cursor, int* __restrict__ perm) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    perm[atomicAdd(cursor + sortBin(c, E[i], mat ? mat[i] : 1, nEb, keyLo), 1)] = (int)i;
}
// initMajorant (:1545-1617): majorant(i) = max over materials of Sigma_t(E_i), nudged up by 1e-6
__global__ void k_ce_majorant(const CeDev c, double* uMaj) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < c.nUnion; j += gridDim.x * blockDim.x) {
    const double e = c.uGrid[j];
    const int u = j + 1;
    double maj = 0.0;
    for (int a = 0; a < c.nActive; ++a) maj = fmax(maj, matTotal(c, u, e, __ldg(c.activeMat + a)));
    uMaj[j] = maj * (1.0 + 1.0e-06);
  }
}

// total / macro set (8 values) / majorant / per-nuclide index of nuclide `probeNuc`, whichever output pointers are given
__global__ void __launch_bounds__(256) k_ce_lookup(const CeDev c, long long n, const double* __restrict__ E, const int* __restrict__ mat,
                                                   double* __restrict__ total, double* __restrict__ macro, double* __restrict__ maj,
                                                   int* __restrict__ probeIdx, int probeNuc, int* __restrict__ err, const int* __restrict__ perm) {
  for (long long ij = blockIdx.x * (long long)blockDim.x + threadIdx.x; ij < n; ij += (long long)gridDim.x * blockDim.x) {
    const long long i = perm ? perm[ij] : ij;
    const double e = E[i];
    // the union interval is needed by the majorant and by the [union interval][nuclide] table; without either only the bounds
    const int u = (maj || c.idxTab) ? unionSearch(c, e) : ((e >= c.eMin && e <= c.eMax) ? 1 : 0);
    if (u == 0) { atomicMax(err, 1); continue; }               // "Failed to find energy"
    if (maj) {                                                 // updateMajorantXS
      const int uu = u > c.nUnion - 1 ? c.nUnion - 1 : u;
      const double E_low = __ldg(c.uGrid + uu - 1), E_top = __ldg(c.uGrid + uu);
      const double f = (e - E_low) / (E_top - E_low);
      maj[i] = __ldg(c.uMaj + uu) * f + (1.0 - f) * __ldg(c.uMaj + uu - 1);
    }
    if (probeIdx) probeIdx[i] = nucIndex(c, u, e, probeNuc - 1);
    if (!total && !macro) continue;
    const int m = mat[i];
    if (m < 1 || m > c.nMat) { atomicMax(err, 2); continue; }
    if (!macro) { total[i] = matTotal(c, u, e, m); continue; }
    const int k0 = __ldg(c.matOff + m - 1), k1 = __ldg(c.matOff + m);
    double tot = 0.0;
    double xs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = k0; k < k1; ++k) {
      const int nuc = __ldg(c.matNuc + k) - 1;
      const double dens = __ldg(c.matDens + k);
      const int idx = nucIndex(c, u, e, nuc);                  // what binarySearch(eGrid, E) returns for this nuclide
      const double* g = c.grid + __ldg(c.gridOff + nuc) + (idx - 1);
      const double E_low = __ldg(g), E_top = __ldg(g + 1);
      const double f = (e - E_low) / (E_top - E_low);          // nuclide%search
      const int rows = __ldg(c.rows + nuc);
      const double* d = c.data + __ldg(c.dataOff + nuc) + (size_t)(idx - 1) * rows;
      if (macro) {
        if (rows == 8) {
          const double2* p = (const double2*)d;                // 128 B: two energy points x 8 rows
          double2 a0 = __ldg(p), a1 = __ldg(p + 1), a2 = __ldg(p + 2), a3 = __ldg(p + 3), b0 = __ldg(p + 4), b1 = __ldg(p + 5), b2 = __ldg(p + 6), b3 = __ldg(p + 7);
          const double lo[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y}, hi8[8] = {b0.x, b0.y, b1.x, b1.y, b2.x, b2.y, b3.x, b3.y};
#pragma unroll
          for (int r = 0; r < 8; ++r) xs[r] = xs[r] + dens * (hi8[r] * f + (1.0 - f) * lo[r]);
        } else {
          const double2* p = (const double2*)d;                // 64 B: two energy points x 4 rows
          double2 a0 = __ldg(p), a1 = __ldg(p + 1), b0 = __ldg(p + 2), b1 = __ldg(p + 3);
          const double lo[4] = {a0.x, a0.y, a1.x, a1.y}, hi4[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
          for (int r = 0; r < 4; ++r) xs[r] = xs[r] + dens * (hi4[r] * f + (1.0 - f) * lo[r]);
#pragma unroll
          for (int r = 4; r < 8; ++r) xs[r] = xs[r] + dens * 0.0;
        }
      }
      if (total) tot = tot + dens * (__ldg(d + rows) * f + (1.0 - f) * __ldg(d));      // nuclide%totalXS
    }
    if (total) total[i] = tot * 1.0;
    if (macro) {
#pragma unroll
      for (int r = 0; r < 8; ++r) macro[8 * i + r] = xs[r];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side: build the union grid, the index table, the hash buckets; the majorant is filled by the kernel itself
// ---------------------------------------------------------------------------------------------------------
struct CeHost {
  bool loaded = false;
  CeDev dev{};
  std::vector<void*> allocs;
  std::vector<double> uGrid, uMaj;
  int nNuc = 0, nMat = 0;
  long long rawBytes = 0, indexBytes = 0;           // the nuclide tables as the reference holds them / everything this file adds to look them up
  long long algBytesTotal(int nNucInMat) const { return 36LL * nNucInMat + 20; }
};

template <typename T>
static T* ceUpload(CeHost& H, const std::vector<T>& v, std::string& err) {
  T* p = nullptr;
  size_t bytes = std::max<size_t>(sizeof(T) * v.size(), 16);
  if (cudaMalloc(&p, bytes) != cudaSuccess) { err = "cudaMalloc failed (CE tables)"; return nullptr; }
  if (!v.empty() && cudaMemcpy(p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice) != cudaSuccess) { err = "cudaMemcpy failed (CE tables)"; return nullptr; }
  H.allocs.push_back(p);
  return p;
}
static void ceFree(CeHost& H) { for (void* p : H.allocs) cudaFree(p); H.allocs.clear(); H.loaded = false; }

static int ceBuild(CeHost& H, const sb_ce_flat* f, std::string& err, const std::vector<int>* activeIn = nullptr) {
  ceFree(H);
  if (f->n_nuc < 1 || f->n_mat < 1) { err = "sb_load_ce_data: invalid sizes"; return -1; }
  const int nNuc = f->n_nuc, nMat = f->n_mat;
  std::vector<long long> gridOff(nNuc), dataOff(nNuc); std::vector<int> rows(nNuc), gsize(nNuc);
  long long go = 0, dofs = 0;
  for (int n = 0; n < nNuc; ++n) {
    if (f->grid_size[n] < 2 || (f->rows[n] != 4 && f->rows[n] != 8)) { err = "sb_load_ce_data: nuclide grid must have >= 2 points and 4 or 8 rows"; return -1; }
    gridOff[n] = go; dataOff[n] = dofs; rows[n] = f->rows[n]; gsize[n] = f->grid_size[n];
    go += f->grid_size[n]; dofs += (long long)f->grid_size[n] * f->rows[n];
    dofs = (dofs + 3) & ~3LL;                                         // keep every nuclide's data 32-byte aligned for the vector loads
  }
  std::vector<double> grid(go), data(dofs, 0.0);
  { long long gs = 0, ds = 0;
    for (int n = 0; n < nNuc; ++n) {
      std::copy(f->grid + gs, f->grid + gs + gsize[n], grid.begin() + gridOff[n]);
      std::copy(f->data + ds, f->data + ds + (long long)gsize[n] * rows[n], data.begin() + dataOff[n]);
      for (int i = 1; i < gsize[n]; ++i) if (grid[gridOff[n] + i] < grid[gridOff[n] + i - 1]) { err = "sb_load_ce_data: energy grid is not sorted"; return -1; }
      gs += gsize[n]; ds += (long long)gsize[n] * rows[n];
    } }
  std::vector<int> matOff(f->mat_off, f->mat_off + nMat + 1), matNuc(f->mat_nuc, f->mat_nuc + f->mat_off[nMat]);
  std::vector<double> matDens(f->mat_dens, f->mat_dens + f->mat_off[nMat]);
  for (int v : matNuc) if (v < 1 || v > nNuc) { err = "sb_load_ce_data: material refers to an unknown nuclide"; return -1; }
  // energy bounds of the database (aceNeutronDatabase_class.f90:1044-1051) and the unionised grid (initMajorant)
  double eMin = grid[gridOff[0]], eMax = grid[gridOff[0] + gsize[0] - 1];
  for (int n = 0; n < nNuc; ++n) { eMin = std::max(eMin, grid[gridOff[n]]); eMax = std::min(eMax, grid[gridOff[n] + gsize[n] - 1]); }
  std::vector<int> active;
  if (activeIn) active = *activeIn; else for (int m = 1; m <= nMat; ++m) active.push_back(m);
  for (int m : active) if (m < 1 || m > nMat) { err = "sb_load_ce: active material index out of range"; return -1; }
  std::vector<char> used(nNuc, 0); for (int m : active) for (int k = matOff[m - 1]; k < matOff[m]; ++k) used[matNuc[k] - 1] = 1;
  std::vector<double> u;
  for (int n = 0; n < nNuc; ++n) if (used[n]) for (int i = 0; i < gsize[n]; ++i) { double e = grid[gridOff[n] + i]; if (!(e < eMin || e > eMax)) u.push_back(e); }
  std::sort(u.begin(), u.end()); u.erase(std::unique(u.begin(), u.end()), u.end());
  const int nU = (int)u.size();
  if (nU < 2) { err = "sb_load_ce_data: unionised grid has fewer than 2 points"; return -1; }
  // idxTab[j][n] = min(N_n - 1, #{ i : grid_n(i) <= U_j }) : the value of binarySearch for any E in [U_j, U_{j+1})
  // (last row: E == U_last exactly)
  // - only while it stays small: it grows with (union points) x (nuclides); a real library (hundreds of nuclides, 1e6 union
  // points) is served by the per-nuclide hashed index below, whose size is proportional to the data itself
  size_t idxTabMax = (size_t)256 << 20;
  if (const char* e = getenv("SB_CE_IDXTAB_MAX_MB")) idxTabMax = (size_t)atoll(e) << 20;
  const bool haveTab = (size_t)nU * nNuc * sizeof(int) <= idxTabMax;
  std::vector<int> idxTab(haveTab ? (size_t)nU * nNuc : 0, 1);
  if (haveTab)
    for (int n = 0; n < nNuc; ++n) {
      const double* g = &grid[gridOff[n]]; int N = gsize[n], p = 0;
      for (int j = 0; j < nU; ++j) {
        while (p < N && g[p] <= u[j]) ++p;
        idxTab[(size_t)j * nNuc + n] = std::max(1, std::min(N - 1, p));
      }
    }
  // per-nuclide hashed index: the smallest number of mantissa bits that gives at least one bucket per grid point
  std::vector<int> nbShift(nNuc), nbCount(nNuc), nbStart; std::vector<long long> nbOff(nNuc), nbKeyMin(nNuc);
  auto bitsOf = [](double x) { long long b; memcpy(&b, &x, 8); return b; };
  for (int n = 0; n < nNuc; ++n) {
    const double* g = &grid[gridOff[n]]; const int N = gsize[n];
    int shift = 52;
    for (; shift > 52 - 20; --shift) if ((bitsOf(g[N - 1]) >> shift) - (bitsOf(g[0] > 0.0 ? g[0] : 1.0e-300) >> shift) + 1 >= (long long)N) break;
    const long long k0 = bitsOf(g[0] > 0.0 ? g[0] : 1.0e-300) >> shift, k1 = bitsOf(g[N - 1]) >> shift;
    const int nBk = (int)(k1 - k0 + 1);
    nbShift[n] = shift; nbKeyMin[n] = k0; nbCount[n] = nBk; nbOff[n] = (long long)nbStart.size();
    int p = 0;
    for (int b = 0; b <= nBk; ++b) { while (p < N && (bitsOf(g[p]) >> shift) - k0 < b) ++p; nbStart.push_back(p); }
  }
  // pair table of the total cross section: one 32-byte record per nuclide grid interval
  std::vector<long long> pairOff(nNuc); long long po = 0;
  for (int n = 0; n < nNuc; ++n) { pairOff[n] = po; po += gsize[n] - 1; }
  std::vector<double> pairTot((size_t)po * 4);
  for (int n = 0; n < nNuc; ++n)
    for (int i = 0; i + 1 < gsize[n]; ++i) {
      double* q = &pairTot[(size_t)(pairOff[n] + i) * 4];
      q[0] = grid[gridOff[n] + i]; q[1] = grid[gridOff[n] + i + 1];
      q[2] = data[dataOff[n] + (size_t)i * rows[n]]; q[3] = data[dataOff[n] + (size_t)(i + 1) * rows[n]];
    }
  // hash buckets on the IEEE bits
  int uShift = 52 - HASH_MBITS_MIN;
  while (uShift > 52 - HASH_MBITS_MAX && hashKey(u.back(), uShift) - hashKey(u.front(), uShift) + 1 < (long long)nU) --uShift;
  const long long keyMin = hashKey(u.front(), uShift), keyMax = hashKey(u.back(), uShift);
  const int nB = (int)(keyMax - keyMin + 1);
  std::vector<int> bucketStart(nB + 1, 0);
  { int p = 0; for (int b = 0; b <= nB; ++b) { while (p < nU && hashKey(u[p], uShift) - keyMin < b) ++p; bucketStart[b] = p; } }
  CeDev& d = H.dev;
  d.nNuc = nNuc; d.nMat = nMat; d.nUnion = nU; d.nBuckets = nB; d.uShift = uShift; d.keyMin = keyMin; d.eMin = u.front(); d.eMax = u.back();
  H.uGrid = u; H.uMaj.assign(nU, 0.0);
  d.grid = ceUpload(H, grid, err); d.data = ceUpload(H, data, err); d.gridOff = ceUpload(H, gridOff, err); d.dataOff = ceUpload(H, dataOff, err);
  d.rows = ceUpload(H, rows, err); d.gridSize = ceUpload(H, gsize, err);
  d.matOff = ceUpload(H, matOff, err); d.matNuc = ceUpload(H, matNuc, err); d.matDens = ceUpload(H, matDens, err);
  d.pairTot = ceUpload(H, pairTot, err); d.pairOff = ceUpload(H, pairOff, err);
  d.activeMat = ceUpload(H, active, err); d.nActive = (int)active.size();
  d.uGrid = ceUpload(H, u, err); d.uMaj = ceUpload(H, H.uMaj, err); d.idxTab = haveTab ? ceUpload(H, idxTab, err) : nullptr; d.bucketStart = ceUpload(H, bucketStart, err);
  std::vector<NucMeta> meta(nNuc);
  for (int n = 0; n < nNuc; ++n) { meta[n].keyMin = nbKeyMin[n]; meta[n].nbOff = nbOff[n]; meta[n].pairOff = pairOff[n]; meta[n].shift = nbShift[n]; meta[n].last = gsize[n] - 2; }
  d.nbMeta = ceUpload(H, meta, err);
  d.nbStart = ceUpload(H, nbStart, err); d.nbOff = ceUpload(H, nbOff, err); d.nbShift = ceUpload(H, nbShift, err); d.nbKeyMin = ceUpload(H, nbKeyMin, err); d.nbCount = ceUpload(H, nbCount, err);
  H.rawBytes = (long long)(grid.size() + data.size()) * 8;
  H.indexBytes = (long long)(pairTot.size() * 8 + nbStart.size() * 4 + idxTab.size() * 4 + bucketStart.size() * 4 + (u.size() * 2) * 8);
  if (!err.empty()) return -1;
  H.nNuc = nNuc; H.nMat = nMat; H.loaded = true;
  return 0;
}

}  // namespace sbce
