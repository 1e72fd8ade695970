// scone_b200 engine: CUDA kernels for sm_100a + the C ABI of include/scone_b200.h.
//
// Design (DESIGN.md has the long form):
//  * k_histories  - persistent event loop. One lane owns one history at a time and keeps its state in
//                   registers; every loop iteration is one event round (tentative-flight, then collision
//                   for the lanes whose flight ended in a real collision); a lane whose history died is
//                   refilled from a warp-private chunk of the bank (warp-level compaction by ballot), chunks
//                   are claimed with one global atomic. Tables (geometry graph + MG data) are staged once
//                   per CTA into shared memory with a bulk async copy (TMA, cp.async.bulk) when they fit.
//  * fission bank - sites are appended with one atomic per warp per round (warp scan of the counts), keyed
//                   (broodID, seq); k_scan* + k_sort_sites place them in the reference's stable brood order.
//  * tallies      - per-history k-eff scores are written per history and reduced in a fixed tree
//                   (bitwise reproducible); map tallies use f64 L2 atomics (red.global.add.f64).
//  * resampling   - normSize_Repr: per-site LCG numbers by skip-ahead, exact k-th smallest by a two-level
//                   radix select, keep/duplicate flags, scan, scatter.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sb_device.cuh"
#include "sb_hist.cuh"
#include "sb_track.cuh"
#include "sb_ce.cuh"
#include "sb_cehist.cuh"
#include "sb_ceevent.cuh"

using namespace sbd;
using sbh::Bank; using sbh::CycleDev; using sbh::HistArgs; using sbh::HotLayout; using sbh::HUni;

// Programmatic dependent launch for the chain of small kernels that closes a cycle (brood ordering, reductions, cycle close,
// normSize_Repr, the peer exchange): each of them starts with PDL_ENTER - let the next kernel of the stream be scheduled now, then
// wait until everything before this kernel has completed and is visible - so the launch latency and the prologue of kernel i + 1
// overlap the execution of kernel i; the data dependencies are the same as with plain stream order.
#define PDL_ENTER() do { asm volatile("griddepcontrol.launch_dependents;"); asm volatile("griddepcontrol.wait;" ::: "memory"); } while (0)
template <typename... P, typename... A>
static inline void pdlLaunch(void (*k)(P...), int grid, int block, cudaStream_t st, A&&... a) {
  cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute at{}; at.id = cudaLaunchAttributeProgrammaticStreamSerialization; at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, k, P(a)...);
}

template <typename... P, typename... A>
static inline void pdlLaunchSmem(void (*k)(P...), int grid, int block, size_t smem, cudaStream_t st, A&&... a) {
  cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at{}; at.id = cudaLaunchAttributeProgrammaticStreamSerialization; at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, k, P(a)...);
}

#define CUDA_OK(call)                                                                         \
  do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return -1; } } while (0)

// ------------------------------------------------------------------------------------------------
// exclusive scan of int32 (3 kernels) -- used for brood offsets and resampling compaction
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;                  // per thread
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

__device__ __forceinline__ int blockExclusiveScan(int v, int* sWarp, int& blockTotal) {
  const unsigned FULL = 0xffffffffu;
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += t; }
  if (lane == 31) sWarp[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int ws = (lane < (int)(blockDim.x >> 5)) ? sWarp[lane] : 0;
    int winc = ws;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(FULL, winc, d); if (lane >= d) winc += t; }
    sWarp[lane] = winc - ws;                   // exclusive warp offsets
    if (lane == 31) sWarp[32] = winc;
  }
  __syncthreads();
  int res = sWarp[wid] + inc - v;
  blockTotal = sWarp[32];
  __syncthreads();
  return res;
}

// nPtr: device pointer to the element count (so no host sync is needed); nMax bounds the grid
__global__ void k_scan_reduce(const int* in, const int* nPtr, int* tileSums) {
  PDL_ENTER();
  __shared__ int sWarp[33];
  int n = *nPtr;
  int tile = blockIdx.x;
  if ((long long)tile * SCAN_TILE >= n) { return; }
  int sum = 0;
  int b = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) if (b + i < n) sum += in[b + i];
  int tot; blockExclusiveScan(sum, sWarp, tot);
  if (threadIdx.x == 0) tileSums[tile] = tot;
}
__global__ void __launch_bounds__(1024) k_scan_tiles(int* tileSums, const int* nPtr, int* totalOut) {
  PDL_ENTER();
  __shared__ int sWarp[33];
  __shared__ int carry;
  int n = *nPtr;
  int nTiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b = 0; b < nTiles; b += blockDim.x) {
    int i = b + threadIdx.x;
    int v = (i < nTiles) ? tileSums[i] : 0;
    int tot; int ex = blockExclusiveScan(v, sWarp, tot);
    int c = carry;
    if (i < nTiles) tileSums[i] = c + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0 && totalOut) *totalOut = carry;
}
__global__ void k_scan_apply(const int* in, const int* nPtr, const int* tileSums, int* out) {
  PDL_ENTER();
  __shared__ int sWarp[33];
  int n = *nPtr;
  int tile = blockIdx.x;
  if ((long long)tile * SCAN_TILE >= n) return;
  int b = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS]; int sum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = (b + i < n) ? in[b + i] : 0; sum += v[i]; }
  int tot; int ex = blockExclusiveScan(sum, sWarp, tot) + tileSums[tile];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) { if (b + i < n) out[b + i] = ex; ex += v[i]; }
}

// ------------------------------------------------------------------------------------------------
// fission bank ordering: particleDungeon%sortByBroodID (particleDungeon_class.f90:923-981) is a
// stable counting sort by broodID; with the per-history site counts known, the destination of site
// (brood, seq) is offset[brood] + seq.
// ------------------------------------------------------------------------------------------------
__global__ void k_sort_sites(Bank src, Bank dst, const int* offsets, const CycleDev* cd, int cap) {
  PDL_ENTER();
  int n = min(cd->nSites, cap);
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    int b = src.brood[s];
    int d = offsets[b] + src.seq[s];
    dst.rx[d] = src.rx[s]; dst.ry[d] = src.ry[s]; dst.rz[d] = src.rz[s];
    dst.ux[d] = src.ux[s]; dst.uy[d] = src.uy[s]; dst.uz[d] = src.uz[s];
    dst.w[d] = src.w[s]; dst.G[d] = src.G[s]; dst.E[d] = src.E[s]; dst.brood[d] = b; dst.seq[d] = src.seq[s];
  }
}

// ------------------------------------------------------------------------------------------------
// deterministic reduction of the per-history scores (fixed tiling, fixed tree): the result depends on
// n only, never on the execution order of the history kernel
// ------------------------------------------------------------------------------------------------
constexpr int RED_BLOCKS = 592;      // 4 per SM on 148 SMs; the tiling is a constant of the algorithm
constexpr int RED_THREADS = 256;
struct RedOut { double prod, abs, leak, scat, wgt, endw; };

__device__ __forceinline__ double warpSum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  return v;
}
// blocks [0, RED_BLOCKS): per-history scores + start weights of this cycle's bank;
// the same blocks also sum the weights of the (sorted) next-cycle bank: popWeight for keffAnalogClerk%reportCycleEnd
__global__ void __launch_bounds__(RED_THREADS) k_reduce_hist(int n, const double* hProd, const double* hAbs, const double* hLeak, const double* hScat,
                                                             const double* wIn, const double* wSites, const CycleDev* cd, int cap, RedOut* partial) {
  PDL_ENTER();
  __shared__ double sd[6][RED_THREADS / 32];
  // contiguous slice per block, strided by thread inside the slice: fixed order for fixed n
  double v[6] = {0, 0, 0, 0, 0, 0};
  {
    int per = (n + RED_BLOCKS - 1) / RED_BLOCKS;
    int b0 = blockIdx.x * per, b1 = min(n, b0 + per);
    for (int i = b0 + threadIdx.x; i < b1; i += RED_THREADS) { v[0] += hProd[i]; v[1] += hAbs[i]; v[2] += hLeak[i]; v[3] += hScat[i]; v[4] += wIn[i]; }
  }
  {
    int m = min(cd->nSites, cap);
    int per = (m + RED_BLOCKS - 1) / RED_BLOCKS;
    int b0 = blockIdx.x * per, b1 = min(m, b0 + per);
    for (int i = b0 + threadIdx.x; i < b1; i += RED_THREADS) v[5] += wSites[i];
  }
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 6; ++k) { double s = warpSum(v[k]); if (lane == 0) sd[k][wid] = s; }
  __syncthreads();
  if (threadIdx.x < 6) {
    double s = 0.0;
    for (int i = 0; i < RED_THREADS / 32; ++i) s += sd[threadIdx.x][i];
    ((double*)(partial + blockIdx.x))[threadIdx.x] = s;
  }
}

// one warp: lane l sums partials l, l+32, ... in order, then a fixed shuffle tree -> ksum[6] =
// { implicit production, implicit absorption, analog leakage, scatter production, start weight, end weight }
// (the bins of keffImplicitClerk / keffAnalogClerk that are reduced across ranks every cycle: mpiSync = 1,
//  eigenPhysicsPackage_class.f90:605-640, scoreMemory_class.f90:404-431)
__global__ void k_sum_partials(const RedOut* partial, double* ksum, const CycleDev* cd, int cap) {
  PDL_ENTER();
  const int lane = threadIdx.x;
  double v[6] = {0, 0, 0, 0, 0, 0};
  for (int i = lane; i < RED_BLOCKS; i += 32) {
    const double* p = (const double*)(partial + i);
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] += p[k];
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) v[k] = warpSum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 6; ++k) ksum[k] = v[k];
    ksum[6] = (double)min(cd->nSites, cap); ksum[7] = 0.0;      // the bank size rides along with the sums (one all-gather per cycle)
  }
}
// k-eff clerks that a tally block names itself (next to the attachment clerks the engine always runs): kind and 1-based address
struct UserKeff { int n; int kind[4]; int addr[4]; };
// tallyAdmin%reportCycleEnd (tallyAdmin_class.f90:735-794) for the attachment clerks + normalisation factor
//   keffAnalogClerk%closeCycle (keffAnalogClerk_class.f90:156-176), keffImplicitClerk%closeCycle (:292-312)
// ksum: the score sums the attachment clerks close with (reduced over the ranks in a ranked run: that tally is mpiSync in the
// reference, eigenPhysicsPackage_class.f90:605-640); ksumLocal: this rank's own sums, which is what the k-eff clerks of the
// deck's own tally blocks hold (those are not synchronised per cycle; collectDistributed sums them once at the end)
__global__ void k_close_cycle_head(const double* ksum, const double* ksumLocal, CycleDev* cd, int phase, double kNorm,
                                   double* bins, int normAddr, double normVal, UserKeff uk, double* csum, double* csum2) {
  PDL_ENTER();
  if (threadIdx.x != 0) return;
  const double prod = ksum[0], abs_ = ksum[1], leak = ksum[2], scat = ksum[3], wgt = ksum[4], endW = ksum[5];
  for (int i = 0; i < uk.n; ++i) {                          // user clerks: scores into BIN (closed with the other bins), k accumulated directly
    const int a0 = uk.addr[i] - 1;
    const double lprod = ksumLocal[0], labs = ksumLocal[1], lleak = ksumLocal[2], lscat = ksumLocal[3], lwgt = ksumLocal[4], lend = ksumLocal[5];
    if (uk.kind[i] == SB_CLERK_KEFF_ANALOG) {               // keffAnalogClerk_class.f90:132-176
      bins[a0] = lwgt; bins[a0 + 1] = lend;
      const double k = lend / lwgt * kNorm;
      csum[a0 + 2] = csum[a0 + 2] + k; csum2[a0 + 2] = csum2[a0 + 2] + k * k;
    } else {                                                // keffImplicitClerk_class.f90:180-312
      bins[a0] = lprod; bins[a0 + 1] = labs; bins[a0 + 2] = lscat; bins[a0 + 3] = lleak;
      const double k = lprod / (labs + lleak - lscat);
      csum[a0 + 4] = csum[a0 + 4] + k; csum2[a0 + 4] = csum2[a0 + 4] + k * k;
    }
  }
  cd->startWgt = wgt; cd->endWgt = endW;
  cd->impProd = prod; cd->impAbs = abs_; cd->anaLeak = leak; cd->scatProd = scat;
  cd->kAnalog = endW / wgt * kNorm;
  cd->kImplicit = prod / (abs_ + leak - scat);
  double k = (phase == 0) ? cd->kAnalog : cd->kImplicit;
  cd->kCsum[phase] = cd->kCsum[phase] + k;
  cd->kCsum2[phase] = cd->kCsum2[phase] + k * k;
  cd->kBatches[phase] += 1;
  int N = cd->kBatches[phase];                              // scoreMemory%getResult (scoreMemory_class.f90:537-573)
  double mean = cd->kCsum[phase] / N;
  double inv_N = 1.0 / N, inv_Nm1 = (N != 1) ? 1.0 / (N - 1) : 1.0;
  double sd = cd->kCsum2[phase] * inv_N * inv_Nm1 - mean * mean * inv_Nm1;
  cd->kCum = mean; cd->kCumStd = sqrt(sd);
  double nf = 1.0;
  if (normAddr > 0) {
    double sc = bins[normAddr - 1];
    if (sc == 0.0) { atomicMax(&cd->error, SB_ERR_NORM); sc = 1.0; }
    nf = normVal / sc;
  }
  cd->normFactor = nf;
}
// scoreMemory%closeCycle (scoreMemory_class.f90:309-342)
__global__ void k_close_cycle_bins(double* bins, double* lastBins, double* csum, double* csum2, int nBins, const CycleDev* cd) {
  PDL_ENTER();
  double nf = cd->normFactor;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nBins; i += gridDim.x * blockDim.x) {
    double b = bins[i];
    double res = b * nf;
    lastBins[i] = b;
    bins[i] = 0.0;
    csum[i] = csum[i] + res;
    csum2[i] = csum2[i] + res * res;
  }
}

// ------------------------------------------------------------------------------------------------
// normSize_Repr (particleDungeon_class.f90:431-602) on the device.
// The random numbers that decide the fate of the sites are one sequential LCG stream over the
// concatenation of all ranks' banks (rank r starts where rank r-1 ended, :497-511); every rank generates
// the whole stream itself (a pure function of the master RNG state and the sizes) and finds the same
// threshold, then applies it to its own slice [offLocal, offLocal + nLocal).
// ------------------------------------------------------------------------------------------------
struct NormDev { int nGlobal, offLocal, nLocal, totPop, check, pad; };
__global__ void k_norm_setup(NormDev* nd, const CycleDev* cd, int cap, int totPop, int nGlobal, int offLocal, int check) {
  PDL_ENTER();
  int nLocal = min(cd->nSites, cap);
  nd->nLocal = nLocal; nd->totPop = totPop; nd->check = check;
  nd->nGlobal = (nGlobal < 0) ? nLocal : nGlobal;
  nd->offLocal = (nGlobal < 0) ? 0 : offLocal;
}
constexpr int RN_CHUNK = 16;
// rn_j = j-th number of the LCG stream started at `state0` (j = 1..n), stored as the integer state
__global__ void k_rn_generate(unsigned long long* rn, const NormDev* nd, uint64_t state0) {
  PDL_ENTER();
  int n = nd->nGlobal;
  int nChunks = (n + RN_CHUNK - 1) / RN_CHUNK;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nChunks; c += gridDim.x * blockDim.x) {
    int j0 = c * RN_CHUNK;
    uint64_t s = rng_skip(state0, (int64_t)j0);
    for (int j = j0; j < min(n, j0 + RN_CHUNK); ++j) { s = (RNG_G * s) & RNG_MASK; s = (s + 1ULL) & RNG_MASK; rn[j] = s; }
  }
}
constexpr int SEL_BITS = 16;
constexpr int SEL_BINS = 1 << SEL_BITS;
constexpr int SEL_CAND_CAP = 1 << 18;
__global__ void k_sel_hist(const unsigned long long* rn, const NormDev* nd, int* hist) {
  PDL_ENTER();
  int n = nd->nGlobal;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    atomicAdd(&hist[(int)(rn[j] >> (63 - SEL_BITS))], 1);
}
// heapSize-th smallest (1-based rank k): find the 16-bit bin holding it.
// 1024 threads: each sums its 64 consecutive bins, one block scan, the owner of the rank walks its bins
__global__ void __launch_bounds__(1024) k_sel_find_bin(const int* hist, CycleDev* cd, const NormDev* nd) {
  PDL_ENTER();
  __shared__ int sWarp[33];
  int totSites = nd->nGlobal;
  int excess = totSites - nd->totPop;
  int k = (totSites <= 0) ? 0 : ((excess < 0) ? (int)(((long long)(-excess)) % totSites) : excess);     // heapSize
  if (threadIdx.x == 0) { cd->selBin = -1; cd->selRank = 0; cd->nCand = 0; }
  __syncthreads();
  if (k == 0) return;
  constexpr int PER = SEL_BINS / 1024;
  const int4* my = (const int4*)(hist + threadIdx.x * PER);
  int sum = 0;
#pragma unroll 4
  for (int i = 0; i < PER / 4; ++i) { int4 v = my[i]; sum += (v.x + v.y) + (v.z + v.w); }
  int tot; int before = blockExclusiveScan(sum, sWarp, tot);
  if (before < k && before + sum >= k) {
    const int* mine = hist + threadIdx.x * PER;
    for (int i = 0; i < PER; ++i) {
      int v = mine[i];
      if (before < k && before + v >= k) { cd->selBin = threadIdx.x * PER + i; cd->selRank = k - before; break; }
      before += v;
    }
  }
}
__global__ void k_sel_collect(const unsigned long long* rn, CycleDev* cd, const NormDev* nd, unsigned long long* cand) {
  PDL_ENTER();
  int n = nd->nGlobal;
  int bin = cd->selBin;
  if (bin < 0) return;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    unsigned long long s = rn[j];
    if ((int)(s >> (63 - SEL_BITS)) == bin) {
      int p = atomicAdd(&cd->nCand, 1);
      if (p < SEL_CAND_CAP) cand[p] = s;
    }
  }
}
// threshold = selRank-th smallest of the candidates (states are distinct within the LCG period)
__global__ void __launch_bounds__(1024) k_sel_threshold(const unsigned long long* cand, CycleDev* cd) {
  PDL_ENTER();
  int m = cd->nCand;
  if (cd->selBin < 0) { if (threadIdx.x == 0) { cd->thrState = 0; cd->thrReal = 1.0; } return; }
  if (m > SEL_CAND_CAP) { if (threadIdx.x == 0) atomicMax(&cd->error, SB_ERR_NORM); return; }
  int want = cd->selRank - 1;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    unsigned long long v = cand[i];
    int less = 0;
    for (int j = 0; j < m; ++j) less += (cand[j] < v) ? 1 : 0;
    if (less == want) { cd->thrState = v; cd->thrReal = (double)(long long)v * (1.0 / 9223372036854775808.0); }
  }
}
// keep (excess > 0: rn > threshold) or duplicate (excess < 0: rn <= threshold) flags of the local slice
__global__ void k_norm_flags(const unsigned long long* rn, const CycleDev* cd, const NormDev* nd, int* flag) {
  PDL_ENTER();
  int n = nd->nLocal;
  int excess = nd->nGlobal - nd->totPop;
  int nDup = (excess < 0) ? (int)(((long long)(-excess)) % nd->nGlobal) : 0;
  double thr = cd->thrReal;
  const unsigned long long* mine = rn + nd->offLocal;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    double x = (double)(long long)mine[j] * (1.0 / 9223372036854775808.0);
    int f;
    if (excess > 0) f = (x > thr) ? 1 : 0;
    else if (excess < 0) f = (nDup != 0 && x <= thr) ? 1 : 0;
    else f = 1;
    flag[j] = f;
  }
}
// scatter into the new bank. src is brood-sorted; offs = exclusive scan of flags; hOff/hCnt = per-history
// offsets/counts of src (brood segments). For excess < 0 the result is already in the order that the second
// sortByBroodID of the reference produces (:571-574): per brood, copies first, duplicates after.
__global__ void k_norm_scatter(Bank src, Bank dst, const int* flag, const int* offs, const int* hOff, const int* hCnt,
                               const NormDev* nd, int dstCap, CycleDev* cdw) {
  PDL_ENTER();
  int n = nd->nLocal;
  int excess = nd->nGlobal - nd->totPop;
  int nCopies = (excess < 0) ? (-excess) / nd->nGlobal : 0;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    int d0 = -1, stride = 0, reps = 0, dDup = -1;
    if (excess > 0) { if (flag[j]) { d0 = offs[j]; reps = 1; } }
    else if (excess == 0) { d0 = j; reps = 1; }
    else {
      int b = src.brood[j];
      int sb = hOff[b], nb = hCnt[b];
      int Dsb = offs[sb];
      int start = (nCopies + 1) * sb + Dsb;
      d0 = start + (j - sb); stride = nb; reps = nCopies + 1;
      if (flag[j]) dDup = start + (nCopies + 1) * nb + (offs[j] - Dsb);
    }
    for (int m = 0; m <= reps; ++m) {
      int d = (m < reps) ? d0 + m * stride : dDup;
      if (d < 0) continue;
      if (d >= dstCap) { atomicMax(&cdw->error, SB_ERR_BANK_OVERFLOW); continue; }
      dst.rx[d] = src.rx[j]; dst.ry[d] = src.ry[j]; dst.rz[d] = src.rz[j];
      dst.ux[d] = src.ux[j]; dst.uy[d] = src.uy[j]; dst.uz[d] = src.uz[j];
      dst.w[d] = src.w[j]; dst.G[d] = src.G[j]; dst.E[d] = src.E[j]; dst.brood[d] = src.brood[j]; dst.seq[d] = 0;
    }
  }
}
__global__ void k_norm_count(const int* flag, const int* offs, const NormDev* nd, CycleDev* cdw) {
  PDL_ENTER();
  int n = nd->nLocal;
  if (n <= 0) { cdw->nNew = 0; return; }
  int excess = nd->nGlobal - nd->totPop;
  int selected = offs[n - 1] + flag[n - 1];
  int nCopies = (excess < 0) ? (-excess) / nd->nGlobal : 0;
  int nNew = (excess > 0) ? selected : (excess == 0 ? n : n * (nCopies + 1) + selected);
  cdw->nNew = nNew;
  if (nd->check && nNew != nd->totPop) atomicMax(&cdw->error, SB_ERR_NORM);      // "Normalisation failed!" (:596)
}

// number of flagged sites (kept for excess > 0, duplicated for excess < 0) in every rank's slice of the global stream:
// every rank can then compute every rank's new bank size itself (replaces the mpi_allgather of :593)
struct RankOffs { int n; int off[65]; };
__global__ void k_norm_rank_counts(const unsigned long long* rn, const CycleDev* cd, const NormDev* nd, const RankOffs ro, int* counts) {
  PDL_ENTER();
  __shared__ int sc[64];
  if (threadIdx.x < 64) sc[threadIdx.x] = 0;
  __syncthreads();
  const int n = nd->nGlobal;
  const int excess = n - nd->totPop;
  const int nDup = (excess < 0) ? (int)(((long long)(-excess)) % n) : 0;
  const double thr = cd->thrReal;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    double x = (double)(long long)rn[j] * (1.0 / 9223372036854775808.0);
    int f;
    if (excess > 0) f = (x > thr) ? 1 : 0;
    else if (excess < 0) f = (nDup != 0 && x <= thr) ? 1 : 0;
    else f = 1;
    if (f) { int r = 0; while (r + 1 < ro.n && j >= ro.off[r + 1]) ++r; atomicAdd(&sc[r], 1); }
  }
  __syncthreads();
  if (threadIdx.x < ro.n && sc[threadIdx.x]) atomicAdd(&counts[threadIdx.x], sc[threadIdx.x]);
}

// loadBalancing (particleDungeon_class.f90:607-698): pack sites of the ends of the bank / rebuild the bank as
// [received from below] + kept middle + [received from above]. Packed site layout: 8 f64 arrays of k, then k i32 (G), k i32 (broodID)
__global__ void k_bank_pack_range(Bank b, int first, int k, double* buf) {
  int* g = (int*)(buf + 8 * (size_t)k);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) {
    int s = first + i;
    buf[i] = b.rx[s]; buf[k + i] = b.ry[s]; buf[2 * (size_t)k + i] = b.rz[s];
    buf[3 * (size_t)k + i] = b.ux[s]; buf[4 * (size_t)k + i] = b.uy[s]; buf[5 * (size_t)k + i] = b.uz[s];
    buf[6 * (size_t)k + i] = b.w[s]; buf[7 * (size_t)k + i] = b.E[s]; g[i] = b.G[s]; g[k + i] = b.brood[s];
  }
}
__global__ void k_bank_unpack_range(Bank b, int first, int k, const double* buf) {
  const int* g = (const int*)(buf + 8 * (size_t)k);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) {
    int s = first + i;
    b.rx[s] = buf[i]; b.ry[s] = buf[k + i]; b.rz[s] = buf[2 * (size_t)k + i];
    b.ux[s] = buf[3 * (size_t)k + i]; b.uy[s] = buf[4 * (size_t)k + i]; b.uz[s] = buf[5 * (size_t)k + i];
    b.w[s] = buf[6 * (size_t)k + i]; b.E[s] = buf[7 * (size_t)k + i]; b.G[s] = g[i]; b.brood[s] = g[k + i]; b.seq[s] = 0;
  }
}
__global__ void k_bank_copy_range(Bank src, int first, int k, Bank dst, int dfirst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) {
    int s = first + i, d = dfirst + i;
    dst.rx[d] = src.rx[s]; dst.ry[d] = src.ry[s]; dst.rz[d] = src.rz[s];
    dst.ux[d] = src.ux[s]; dst.uy[d] = src.uy[s]; dst.uz[d] = src.uz[s];
    dst.w[d] = src.w[s]; dst.G[d] = src.G[s]; dst.E[d] = src.E[s]; dst.brood[d] = src.brood[s]; dst.seq[d] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// shannonEntropyClerk (Tallies/TallyClerks/shannonEntropyClerk_class.f90): reportCycleEnd bins the weights of the cycle's fission
// bank (before normSize_Repr) over the clerk's map; closeCycle turns them into -sum p log2 p, accumulates it in the bin of the
// current cycle and resets the weight bins.  Bank sites carry no material (particleState%matIdx stays -1): a materialMap sends
// them to its default bin, as in the reference.
// ------------------------------------------------------------------------------------------------
__global__ void k_shannon_score(const Model M, const char* blob, int phase, int clerkIdx, Bank sites, const CycleDev* cd, int cap, int isCE, double* bins) {
  PDL_ENTER();
  __shared__ double sTot[32];
  const DClerk& c = ((const DClerk*)(blob + M.oClerk[phase]))[clerkIdx];
  const int n = min(cd->nSites, cap);
  double tot = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double r[3] = {sites.rx[i], sites.ry[i], sites.rz[i]};
    const double w = sites.w[i];
    tot += w;
    const int bin = isCE ? sbc::clerkBinCE(c, blob, r, -1, sites.E[i]) : clerkBin(c, blob, r, -1);
    if (bin > 0) atomicAdd(&bins[c.addr - 1 + bin], w);
  }
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_down_sync(0xffffffffu, tot, o);
  if ((threadIdx.x & 31) == 0) sTot[threadIdx.x >> 5] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sTot[k];
    if (t != 0.0) atomicAdd(&bins[c.addr - 1], t);
  }
}
__global__ void __launch_bounds__(256) k_shannon_close(int addr, int nBins, int cycle, double* bins, double* csum, double* csum2) {
  PDL_ENTER();
  __shared__ double sVal[8];
  const double totWgt = bins[addr - 1];
  const double one_log2 = 1.0 / sbm::log(2.0);
  double val = 0.0;
  for (int i = threadIdx.x; i < nBins; i += blockDim.x) {
    const double prob = bins[addr + i] / totWgt;
    if (prob > 0.0 && prob < 1.0) val = val - prob * sbm::log(prob) * one_log2;
  }
  for (int o = 16; o > 0; o >>= 1) val += __shfl_down_sync(0xffffffffu, val, o);
  if ((threadIdx.x & 31) == 0) sVal[threadIdx.x >> 5] = val;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int k = 0; k < 8; ++k) v += sVal[k];
    const int cc = addr - 1 + nBins + cycle;                  // 0-based position of getMemAddress() + N + currentCycle
    csum[cc] = csum[cc] + v; csum2[cc] = csum2[cc] + v * v;   // scoreMemory%accumulate
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= nBins; i += blockDim.x) bins[addr - 1 + i] = 0.0;      // resetBin
}

// ------------------------------------------------------------------------------------------------
// Ranks of one node exchanging through peer memory (NVLink / NVSwitch; CUDA IPC between the processes), no host in the loop.
// What SCONE moves over MPI per cycle is tiny: 6 score sums + the bank size of every rank (scoreMemory%reduceBins,
// particleDungeon_class.f90:464,516-518,593) and the sites that loadBalancing hands to the neighbours (:607-698).  Every rank owns
// one PeerBox in its own memory; the others WRITE into it (posted stores, then a release flag) and the owner spins on its local
// flags - no rank ever reads remote memory.  Slots are double-buffered by cycle parity: the all-to-all of the sums is a barrier,
// so a rank is never more than one cycle ahead of another and a slot is rewritten only after its reader has consumed it.
// ------------------------------------------------------------------------------------------------
constexpr int CE_PIPE = 8;            // chunks in flight of the host-buffer CE lookup
constexpr int PEER_MAX = 64;
struct PeerBox {
  double data[2][PEER_MAX][8];                   // [parity][source rank]: 6 score sums, bank size, 0
  unsigned long long flagSums[2][PEER_MAX];      // cycle sequence number of the data above
  unsigned long long flagSites[2][2];            // [parity][0 = from the rank below, 1 = from the rank above]
  unsigned long long pad[4];
};
struct PeerPtrs { PeerBox* box[PEER_MAX]; char* stage[PEER_MAX]; };     // stage: 4 site buffers {below,above} x parity of every rank
struct PeerPlan {            // what the host would have computed from the gathered sizes, left on the device
  int n, rank, nGlobal, offLocal;
  int popSizes[PEER_MAX], off[PEER_MAX + 1], newSizes[PEER_MAX], finalSizes[PEER_MAX];
  int sendUp, recvUp, sendDown, recvDown, finalLocal, ok;
};
__device__ __forceinline__ unsigned long long ldAcquireSys(const unsigned long long* p) {
  unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void stReleaseSys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globalTimerNs() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ bool spinUntil(const unsigned long long* flag, unsigned long long seq, unsigned long long timeoutNs) {
  const unsigned long long t0 = globalTimerNs();
  for (;;) {
    if (ldAcquireSys(flag) >= seq) return true;
    if (globalTimerNs() - t0 > timeoutNs) return false;
    __nanosleep(200);
  }
}
// all-to-all of { 6 sums, bank size } + the barrier it implies; sums added in rank order (same bits on every rank)
__global__ void k_peer_post_wait(PeerPtrs P, int nRanks, int rank, unsigned long long seq, const double* own, double* total, PeerPlan* plan,
                                 CycleDev* cd, int cap, unsigned long long timeoutNs) {
  PDL_ENTER();
  __shared__ int sOk[PEER_MAX];
  const int par = (int)(seq & 1ULL), t = threadIdx.x;
  if (t < nRanks) {
    PeerBox* dst = P.box[t];
    for (int k = 0; k < 8; ++k) dst->data[par][rank][k] = own[k];
    __threadfence_system();
    stReleaseSys(&dst->flagSums[par][rank], seq);
    sOk[t] = spinUntil(&P.box[rank]->flagSums[par][t], seq, timeoutNs) ? 1 : 0;
  }
  __syncthreads();
  if (t == 0) {
    const PeerBox* me = P.box[rank];
    bool ok = true;
    for (int r = 0; r < nRanks; ++r) ok = ok && sOk[r];
    double sum[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    long long tot = 0;
    plan->n = nRanks; plan->rank = rank;
    for (int r = 0; r < nRanks; ++r) {
      const volatile double* d = me->data[par][r];
      for (int k = 0; k < 6; ++k) sum[k] = sum[k] + (ok ? d[k] : (r == rank ? own[k] : 0.0));
      int sz = ok ? (int)d[6] : (r == rank ? (int)own[6] : 0);
      plan->popSizes[r] = sz; plan->off[r] = (int)tot; tot += sz;
    }
    plan->off[nRanks] = (int)tot;
    if (!ok) atomicMax(&cd->error, SB_ERR_PEER_TIMEOUT);
    if (tot > 2000000000LL || tot <= 0) { atomicMax(&cd->error, SB_ERR_NORM); tot = max(1, min(cd->nSites, cap)); }
    plan->nGlobal = (int)tot; plan->offLocal = plan->off[rank]; plan->ok = ok ? 1 : 0;
    for (int k = 0; k < 6; ++k) total[k] = sum[k];
    total[6] = (double)tot; total[7] = 0.0;
  }
}
__global__ void k_norm_setup_plan(NormDev* nd, const CycleDev* cd, int cap, int totPop, const PeerPlan* plan) {
  PDL_ENTER();
  nd->nLocal = min(cd->nSites, cap); nd->totPop = totPop; nd->check = 0;
  nd->nGlobal = plan->nGlobal; nd->offLocal = plan->offLocal;
}
__global__ void k_norm_rank_counts_plan(const unsigned long long* rn, const CycleDev* cd, const NormDev* nd, const PeerPlan* plan, int* counts) {
  PDL_ENTER();
  __shared__ int sc[PEER_MAX]; __shared__ int so[PEER_MAX + 1];
  const int nr = plan->n;
  if (threadIdx.x < PEER_MAX) sc[threadIdx.x] = 0;
  if (threadIdx.x <= nr) so[threadIdx.x] = plan->off[threadIdx.x];
  __syncthreads();
  const int n = nd->nGlobal;
  const int excess = n - nd->totPop;
  const int nDup = (excess < 0) ? (int)(((long long)(-excess)) % n) : 0;
  const double thr = cd->thrReal;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    double x = (double)(long long)rn[j] * (1.0 / 9223372036854775808.0);
    int f;
    if (excess > 0) f = (x > thr) ? 1 : 0;
    else if (excess < 0) f = (nDup != 0 && x <= thr) ? 1 : 0;
    else f = 1;
    if (f) { int r = 0; while (r + 1 < nr && j >= so[r + 1]) ++r; atomicAdd(&sc[r], 1); }
  }
  __syncthreads();
  if (threadIdx.x < nr && sc[threadIdx.x]) atomicAdd(&counts[threadIdx.x], sc[threadIdx.x]);
}
// every rank's size after normSize_Repr (what the mpi_allgather of :593 returns) and this rank's part of loadBalancing (:607-698)
__global__ void k_peer_plan(PeerPlan* plan, const int* counts, const NormDev* nd, CycleDev* cd, int cap, int stageCap) {
  PDL_ENTER();
  const int nr = plan->n, rank = plan->rank, totPop = nd->totPop;
  const long long tot = nd->nGlobal, excess = tot - totPop, nCopies = excess < 0 ? (-excess) / tot : 0;
  long long off1 = 0, off2 = 0, sum = 0;
  for (int i = 0; i < nr; ++i) {
    int ns = excess > 0 ? counts[i] : (excess == 0 ? plan->popSizes[i] : (int)(plan->popSizes[i] * (nCopies + 1) + counts[i]));
    plan->newSizes[i] = ns;
    if (i < rank) off1 += ns;
    sum += ns;
  }
  off2 = off1 + plan->newSizes[rank];
  if (sum != totPop || plan->newSizes[rank] != cd->nNew) atomicMax(&cd->error, SB_ERR_NORM);
  // getWorkshare / getOffset (mpi_func.f90:133-159)
  auto offsetOf = [&](int r) { return (long long)(totPop / nr) * r + max(0, totPop % nr + r - nr); };
  const long long t1 = offsetOf(rank), t2 = (rank + 1 == nr) ? totPop : offsetOf(rank + 1);
  const long long eEnd = off2 - t2, eBeg = off1 - t1;
  int sendUp = eEnd > 0 ? (int)eEnd : 0, recvUp = eEnd < 0 ? (int)(-eEnd) : 0;
  int sendDown = eBeg < 0 ? (int)(-eBeg) : 0, recvDown = eBeg > 0 ? (int)eBeg : 0;
  long long run = 0;
  for (int i = 0; i < nr; ++i) {                 // every rank's size after the exchange with its two neighbours
    long long a = run, b = run + plan->newSizes[i];
    long long ta = offsetOf(i), tb = (i + 1 == nr) ? totPop : offsetOf(i + 1);
    plan->finalSizes[i] = (int)(plan->newSizes[i] - max(0LL, b - tb) + max(0LL, tb - b) - max(0LL, ta - a) + max(0LL, a - ta));
    if (max(0LL, b - tb) + max(0LL, ta - a) > plan->newSizes[i]) atomicMax(&cd->error, SB_ERR_BALANCE);     // nearest neighbours cannot fix this
    run = b;
  }
  const int fin = plan->newSizes[rank] - sendUp - sendDown + recvUp + recvDown;
  if (fin > cap || max(max(sendUp, sendDown), max(recvUp, recvDown)) > stageCap) { atomicMax(&cd->error, SB_ERR_BANK_OVERFLOW); sendUp = recvUp = sendDown = recvDown = 0; }
  if (cd->error != 0) { sendUp = recvUp = sendDown = recvDown = 0; }
  plan->sendUp = sendUp; plan->recvUp = recvUp; plan->sendDown = sendDown; plan->recvDown = recvDown;
  plan->finalLocal = plan->newSizes[rank] - sendUp - sendDown + recvUp + recvDown;
}
// packed site layout of k sites with capacity `cap`: 8 f64 arrays of cap, then cap i32 (G), cap i32 (broodID)
__device__ __forceinline__ void packSite(const Bank& b, int s, double* buf, int cap, int i) {
  int* g = (int*)(buf + 8 * (size_t)cap);
  buf[i] = b.rx[s]; buf[(size_t)cap + i] = b.ry[s]; buf[2 * (size_t)cap + i] = b.rz[s];
  buf[3 * (size_t)cap + i] = b.ux[s]; buf[4 * (size_t)cap + i] = b.uy[s]; buf[5 * (size_t)cap + i] = b.uz[s];
  buf[6 * (size_t)cap + i] = b.w[s]; buf[7 * (size_t)cap + i] = b.E[s]; g[i] = b.G[s]; g[cap + i] = b.brood[s];
}
__device__ __forceinline__ void unpackSite(Bank& b, int s, const double* buf, int cap, int i) {
  const int* g = (const int*)(buf + 8 * (size_t)cap);
  b.rx[s] = buf[i]; b.ry[s] = buf[(size_t)cap + i]; b.rz[s] = buf[2 * (size_t)cap + i];
  b.ux[s] = buf[3 * (size_t)cap + i]; b.uy[s] = buf[4 * (size_t)cap + i]; b.uz[s] = buf[5 * (size_t)cap + i];
  b.w[s] = buf[6 * (size_t)cap + i]; b.E[s] = buf[7 * (size_t)cap + i]; b.G[s] = g[i]; b.brood[s] = g[cap + i]; b.seq[s] = 0;
}
__host__ __device__ inline size_t peerStageBytes(int cap) { return ((size_t)cap * (8 * sizeof(double) + 2 * sizeof(int)) + 255) / 256 * 256; }
// stage buffer of rank r: index 2*parity + (0 = sites arriving from below, 1 = from above)
__device__ __forceinline__ double* peerStage(const PeerPtrs& P, int r, int par, int fromAbove, int cap) {
  return (double*)(P.stage[r] + (size_t)(2 * par + fromAbove) * peerStageBytes(cap));
}
// the sites this rank gives away go straight into the neighbours' stage buffers (stores over NVLink)
__global__ void k_peer_push(PeerPtrs P, const PeerPlan* plan, Bank src, int cap, int par) {
  PDL_ENTER();
  const int rank = plan->rank, nLocal = plan->newSizes[rank];
  const int up = plan->sendUp, down = plan->sendDown;
  double* bufUp = up > 0 ? peerStage(P, rank + 1, par, 0, cap) : nullptr;          // arrives "from below" at rank + 1
  double* bufDown = down > 0 ? peerStage(P, rank - 1, par, 1, cap) : nullptr;      // arrives "from above" at rank - 1
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < up + down; i += gridDim.x * blockDim.x) {
    if (i < up) packSite(src, nLocal - up + i, bufUp, cap, i);
    else packSite(src, i - up, bufDown, cap, i - up);
  }
}
__global__ void k_peer_flag_sites(PeerPtrs P, const PeerPlan* plan, unsigned long long seq) {
  PDL_ENTER();
  const int rank = plan->rank, par = (int)(seq & 1ULL);
  __threadfence_system();
  if (threadIdx.x == 0 && plan->sendUp > 0) stReleaseSys(&P.box[rank + 1]->flagSites[par][0], seq);
  if (threadIdx.x == 1 && plan->sendDown > 0) stReleaseSys(&P.box[rank - 1]->flagSites[par][1], seq);
}
__global__ void k_peer_wait_sites(PeerPtrs P, const PeerPlan* plan, unsigned long long seq, CycleDev* cd, unsigned long long timeoutNs) {
  PDL_ENTER();
  const int rank = plan->rank, par = (int)(seq & 1ULL);
  bool ok = true;
  if (threadIdx.x == 0 && plan->recvDown > 0) ok = spinUntil(&P.box[rank]->flagSites[par][0], seq, timeoutNs);
  if (threadIdx.x == 1 && plan->recvUp > 0) ok = spinUntil(&P.box[rank]->flagSites[par][1], seq, timeoutNs);
  if (!ok) atomicMax(&cd->error, SB_ERR_PEER_TIMEOUT);
}
// the bank after loadBalancing: [received from below] + kept middle + [received from above]
__global__ void k_peer_splice(PeerPtrs P, const PeerPlan* plan, Bank src, Bank dst, int cap, int par, const CycleDev* cd) {
  PDL_ENTER();
  const int rank = plan->rank;
  const bool bad = cd->error != 0;
  const int addF = bad ? 0 : plan->recvDown, addB = bad ? 0 : plan->recvUp, dropF = plan->sendDown, dropB = plan->sendUp;
  const int keep = plan->newSizes[rank] - dropF - dropB;
  const double* bufF = peerStage(P, rank, par, 0, cap);
  const double* bufB = peerStage(P, rank, par, 1, cap);
  const int n = addF + keep + addB;
  for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < n; d += gridDim.x * blockDim.x) {
    if (d < addF) unpackSite(dst, d, bufF, cap, d);
    else if (d < addF + keep) {
      int s = dropF + (d - addF);
      dst.rx[d] = src.rx[s]; dst.ry[d] = src.ry[s]; dst.rz[d] = src.rz[s];
      dst.ux[d] = src.ux[s]; dst.uy[d] = src.uy[s]; dst.uz[d] = src.uz[s];
      dst.w[d] = src.w[s]; dst.G[d] = src.G[s]; dst.E[d] = src.E[s]; dst.brood[d] = src.brood[s]; dst.seq[d] = 0;
    } else unpackSite(dst, d, bufB, cap, d - addF - keep);
  }
}

// ------------------------------------------------------------------------------------------------
// fissionSource (ParticleObjects/Source/fissionSource_class.f90:149-271, source_inter.f90:98-118)
// ------------------------------------------------------------------------------------------------
__global__ void k_source(const Model M, const char* blob, Bank out, int n, uint64_t rng0, int offset,
                         double b0, double b1, double b2, double t0, double t1, double t2, CycleDev* cd) {
  const Tables T = bind(M, blob);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint64_t rng = rng_skip(rng0, RNG_STRIDE * (int64_t)(offset + i + 1));
    const double bottom[3] = {b0, b1, b2}, top[3] = {t0, t1, t2};
    bool ok = false;
    for (int att = 0; att < 10000 && !ok; ++att) {
      double r3[3]; r3[0] = rng_get(rng); r3[1] = rng_get(rng); r3[2] = rng_get(rng);
      double r[3], u[3] = {1.0, 0.0, 0.0};
      for (int k = 0; k < 3; ++k) r[k] = (top[k] - bottom[k]) * r3[k] + bottom[k];
      int mat, uid;
      geomPlace(M, T, r, u, mat, uid);
      if (mat == SB_VOID_MAT || mat == SB_OUTSIDE_MAT) continue;
      if (mat == SB_UNDEF_MAT) { atomicMax(&cd->error, SB_ERR_UNDEF_MAT); break; }
      if (mat == SB_OVERLAP_MAT) { atomicMax(&cd->error, SB_ERR_OVERLAP_MAT); break; }
      if (!T.fissile[mat - 1]) continue;
      double mu, phi;
      int Gout = mgFissionSample(M, T, mat, mu, phi, rng);
      if (Gout == 0) { atomicMax(&cd->error, SB_ERR_SAMPLING); Gout = 1; }
      double d[3] = {1.0, 0.0, 0.0};
      rotateVector(d, mu, phi);
      out.rx[i] = r[0]; out.ry[i] = r[1]; out.rz[i] = r[2];
      out.ux[i] = d[0]; out.uy[i] = d[1]; out.uz[i] = d[2];
      out.w[i] = 1.0; out.G[i] = Gout; out.E[i] = 0.0; out.brood[i] = 0; out.seq[i] = 0;
      ok = true;
    }
    if (!ok) atomicMax(&cd->error, SB_ERR_SOURCE);
  }
}

// pointSource%sampleParticle (ParticleObjects/Source/pointSource_class.f90:142-200; configSource_inter.f90:75-90: type, position,
// energy-angle, energy): particle i draws from rng0 skipped by stride*(offset + i + 1) as source%generate does (source_inter.f90:98-118)
__global__ void k_source_point(Bank out, int n, uint64_t rng0, int offset, double r0, double r1, double r2, double u0, double u1, double u2,
                               int isotropic, int isMG, double E, int G, int nProb, const double* probG) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint64_t rng = rng_skip(rng0, RNG_STRIDE * (int64_t)(offset + i + 1));
    double d[3] = {u0, u1, u2};
    if (isotropic) {
      double mu = 2.0 * rng_get(rng) - 1.0;
      double phi = TWO_PI * rng_get(rng);
      d[0] = 1.0; d[1] = 0.0; d[2] = 0.0;
      rotateVector(d, mu, phi);
    }
    int g = G;
    if (isMG && nProb > 0) {
      double r = rng_get(rng);
      for (g = 1; g <= nProb; ++g) { r = r - probG[g - 1]; if (r < 0.0) break; }
    }
    out.rx[i] = r0; out.ry[i] = r1; out.rz[i] = r2;
    out.ux[i] = d[0]; out.uy[i] = d[1]; out.uz[i] = d[2];
    out.w[i] = 1.0; out.G[i] = isMG ? g : 0; out.E[i] = isMG ? 0.0 : E; out.brood[i] = 0; out.seq[i] = 0;
  }
}

// materialSource%sampleParticle (ParticleObjects/Source/materialSource_class.f90:136-210)
__global__ void k_source_material(const Model M, const char* blob, Bank out, int n, uint64_t rng0, int offset, sb_material_source S, CycleDev* cd) {
  const Tables T = bind(M, blob);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint64_t rng = rng_skip(rng0, RNG_STRIDE * (int64_t)(offset + i + 1));
    bool ok = false;
    for (int att = 1; att <= 200 && !ok; ++att) {
      double r3[3]; r3[0] = rng_get(rng); r3[1] = rng_get(rng); r3[2] = rng_get(rng);
      double r[3], u[3] = {1.0, 0.0, 0.0};
      for (int k = 0; k < 3; ++k) r[k] = (S.top[k] - S.bottom[k]) * r3[k] + S.bottom[k];
      (void)rng_get(rng);                                       // time = tLow + rand * (tHigh - tLow): drawn in every attempt
      int mat, uid;
      geomPlace(M, T, r, u, mat, uid);
      if (mat == SB_OUTSIDE_MAT) continue;
      if (mat >= SB_OVERLAP_MAT) { atomicMax(&cd->error, mat == SB_VOID_MAT ? SB_ERR_MAT_SOURCE_VOID : (mat == SB_UNDEF_MAT ? SB_ERR_UNDEF_MAT : SB_ERR_OVERLAP_MAT)); ok = true; break; }
      if (mat != S.mat_idx) continue;
      double mu = 2.0 * rng_get(rng) - 1.0;
      double phi = TWO_PI * rng_get(rng);
      double d[3] = {1.0, 0.0, 0.0};
      rotateVector(d, mu, phi);
      out.rx[i] = r[0]; out.ry[i] = r[1]; out.rz[i] = r[2];
      out.ux[i] = d[0]; out.uy[i] = d[1]; out.uz[i] = d[2];
      out.w[i] = 1.0; out.G[i] = S.is_mg ? S.G : 0; out.E[i] = S.is_mg ? 0.0 : S.E; out.brood[i] = 0; out.seq[i] = 0;
      ok = true;
    }
    if (!ok) atomicMax(&cd->error, SB_ERR_MAT_SOURCE);
  }
}
// fileSource%sampleParticle (ParticleObjects/Source/fileSource_class.f90:151-196)
__global__ void k_source_file(const Model M, const char* blob, Bank out, int n, uint64_t rng0, int offset, const double* rows, long long nRows, int isMG,
                              CycleDev* cd) {
  const Tables T = bind(M, blob);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint64_t rng = rng_skip(rng0, RNG_STRIDE * (int64_t)(offset + i + 1));
    long long idx = (long long)(rng_get(rng) * (double)nRows);
    if (idx >= nRows) { atomicMax(&cd->error, SB_ERR_FILE_SOURCE); idx = nRows - 1; }
    const double* row = rows + 10 * idx;
    double r[3] = {row[0], row[1], row[2]}, u[3] = {row[3], row[4], row[5]};
    int mat, uid;
    geomPlace(M, T, r, u, mat, uid);
    if (mat == SB_OUTSIDE_MAT || mat == SB_UNDEF_MAT) atomicMax(&cd->error, SB_ERR_FILE_SOURCE);
    out.rx[i] = row[0]; out.ry[i] = row[1]; out.rz[i] = row[2];
    out.ux[i] = row[3]; out.uy[i] = row[4]; out.uz[i] = row[5];
    out.w[i] = row[9]; out.G[i] = isMG ? (int)row[7] : 0; out.E[i] = isMG ? 0.0 : row[6]; out.brood[i] = 0; out.seq[i] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// batch query kernels (parity tests)
// ------------------------------------------------------------------------------------------------
__global__ void k_geom_query(const Model M, const char* blob, long long n, double* r, double* dir, const double* dist, int* mat, int* uid) {
  const Tables T = bind(M, blob);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double rr[3] = {r[3 * i], r[3 * i + 1], r[3 * i + 2]}, uu[3] = {dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]};
    int m, q;
    if (dist) geomTeleport(M, T, rr, uu, dist[i], m, q);
    else geomPlace(M, T, rr, uu, m, q);
    for (int k = 0; k < 3; ++k) { r[3 * i + k] = rr[k]; dir[3 * i + k] = uu[k]; }
    mat[i] = m; uid[i] = q;
  }
}
__global__ void k_mg_query(const Model M, const char* blob, long long n, const int* mat, const int* G, double* total, double* maj) {
  const Tables T = bind(M, blob);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    total[i] = mgRow(M, T, mat[i], G[i])[XS_TOTAL] + 0.0;
    maj[i] = mgMajorant(M, T, G[i]);
  }
}
__global__ void k_rng_query(long long n, const unsigned long long* st, const long long* skip, unsigned long long* outS, double* outR) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint64_t s = rng_skip(st[i], skip[i]);
    double x = rng_get(s);
    outS[i] = s; outR[i] = x;
  }
}
__global__ void k_math_query(long long n, const double* x, double* lg, double* sn, double* cs) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    lg[i] = sbm::log(x[i]);
    double s, c; sbm::sincos(x[i], &s, &c); sn[i] = s; cs[i] = c;
  }
}
// the merged-range division / square root of sb_device.cuh against the plain operators, on generated operands:
// a = 2^ea * (1 + fraction), b likewise, exponents within the fast range; counts results that differ in any bit
__global__ void k_fastmath_check(long long n, unsigned long long seed, int expSpan, unsigned long long* bad) {
  unsigned long long nb = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint64_t s = rng_skip(seed, 3 * i + 1);
    double fa = rng_get(s), fb = rng_get(s), fe = rng_get(s);
    int ea = (int)(fe * (2 * expSpan + 1)) - expSpan, eb = (int)((fe * 7919.0 - floor(fe * 7919.0)) * (2 * expSpan + 1)) - expSpan;
    double a = ldexp(1.0 + fa, ea), b = ldexp(1.0 + fb, eb);
    if (i & 1) a = -a;
    if (i % 5 == 0) b = 1.0 - fb * fb;                      // the operands rotateVector sees: 1 - mu^2, sums of squares near one
    if (i % 7 == 0) a = fa * fb;
    if (fastRange(a) && fastRange(b)) {
      double y = rcpRefined(b);
      if (__double_as_longlong(divBy(a, b, y)) != __double_as_longlong(a / b)) ++nb;
      if (b > 0.0 && __double_as_longlong(sqrtFast(b)) != __double_as_longlong(sqrt(b))) ++nb;
    }
  }
  if (nb) atomicAdd(bad, nb);
}

// fissionMG%sampleOut + rotateVector for the sites the delta-tracking history kernel left unfinished (sb_hist.cuh):
// slot = {r, parent direction, weight, G = material, E = bits of the stream state in front of the site's numbers}.
// fissionMG_class.f90:183-206 (mu, phi, then the chi walk), neutronMGstd_class.f90:131-199
__global__ void k_finish_sites(const Model M, const char* blob, Bank b, CycleDev* cd, int cap) {
  PDL_ENTER();
  const Tables T = bind(M, blob);
  const int n = min(cd->nSites, cap);
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    uint64_t rng = (uint64_t)__double_as_longlong(b.E[s]);
    const int mat = b.G[s];
    double mu, phi;
    int Gout = mgFissionSample(M, T, mat, mu, phi, rng);
    if (Gout == 0) { atomicMax(&cd->error, SB_ERR_SAMPLING); Gout = 1; }
    double d[3] = {b.ux[s], b.uy[s], b.uz[s]};
    rotateVector(d, mu, phi);
    b.ux[s] = d[0]; b.uy[s] = d[1]; b.uz[s] = d[2];
    b.G[s] = Gout; b.E[s] = 0.0;
  }
}

__global__ void k_cycle_begin(CycleDev* cd, int* nCur, int n) {
  cd->nStart = n; cd->nSites = 0; cd->nextHistory = 0; cd->error = 0;
  cd->selBin = -1; cd->selRank = 0; cd->nCand = 0; cd->nNew = 0;
  cd->nSeg = 0ULL; cd->nColl = 0ULL; cd->nScore = 0ULL; cd->nXsTerms = 0ULL; cd->maxSeg = 256;
  *nCur = n;
}
// particleState arrays cross the boundary as r(3,n), dir(3,n) (Fortran order); banks are SoA on the device
__global__ void k_bank_unpack(const double* r, const double* d, Bank b, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    b.rx[i] = r[3 * (size_t)i]; b.ry[i] = r[3 * (size_t)i + 1]; b.rz[i] = r[3 * (size_t)i + 2];
    b.ux[i] = d[3 * (size_t)i]; b.uy[i] = d[3 * (size_t)i + 1]; b.uz[i] = d[3 * (size_t)i + 2];
    b.brood[i] = 0; b.seq[i] = 0;
  }
}
__global__ void k_bank_pack(Bank b, double* r, double* d, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    r[3 * (size_t)i] = b.rx[i]; r[3 * (size_t)i + 1] = b.ry[i]; r[3 * (size_t)i + 2] = b.rz[i];
    d[3 * (size_t)i] = b.ux[i]; d[3 * (size_t)i + 1] = b.uy[i]; d[3 * (size_t)i + 2] = b.uz[i];
  }
}
__global__ void k_zero_int(int* p, int n) { PDL_ENTER(); for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0; }

// ================================================================================================
// host side: engine handle + C ABI
// ================================================================================================
struct sb_engine {
  int device = 0, numSM = 148;
  std::string err;
  cudaStream_t stream = nullptr;
  long long launches = 0;
  // model
  bool haveGeom = false, haveData = false;
  sb_geom_flat g{}; std::vector<int> gi_surfType, gi_cellOff, gi_cellSurf, gi_uniType, gi_uniIpar, gi_auxI, gi_gidx, gi_gid;
  std::vector<double> gd_surfPar, gd_uniDpar, gd_auxD;
  int nMat = 0, nG = 0, isP1 = 0; double collisionXS = 0.0;
  std::vector<double> xs, P0, prod, P1, chi, majorant; std::vector<int> fissile;
  std::vector<DClerk> clerks[2]; std::vector<std::vector<char>> clerkAux[2]; int nBins[2] = {0, 0}; int normAddr[2] = {0, 0}; double normVal[2] = {1.0, 1.0};
  std::vector<sb_clerk> clerkDefs[2]; std::vector<std::vector<double>> mapBounds[2]; std::vector<std::vector<int>> mapMat[2];
  Model M{}; char* dBlob = nullptr; bool blobDirty = true; int useSmem = 0, trackSmem = 0;
  sb_options opt{SB_TRACK_DT, 0.9, 1, 0, 0, 0};
  // banks
  int cap = 0; Bank bank[3]{}; int cur = 0;          // bank[cur] = this cycle; others: raw sites, sorted/next
  int nCur = 0;
  int *dNsites = nullptr, *dOffsets = nullptr, *dTile = nullptr, *dFlag = nullptr, *dFlagOff = nullptr, *dHist = nullptr;
  double *dHProd = nullptr, *dHAbs = nullptr, *dHLeak = nullptr, *dHScat = nullptr;
  unsigned long long *dRn = nullptr, *dCand = nullptr;
  RedOut* dPartial = nullptr;
  char* dHot = nullptr; HotLayout hot{}; ulonglong2* dSeedTab = nullptr;
  CycleDev* dCd = nullptr; CycleDev* hCd = nullptr; int* dNcur = nullptr;
  double *dBins[2] = {nullptr, nullptr}, *dLast[2] = {nullptr, nullptr}, *dCsum[2] = {nullptr, nullptr}, *dCsum2[2] = {nullptr, nullptr};
  int batchN[2] = {0, 0};
  double bounds[6] = {0, 0, 0, 0, 0, 0};
  double kNormNext = 1.0;   // nextCycle%k_eff of the dungeon that will receive the sites (keffAnalogClerk k_norm)
  bool sortedReady = false; int phaseOpen = -1; double kCumLast = 1.0;
  int* dRankCounts = nullptr; int* hRankCounts = nullptr;
  double* dKsum = nullptr; NormDev* dNd = nullptr; unsigned long long* dRnGlobal = nullptr; size_t rnGlobalCap = 0;
  int refillMin = 1;
  int maxSegMin = 256, loneMode = 1, cellCache = 48, assist = -1; unsigned laneMask = 0xffffffffu;
  sbh::LoneRec* dLoneQ = nullptr; int* dLoneCtl = nullptr; int* dLoneReady = nullptr; int loneCap = 0, loneTag = 0;      // queue between k_histories and k_lone; { count, next, done }
  int *dCePerm = nullptr, *dCeHist = nullptr, *dCeCursor = nullptr, *dCeTile = nullptr, *dCeNbin = nullptr; size_t cePermCap = 0, ceBinCap = 0;   // sorted CE lookups
  long long* dProfRounds = nullptr;
  // measurement
  bool profiling = false; cudaEvent_t evK0 = nullptr, evK1 = nullptr, evT0 = nullptr, evT1 = nullptr;
  cudaEvent_t evP1 = nullptr, evP2 = nullptr; double msPeerWait = 0.0, msPeerTail = 0.0; bool peerStagesOpen = false;
  double msHistories = 0.0; long long nHistLaunches = 0; long long segProfiled = 0, scoreProfiled = 0;
  double* dStage = nullptr; size_t stageBytes = 0;
  // continuous-energy transport model (sb_load_ce_model)
  char* dCeSlots = nullptr; size_t ceSlotCount = 0;
  UserKeff userKeff[2] = {};
  // fixed-source calculations: private secondary buffers of the lanes
  bool fixedSource = false; int stkCap = 50; double* dStkD = nullptr; int* dStkG = nullptr; size_t stkLanes = 0; int stkAllocCap = 0;
  // peer-memory exchange between the ranks of a node (sb_peer_*)
  int peerRanks = 0, peerRank = 0, peerCap = 0; char* peerRegion = nullptr; PeerPtrs peerPtrs{}; void* peerOpened[PEER_MAX] = {};
  double* dKsumRed = nullptr;          // score sums reduced over the ranks by the caller (sb_cycle_end_resample_ranked)
  PeerPlan* dPlan = nullptr; PeerPlan* hPlan = nullptr; double* dKsumTot = nullptr; unsigned long long peerSeq = 0; double peerTimeoutS = 20.0;
  struct ShannonRec { int clerk, addr, nBins, maxCycles, cycle; };
  std::vector<ShannonRec> shannon[2];     // shannonEntropyClerk records of each phase; cycle = reportCycleEnd calls so far
  cudaStream_t ceStreamIn = nullptr, ceStreamOut = nullptr; cudaEvent_t ceEvIn[CE_PIPE] = {}, ceEvK[CE_PIPE] = {};      // sb_ce_lookup pipeline
  bool broodValid = false;       // the current bank came out of a cycle (its sites have parents); false for source / uploaded banks
  double* dFileSrc = nullptr; long long nFileSrc = 0; bool fileSrcMG = false;     // fileSource rows (printToFile records)
  bool ceMode = false; sbc::CeModelDev ceModel{}; std::vector<sbk::CardOut> ceCards; std::vector<void*> ceAllocs;
  sbce::CeHost ce; int* dCeErr = nullptr; cudaEvent_t evC0 = nullptr, evC1 = nullptr; float ceLastMs = 0.f;
};

static std::string g_globalErr;

static int allocBank(sb_engine* h, Bank& b, int cap) {
  CUDA_OK(cudaMalloc(&b.rx, sizeof(double) * cap)); CUDA_OK(cudaMalloc(&b.ry, sizeof(double) * cap)); CUDA_OK(cudaMalloc(&b.rz, sizeof(double) * cap));
  CUDA_OK(cudaMalloc(&b.ux, sizeof(double) * cap)); CUDA_OK(cudaMalloc(&b.uy, sizeof(double) * cap)); CUDA_OK(cudaMalloc(&b.uz, sizeof(double) * cap));
  CUDA_OK(cudaMalloc(&b.w, sizeof(double) * cap)); CUDA_OK(cudaMalloc(&b.G, sizeof(int) * cap));
  CUDA_OK(cudaMalloc(&b.E, sizeof(double) * cap)); CUDA_OK(cudaMemset(b.E, 0, sizeof(double) * cap));
  CUDA_OK(cudaMalloc(&b.brood, sizeof(int) * cap)); CUDA_OK(cudaMalloc(&b.seq, sizeof(int) * cap));
  return 0;
}
static void freeBank(Bank& b) {
  cudaFree(b.rx); cudaFree(b.ry); cudaFree(b.rz); cudaFree(b.ux); cudaFree(b.uy); cudaFree(b.uz); cudaFree(b.w); cudaFree(b.G); cudaFree(b.E); cudaFree(b.brood); cudaFree(b.seq);
  b = Bank{};
}

static int ensureCapacity(sb_engine* h, int maxPop) {
  int cap = 2 * maxPop;                      // eigenPhysicsPackage_class.f90:355-356 dungeons of 2*pop
  if (cap <= h->cap) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  for (int i = 0; i < 3; ++i) { freeBank(h->bank[i]); if (allocBank(h, h->bank[i], cap)) return -1; }
  cudaFree(h->dNsites); cudaFree(h->dOffsets); cudaFree(h->dTile); cudaFree(h->dFlag); cudaFree(h->dFlagOff);
  cudaFree(h->dHProd); cudaFree(h->dHAbs); cudaFree(h->dHLeak); cudaFree(h->dHScat); cudaFree(h->dRn);
  CUDA_OK(cudaMalloc(&h->dNsites, sizeof(int) * cap)); CUDA_OK(cudaMalloc(&h->dOffsets, sizeof(int) * cap));
  CUDA_OK(cudaMalloc(&h->dTile, sizeof(int) * (cap / SCAN_TILE + 2)));
  CUDA_OK(cudaMalloc(&h->dFlag, sizeof(int) * cap)); CUDA_OK(cudaMalloc(&h->dFlagOff, sizeof(int) * cap));
  CUDA_OK(cudaMalloc(&h->dHProd, sizeof(double) * cap)); CUDA_OK(cudaMalloc(&h->dHAbs, sizeof(double) * cap));
  CUDA_OK(cudaMalloc(&h->dHLeak, sizeof(double) * cap)); CUDA_OK(cudaMalloc(&h->dHScat, sizeof(double) * cap));
  CUDA_OK(cudaMalloc(&h->dRn, sizeof(unsigned long long) * cap));
  CUDA_OK(cudaDeviceSynchronize());            // allocBank's null-stream memsets land before the engine stream uses the banks
  h->cap = cap; h->cur = 0; h->nCur = 0;
  return 0;
}

template <typename T>
static int put(std::vector<char>& blob, const std::vector<T>& v) {
  while (blob.size() % 16) blob.push_back(0);
  int off = (int)blob.size();
  const char* p = (const char*)v.data();
  blob.insert(blob.end(), p, p + sizeof(T) * v.size());
  return off;
}

// (re)build the device tables from the loaded model:
//   generic blob = [Model header][tables]  (global memory; query kernels, source kernel, cold paths)
//   hot blob     = compact records of sb_hist.cuh (staged into shared memory by the history kernel)
static int buildBlob(sb_engine* h) {
  if (!h->blobDirty) return 0;
  if (!h->haveGeom || !h->haveData) { h->err = "geometry and nuclear data must be loaded first"; return -1; }
  Model& M = h->M;
  std::vector<char> blob(sizeof(Model), 0);
  M.nSurf = h->g.n_surf; M.nCell = h->g.n_cell; M.nUni = h->g.n_uni; M.nGraph = h->g.n_graph;
  M.rootIdx = h->g.root_idx; M.borderIdx = h->g.border_idx;
  for (int i = 0; i < 6; ++i) M.bc[i] = h->g.bc[i];
  M.oSurfType = put(blob, h->gi_surfType); M.oSurfPar = put(blob, h->gd_surfPar);
  M.oCellOff = put(blob, h->gi_cellOff); M.oCellSurf = put(blob, h->gi_cellSurf);
  M.oUniType = put(blob, h->gi_uniType); M.oUniIpar = put(blob, h->gi_uniIpar); M.oUniDpar = put(blob, h->gd_uniDpar);
  M.oAuxD = put(blob, h->gd_auxD); M.oAuxI = put(blob, h->gi_auxI);
  std::vector<int> graph(2 * (size_t)h->g.n_graph);
  for (int i = 0; i < h->g.n_graph; ++i) { graph[2 * i] = h->gi_gidx[i]; graph[2 * i + 1] = h->gi_gid[i]; }
  M.oGraph = put(blob, graph);
  M.nMat = h->nMat; M.nG = h->nG; M.isP1 = h->isP1; M.collisionXS = h->collisionXS;
  M.oXs = put(blob, h->xs); M.oP0 = put(blob, h->P0); M.oProd = put(blob, h->prod);
  M.oP1 = h->isP1 ? put(blob, h->P1) : M.oP0;
  M.oChi = put(blob, h->chi); M.oFissile = put(blob, h->fissile); M.oMajorant = put(blob, h->majorant);
  for (int ph = 0; ph < 2; ++ph) {
    // per-map auxiliary tables first, then the clerk records that point at them
    std::vector<DClerk> cl = h->clerks[ph];
    for (size_t c = 0; c < cl.size(); ++c)
      for (int m = 0; m < cl[c].nMaps; ++m) {
        size_t key = c * SB_MAX_MAPS + m;
        if (cl[c].mapType[m] == SB_MAP_MATERIAL) cl[c].mapOff[m] = put(blob, h->mapMat[ph][key]);
        else if (cl[c].mapGrid[m] == SB_GRID_UNSTRUCT) cl[c].mapOff[m] = put(blob, h->mapBounds[ph][key]);
        else cl[c].mapOff[m] = 0;
      }
    M.nClerk[ph] = (int)cl.size(); M.nBins[ph] = h->nBins[ph];
    if (cl.empty()) cl.push_back(DClerk{});
    M.oClerk[ph] = put(blob, cl);
  }
  while (blob.size() % 16) blob.push_back(0);
  M.blobBytes = (int)blob.size();
  memcpy(blob.data(), &M, sizeof(Model));
  CUDA_OK(cudaSetDevice(h->device));
  cudaFree(h->dBlob);
  CUDA_OK(cudaMalloc(&h->dBlob, blob.size()));
  CUDA_OK(cudaMemcpy(h->dBlob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  h->trackSmem = (M.blobBytes <= 50 * 1024) ? 1 : 0;           // 4 CTAs of 128 threads per SM
  if (h->trackSmem) {
    CUDA_OK(cudaFuncSetAttribute(sbt::k_histories_track<128, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, M.blobBytes));
    CUDA_OK(cudaFuncSetAttribute(sbt::k_histories_track<512, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, M.blobBytes));
  }

  // ---- hot blob ----------------------------------------------------------------------------------
  {
    HotLayout& L = h->hot; L = HotLayout{};
    std::vector<char> hb;
    std::vector<HUni> unis((size_t)h->g.n_uni);
    for (int ui = 0; ui < h->g.n_uni; ++ui) {
      HUni& U = unis[ui]; memset(&U, 0, sizeof(U));
      const int* ip = &h->gi_uniIpar[(size_t)ui * SB_UNI_NIPAR]; const double* dp = &h->gd_uniDpar[(size_t)ui * SB_UNI_NDPAR];
      const int t = h->gi_uniType[ui];
      if (ip[0]) U.flags |= sbh::HF_ROT;
      if (ip[1]) U.flags |= sbh::HF_GLOBAL;
      for (int i = 0; i < 3; ++i) U.org[i] = dp[i];
      if (dp[0] == 0.0 && dp[1] == 0.0 && dp[2] == 0.0) U.flags |= sbh::HF_ORG0;
      U.type = sbh::HU_COLD;
      if (t == SB_UNI_LAT) {
        U.type = sbh::HU_LAT;
        for (int i = 0; i < 3; ++i) { U.ph[i].x = dp[12 + i]; U.ci[i].x = dp[15 + i]; U.ab[i] = dp[18 + i]; U.ci[i].y = 1.0 / dp[12 + i]; U.ph[i].y = 0.5 * dp[12 + i]; }
        U.n0 = ip[2]; U.n1 = ip[3]; U.n2 = ip[4]; U.outID = ip[5]; U.aux = ip[7];
        if (ip[6] == 1) U.flags |= sbh::HF_OFFALL; else if (ip[6] == 2) U.flags |= sbh::HF_OFFMAP;
        if (ip[4] == 1 && dp[14] >= 2.0 * INF && dp[17] == -INF) U.flags |= sbh::HF_LAT2D;
      } else if (t == SB_UNI_PIN) {
        U.type = sbh::HU_PIN; U.n0 = ip[2]; U.aux = ip[3];
      } else if (t == SB_UNI_ROOT) {
        int sidx = ip[2] - 1;
        if (h->gi_surfType[sidx] == SB_SURF_BOX) {
          const double* p = &h->gd_surfPar[(size_t)sidx * SB_SURF_NPAR];
          U.type = sbh::HU_ROOTBOX;
          for (int i = 0; i < 3; ++i) { U.ci[i].x = p[i]; U.ph[i].x = p[3 + i]; }
          U.ab[0] = p[6];
        }
      }
    }
    L.oUni = put(hb, unis); L.oGraph = put(hb, graph); L.oAuxD = put(hb, h->gd_auxD); L.oAuxI = put(hb, h->gi_auxI);
    L.oXs = put(hb, h->xs); L.oP0 = put(hb, h->P0); L.oProd = put(hb, h->prod); L.oP1 = h->isP1 ? put(hb, h->P1) : L.oP0;
    {
      std::vector<int> first((size_t)h->nMat * h->nG, h->nG);
      for (size_t row = 0; row < first.size(); ++row)
        for (int g = 0; g < h->nG; ++g) if (h->P0[row * h->nG + g] != 0.0) { first[row] = g; break; }
      L.oP0First = put(hb, first);
    }
    L.oFissile = put(hb, h->fissile);
    std::vector<double> majT(h->nG), majInv(h->nG);
    for (int g = 0; g < h->nG; ++g) { majT[g] = std::fmax(h->majorant[g] + 0.0, h->collisionXS); majInv[g] = 1.0 / majT[g]; }   // getTrackingXS(MAJORANT_XS)
    L.oMajT = put(hb, majT); L.oMajInv = put(hb, majInv);
    for (int ph = 0; ph < 2; ++ph) {
      // per-map auxiliary tables first, then the clerk records that point at them
      std::vector<DClerk> cl = h->clerks[ph];
      for (size_t c = 0; c < cl.size(); ++c)
        for (int m = 0; m < cl[c].nMaps; ++m) {
          size_t key = c * SB_MAX_MAPS + m;
          if (cl[c].mapType[m] == SB_MAP_MATERIAL) cl[c].mapOff[m] = put(hb, h->mapMat[ph][key]);
          else if (cl[c].mapGrid[m] == SB_GRID_UNSTRUCT) cl[c].mapOff[m] = put(hb, h->mapBounds[ph][key]);
          else cl[c].mapOff[m] = 0;
        }
      L.nClerk[ph] = (int)cl.size();
      // can any clerk of this phase score a non-zero value in (mat, G)?  (response values are functions of the XS row only)
      std::vector<unsigned char> mask((size_t)h->nMat * h->nG + 16, 0);
      for (int m = 0; m <= h->nMat; ++m)
        for (int g = 0; g < (m < h->nMat ? h->nG : 1); ++g) {
          bool any = false;
          for (const DClerk& k : cl)
            for (int i = 0; i < k.nResp; ++i) {
              if (k.respMT[i] == 0) any = true;
              else if (m < h->nMat) {
                const double* x = &h->xs[((size_t)m * h->nG + g) * 6]; const bool fis = h->fissile[m] != 0;
                double r = 0.0;
                switch (k.respMT[i]) {                       // neutronMacroXSs%get (neutronXsPackages_class.f90:143-190)
                  case -1: r = x[0]; break; case -2: r = x[2]; break; case -22: r = x[1] + (fis ? x[3] : 0.0) + x[2]; break;
                  case -4: case -20: r = x[1]; break; case -6: r = fis ? x[3] : 0.0; break; case -7: case -9: r = fis ? x[4] : 0.0; break;
                  case -80: r = fis ? x[5] : 0.0; break; case -21: r = (fis ? x[3] : 0.0) + x[2]; break; default: r = 0.0;
                }
                if (r != 0.0) any = true;
              }
            }
          mask[m < h->nMat ? (size_t)m * h->nG + g : (size_t)h->nMat * h->nG] = any ? 1 : 0;
        }
      L.oScoreMask[ph] = put(hb, mask);
      if (cl.empty()) cl.push_back(DClerk{});
      L.oClerk[ph] = put(hb, cl);
    }
    while (hb.size() % 16) hb.push_back(0);
    L.bytes = (int)hb.size();
    L.nG = h->nG; L.nMat = h->nMat; L.isP1 = h->isP1; L.rootIdx = h->g.root_idx; L.borderS = h->g.border_idx - 1;
    L.borderIsBox = h->gi_surfType[h->g.border_idx - 1] >= SB_SURF_BOX ? 1 : 0;
    {                                                   // the border is the box the root universe record holds: boundary transformations inline
      const int ri = h->g.root_idx - 1; const int* rip = &h->gi_uniIpar[(size_t)ri * SB_UNI_NIPAR];
      if (unis[ri].type == sbh::HU_ROOTBOX && rip[2] == h->g.border_idx && h->gi_surfType[h->g.border_idx - 1] == SB_SURF_BOX) {
        L.borderIsBox = 2;
        for (int i = 0; i < 6; ++i) L.bc[i] = h->g.bc[i];
        L.borderTol = h->gd_surfPar[(size_t)(h->g.border_idx - 1) * SB_SURF_NPAR + 6];
      }
    }
    cudaFree(h->dHot);
    CUDA_OK(cudaMalloc(&h->dHot, hb.size()));
    CUDA_OK(cudaMemcpy(h->dHot, hb.data(), hb.size(), cudaMemcpyHostToDevice));
    // shared-memory staging when two CTAs per SM fit beside each other (227 KB per SM on sm_100)
    h->useSmem = (L.bytes <= 72 * 1024) ? 1 : 0;
    {
      const int hotB = h->useSmem ? L.bytes : 0;
      CUDA_OK(cudaFuncSetAttribute(sbh::k_histories<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, hotB + sbh::histScratchBytes(256)));
      CUDA_OK(cudaFuncSetAttribute(sbh::k_histories<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, hotB + sbh::histScratchBytes(256)));
      CUDA_OK(cudaFuncSetAttribute(sbh::k_histories<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, hotB + sbh::histScratchBytes(256)));
      CUDA_OK(cudaFuncSetAttribute(sbh::k_histories<true, 1, 384>, cudaFuncAttributeMaxDynamicSharedMemorySize, hotB + sbh::histScratchBytes(384)));
      CUDA_OK(cudaFuncSetAttribute(sbh::k_histories<true, 1, 384, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, hotB + sbh::histScratchBytes(384)));
      CUDA_OK(cudaFuncSetAttribute(sbh::k_histories<true, 1, 384, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, hotB + sbh::histScratchBytes(384)));
      if (h->useSmem) CUDA_OK(cudaFuncSetAttribute(sbh::k_lone<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, hotB + sbh::loneScratchBytes(128)));
      if (h->useSmem) CUDA_OK(cudaFuncSetAttribute(sbh::k_lone<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, hotB + sbh::loneScratchBytes(128)));
      CUDA_OK(cudaFuncSetAttribute(sbh::k_histories<true, 1, 448>, cudaFuncAttributeMaxDynamicSharedMemorySize, hotB + sbh::histScratchBytes(448)));
      CUDA_OK(cudaFuncSetAttribute(sbh::k_histories<true, 1, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, hotB + sbh::histScratchBytes(512)));
      CUDA_OK(cudaFuncSetAttribute(sbh::k_histories<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, sbh::histScratchBytes(256)));
      if (getenv("SB_DEBUG_OCC")) {
        int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sbh::k_histories<true, 2>, 256, hotB + sbh::histScratchBytes(256));
        fprintf(stderr, "k_histories<true,2>: %d CTAs per SM with %d bytes of dynamic shared memory\n", nb, hotB + sbh::histScratchBytes(256));
      }
    }
  }
  for (int ph = 0; ph < 2; ++ph) {
    cudaFree(h->dBins[ph]); cudaFree(h->dLast[ph]); cudaFree(h->dCsum[ph]); cudaFree(h->dCsum2[ph]);
    size_t nb = (size_t)std::max(1, h->nBins[ph]);
    CUDA_OK(cudaMalloc(&h->dBins[ph], sizeof(double) * nb)); CUDA_OK(cudaMalloc(&h->dLast[ph], sizeof(double) * nb));
    CUDA_OK(cudaMalloc(&h->dCsum[ph], sizeof(double) * nb)); CUDA_OK(cudaMalloc(&h->dCsum2[ph], sizeof(double) * nb));
    CUDA_OK(cudaMemset(h->dBins[ph], 0, sizeof(double) * nb)); CUDA_OK(cudaMemset(h->dLast[ph], 0, sizeof(double) * nb));
    CUDA_OK(cudaMemset(h->dCsum[ph], 0, sizeof(double) * nb)); CUDA_OK(cudaMemset(h->dCsum2[ph], 0, sizeof(double) * nb));
    h->batchN[ph] = 0;
  }
  // the uploads and memsets above ran on the legacy null stream from pageable memory; the engine's stream is
  // non-blocking and does not order itself after them, so make them complete before any kernel can be enqueued
  CUDA_OK(cudaDeviceSynchronize());
  h->blobDirty = false;
  return 0;
}

static int gridFor(sb_engine* h, long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  long long cap = (long long)h->numSM * 8;
  if (b < 1) b = 1;
  return (int)(b < cap ? b : cap);
}

template <typename T>
static T* ceModelUpload(sb_engine* h, const std::vector<T>& v) {
  T* p = nullptr;
  if (cudaMalloc(&p, std::max<size_t>(sizeof(T) * v.size(), 16)) != cudaSuccess) { h->err = "cudaMalloc failed (CE model)"; return nullptr; }
  if (!v.empty() && cudaMemcpy(p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice) != cudaSuccess) { h->err = "cudaMemcpy failed (CE model)"; return nullptr; }
  h->ceAllocs.push_back(p);
  return p;
}
extern "C" {

int sb_create(sb_engine** out, int device) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { g_globalErr = "scone_b200: no CUDA device available (the engine has no CPU fallback)"; return -1; }
  if (device < 0 || device >= ndev) { g_globalErr = "scone_b200: invalid device ordinal"; return -1; }
  sb_engine* h = new sb_engine();
  h->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { g_globalErr = "cudaSetDevice failed"; delete h; return -1; }
  cudaDeviceProp p; cudaGetDeviceProperties(&p, device);
  h->numSM = p.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { g_globalErr = "cudaStreamCreate failed"; delete h; return -1; }
  cudaMalloc(&h->dCd, sizeof(CycleDev)); cudaMemset(h->dCd, 0, sizeof(CycleDev));
  cudaMallocHost(&h->hCd, sizeof(CycleDev));
  cudaMalloc(&h->dPartial, sizeof(RedOut) * RED_BLOCKS);
  {                                        // LCG jump table for per-history seeding: maps for stride*i*1024^level
    std::vector<ulonglong2> tab(3 * 1024 + 32);
    for (int lvl = 0; lvl < 3; ++lvl)
      for (int i = 0; i < 1024; ++i) {
        int64_t k = RNG_STRIDE * ((int64_t)i << (10 * lvl));
        uint64_t c = rng_skip(0ULL, k);                       // f^k(0) = C_k
        uint64_t g = (rng_skip(1ULL, k) - c) & RNG_MASK;     // f^k(1) - C_k = G_k
        tab[lvl * 1024 + i] = make_ulonglong2(g, c);
      }
    for (int j = 0; j < 32; ++j) {                            // j + 1 consecutive draws as one map (draw window of sb_hist.cuh)
      uint64_t c = rng_skip(0ULL, j + 1);
      tab[3 * 1024 + j] = make_ulonglong2((rng_skip(1ULL, j + 1) - c) & RNG_MASK, c);
    }
    cudaMalloc(&h->dSeedTab, sizeof(ulonglong2) * tab.size());
    cudaMemcpy(h->dSeedTab, tab.data(), sizeof(ulonglong2) * tab.size(), cudaMemcpyHostToDevice);
  }
  cudaMalloc(&h->dHist, sizeof(int) * SEL_BINS); cudaMalloc(&h->dCand, sizeof(unsigned long long) * SEL_CAND_CAP);
  cudaMalloc(&h->dNcur, sizeof(int)); cudaMalloc(&h->dKsum, 8 * sizeof(double)); cudaMalloc(&h->dNd, sizeof(NormDev));
  if (const char* e = getenv("SB_REFILL_MIN")) h->refillMin = std::max(1, atoi(e));     // tuning knobs (measurement only)
  if (const char* e = getenv("SB_BLOCKS_PER_SM")) h->opt.blocks_per_sm = atoi(e);
  if (const char* e = getenv("SB_MAXSEG_MIN")) h->maxSegMin = atoi(e);
  if (const char* e = getenv("SB_LONE_MODE")) h->loneMode = atoi(e);
  if (const char* e = getenv("SB_ASSIST")) h->assist = atoi(e);
  if (const char* e = getenv("SB_LANES")) { int k = atoi(e); h->laneMask = (k >= 32 || k < 1) ? 0xffffffffu : ((1u << k) - 1u); }
  if (const char* e = getenv("SB_CELL_CACHE")) h->cellCache = atoi(e) < 0 ? 0x7fffffff : atoi(e);
  cudaEventCreate(&h->evP1); cudaEventCreate(&h->evP2);
  cudaEventCreate(&h->evK0); cudaEventCreate(&h->evK1); cudaEventCreate(&h->evT0); cudaEventCreate(&h->evT1);
  if (cudaDeviceSynchronize() != cudaSuccess) { g_globalErr = "scone_b200: device initialisation failed"; delete h; return -1; }   // null-stream uploads above
  *out = h;
  return 0;
}

void sb_destroy(sb_engine* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (int i = 0; i < 3; ++i) freeBank(h->bank[i]);
  cudaFree(h->dNsites); cudaFree(h->dOffsets); cudaFree(h->dTile); cudaFree(h->dFlag); cudaFree(h->dFlagOff);
  cudaFree(h->dHProd); cudaFree(h->dHAbs); cudaFree(h->dHLeak); cudaFree(h->dHScat); cudaFree(h->dRn);
  cudaFree(h->dCand); cudaFree(h->dHist); cudaFree(h->dPartial); cudaFree(h->dHot); cudaFree(h->dSeedTab); cudaFree(h->dKsum); cudaFree(h->dNd); cudaFree(h->dRnGlobal); cudaFree(h->dRankCounts); cudaFreeHost(h->hRankCounts); cudaFree(h->dCd); cudaFree(h->dNcur); cudaFreeHost(h->hCd); cudaFree(h->dBlob);
  for (int ph = 0; ph < 2; ++ph) { cudaFree(h->dBins[ph]); cudaFree(h->dLast[ph]); cudaFree(h->dCsum[ph]); cudaFree(h->dCsum2[ph]); }
  cudaFree(h->dCePerm); cudaFree(h->dCeHist); cudaFree(h->dCeCursor); cudaFree(h->dCeTile); cudaFree(h->dCeNbin);
  cudaFree(h->dStage); sbce::ceFree(h->ce); cudaFree(h->dCeErr); cudaFree(h->dCeSlots); cudaFree(h->dStkD); cudaFree(h->dStkG); cudaFree(h->dFileSrc);
  if (h->ceStreamIn) { cudaStreamDestroy(h->ceStreamIn); cudaStreamDestroy(h->ceStreamOut); for (int i = 0; i < CE_PIPE; ++i) { cudaEventDestroy(h->ceEvIn[i]); cudaEventDestroy(h->ceEvK[i]); } }
  for (int r = 0; r < PEER_MAX; ++r) if (h->peerOpened[r]) cudaIpcCloseMemHandle(h->peerOpened[r]);
  cudaFree(h->peerRegion); cudaFree(h->dPlan); cudaFreeHost(h->hPlan); cudaFree(h->dKsumTot); cudaFree(h->dKsumRed);
  for (void* p : h->ceAllocs) cudaFree(p);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char* sb_last_error(sb_engine* h) { return h ? h->err.c_str() : g_globalErr.c_str(); }
int64_t sb_launch_count(sb_engine* h) { return h->launches; }

int sb_load_geometry(sb_engine* h, const sb_geom_flat* g) {
  if (g->n_uni < 1 || g->n_graph < 1 || g->root_idx < 1 || g->root_idx > g->n_uni) { h->err = "sb_load_geometry: invalid sizes"; return -1; }
  h->g = *g;
  h->gi_surfType.assign(g->surf_type, g->surf_type + g->n_surf);
  h->gd_surfPar.assign(g->surf_par, g->surf_par + (size_t)g->n_surf * SB_SURF_NPAR);
  h->gi_cellOff.assign(1, 0);
  if (g->n_cell > 0) { h->gi_cellOff.assign(g->cell_off, g->cell_off + g->n_cell + 1); h->gi_cellSurf.assign(g->cell_surf, g->cell_surf + g->cell_off[g->n_cell]); }
  else h->gi_cellSurf.clear();
  h->gi_uniType.assign(g->uni_type, g->uni_type + g->n_uni);
  h->gi_uniIpar.assign(g->uni_ipar, g->uni_ipar + (size_t)g->n_uni * SB_UNI_NIPAR);
  h->gd_uniDpar.assign(g->uni_dpar, g->uni_dpar + (size_t)g->n_uni * SB_UNI_NDPAR);
  h->gd_auxD.assign(g->aux_d, g->aux_d + g->n_aux_d); h->gi_auxI.assign(g->aux_i, g->aux_i + g->n_aux_i);
  if (h->gd_auxD.empty()) h->gd_auxD.push_back(0.0);
  if (h->gi_auxI.empty()) h->gi_auxI.push_back(0);
  if (h->gi_cellSurf.empty()) h->gi_cellSurf.push_back(0);
  h->gi_gidx.assign(g->graph_idx, g->graph_idx + g->n_graph); h->gi_gid.assign(g->graph_id, g->graph_id + g->n_graph);
  if (g->border_idx < 1 || g->border_idx > g->n_surf) { h->err = "sb_load_geometry: invalid border surface index"; return -1; }
  // geometry % bounds() (geometryStd_class.f90:188-206) of the border surface, for the fission source
  {
    int t = g->surf_type[g->border_idx - 1]; const double* p = g->surf_par + (size_t)(g->border_idx - 1) * SB_SURF_NPAR;
    double lo[3] = {-INF, -INF, -INF}, hi[3] = {INF, INF, INF};
    if (t == SB_SURF_BOX) for (int a = 0; a < 3; ++a) { lo[a] = p[a] - p[3 + a]; hi[a] = p[a] + p[3 + a]; }
    else if (t >= SB_SURF_XTCYL) { int ax = t - SB_SURF_XTCYL; for (int a = 0; a < 3; ++a) { double hw = (a == ax) ? p[5] : p[3]; lo[a] = p[a] - hw; hi[a] = p[a] + hw; } }
    else if (t >= SB_SURF_XSQCYL) { int ax = t - SB_SURF_XSQCYL; for (int a = 0; a < 3; ++a) if (a != ax) { lo[a] = p[a] - p[3 + a]; hi[a] = p[a] + p[3 + a]; } }
    else if (t == SB_SURF_SPHERE) for (int a = 0; a < 3; ++a) { lo[a] = p[a] - p[3]; hi[a] = p[a] + p[3]; }
    else if (t >= SB_SURF_XCYL && t <= SB_SURF_ZCYL) { int ax = t - SB_SURF_XCYL; for (int a = 0; a < 3; ++a) if (a != ax) { lo[a] = p[a] - p[3]; hi[a] = p[a] + p[3]; } }
    else if (t >= SB_SURF_XPLANE && t <= SB_SURF_ZPLANE) { lo[t - SB_SURF_XPLANE] = p[0]; hi[t - SB_SURF_XPLANE] = p[0]; }
    for (int a = 0; a < 3; ++a) { if (lo[a] <= -INF && hi[a] >= INF) { lo[a] = 0.0; hi[a] = 0.0; } h->bounds[a] = lo[a]; h->bounds[a + 3] = hi[a]; }
  }
  h->haveGeom = true; h->blobDirty = true;
  return 0;
}

int sb_load_mg_data(sb_engine* h, const sb_mg_flat* d) {
  if (d->n_mat < 1 || d->n_g < 1) { h->err = "sb_load_mg_data: invalid sizes"; return -1; }
  size_t nm = d->n_mat, ng = d->n_g;
  h->nMat = d->n_mat; h->nG = d->n_g; h->isP1 = d->P1 ? 1 : 0; h->collisionXS = d->collision_xs;
  h->xs.assign(d->data, d->data + nm * ng * 6);
  h->P0.assign(d->P0, d->P0 + nm * ng * ng); h->prod.assign(d->prod, d->prod + nm * ng * ng);
  if (d->P1) h->P1.assign(d->P1, d->P1 + nm * ng * ng); else h->P1.clear();
  h->chi.assign(d->chi, d->chi + nm * ng);
  h->fissile.assign(d->fissile, d->fissile + nm);
  h->majorant.assign(d->majorant, d->majorant + ng);
  h->haveData = true; h->blobDirty = true;
  return 0;
}

int sb_define_tallies(sb_engine* h, int phase, const sb_clerk* clerks, int n, int normClerk, double normVal) {
  if (phase < 0 || phase > 1) { h->err = "sb_define_tallies: phase must be 0 or 1"; return -1; }
  h->clerks[phase].clear(); h->mapBounds[phase].clear(); h->mapMat[phase].clear(); h->shannon[phase].clear();
  h->mapBounds[phase].resize((size_t)std::max(1, n) * SB_MAX_MAPS); h->mapMat[phase].resize((size_t)std::max(1, n) * SB_MAX_MAPS);
  int memLoc = 1;
  h->normAddr[phase] = 0; h->normVal[phase] = normVal;
  h->userKeff[phase] = UserKeff{};
  for (int c = 0; c < n; ++c) {
    const sb_clerk& s = clerks[c];
    if (s.kind == SB_CLERK_KEFF_ANALOG || s.kind == SB_CLERK_KEFF_IMPLICIT) {          // a clerk record without maps and responses keeps the order
      UserKeff& uk = h->userKeff[phase];
      if (uk.n >= 4) { h->err = "sb_define_tallies: at most 4 k-eff clerks per tally"; return -1; }
      uk.kind[uk.n] = s.kind; uk.addr[uk.n] = memLoc; uk.n++;
      DClerk d; memset(&d, 0, sizeof(d)); d.addr = memLoc; d.handleVirtual = 1;
      if (normClerk == c + 1) h->normAddr[phase] = memLoc;
      memLoc += (s.kind == SB_CLERK_KEFF_ANALOG) ? 3 : 5;
      h->clerks[phase].push_back(d);
      continue;
    }
    if (s.kind != SB_CLERK_COLLISION && s.kind != SB_CLERK_TRACK && s.kind != SB_CLERK_SHANNON) { h->err = "sb_define_tallies: unknown clerk kind"; return -1; }
    const bool shannon = s.kind == SB_CLERK_SHANNON;
    if (s.n_maps < 0 || s.n_maps > SB_MAX_MAPS || (!shannon && (s.n_resp < 1 || s.n_resp > SB_MAX_RESP))) { h->err = "sb_define_tallies: invalid clerk"; return -1; }
    if (shannon && (s.n_maps < 1 || s.cycles < 0)) { h->err = "sb_define_tallies: shannonEntropyClerk needs a map and a number of cycles"; return -1; }
    DClerk d; memset(&d, 0, sizeof(d));
    d.addr = memLoc; d.nMaps = s.n_maps; d.nResp = shannon ? 0 : s.n_resp; d.handleVirtual = s.handle_virtual; d.kind = s.kind; d.padk = shannon ? s.cycles : 0;
    int mul = 1;
    for (int m = 0; m < s.n_maps; ++m) {
      const sb_map1d& mp = s.maps[m];
      d.mapType[m] = mp.type; d.mapAxis[m] = mp.axis; d.mapGrid[m] = mp.grid; d.mapN[m] = mp.n_bins; d.mapMul[m] = mul;
      d.mapFirst[m] = mp.first; d.mapStep[m] = mp.step; d.mapInv[m] = (mp.step != 0.0) ? 1.0 / mp.step : 0.0; d.mapDef[m] = mp.default_bin;
      size_t key = (size_t)c * SB_MAX_MAPS + m;
      if (mp.type == SB_MAP_MATERIAL) {
        if (!mp.mat_bin) { h->err = "sb_define_tallies: materialMap without mat_bin"; return -1; }
        h->mapMat[phase][key].assign(mp.mat_bin, mp.mat_bin + h->nMat);
        d.mapGrid[m] = h->nMat;
      } else if (mp.grid == SB_GRID_UNSTRUCT) {
        if (!mp.bounds) { h->err = "sb_define_tallies: unstructured grid without bounds"; return -1; }
        h->mapBounds[phase][key].assign(mp.bounds, mp.bounds + mp.n_bins + 1);
      }
      mul *= mp.n_bins;
    }
    for (int i = 0; i < d.nResp; ++i) d.respMT[i] = s.resp_mt[i];
    if (normClerk == c + 1) h->normAddr[phase] = memLoc;
    if (shannon) { h->shannon[phase].push_back(sb_engine::ShannonRec{(int)h->clerks[phase].size(), memLoc, mul, s.cycles, 0}); memLoc += mul + 1 + s.cycles; }
    else memLoc += s.n_resp * mul;
    h->clerks[phase].push_back(d);
  }
  h->nBins[phase] = memLoc - 1;
  h->blobDirty = true;
  return 0;
}

int sb_set_options(sb_engine* h, const sb_options* o) {
  if (o->tracking != SB_TRACK_DT && o->tracking != SB_TRACK_ST && o->tracking != SB_TRACK_HT) { h->err = "sb_set_options: unknown tracking"; return -1; }
  { int keep = h->opt.blocks_per_sm; h->opt = *o; if (h->opt.blocks_per_sm <= 0) h->opt.blocks_per_sm = keep; }
  if (o->max_pop > 0) return ensureCapacity(h, o->max_pop);
  return 0;
}

int sb_bank_size(sb_engine* h) { return h->nCur; }
int sb_bank_brood(sb_engine* h, int cap, int32_t* brood) {
  CUDA_OK(cudaSetDevice(h->device));
  if (h->nCur > cap) { h->err = "sb_bank_brood: buffer too small"; return -1; }
  if (h->nCur == 0) return 0;
  if (!h->broodValid) { for (int i = 0; i < h->nCur; ++i) brood[i] = 0; return 0; }
  CUDA_OK(cudaMemcpyAsync(brood, h->bank[h->cur].brood, sizeof(int) * h->nCur, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < h->nCur; ++i) brood[i] += 1;      // the device numbers the histories of a cycle from 0, SCONE from 1
  return 0;
}

static int ensureStage(sb_engine* h, size_t bytes) {
  if (bytes <= h->stageBytes) return 0;
  cudaFree(h->dStage);
  CUDA_OK(cudaMalloc(&h->dStage, bytes));
  h->stageBytes = bytes;
  return 0;
}

int sb_bank_upload(sb_engine* h, int n, const double* r, const double* dir, const double* w, const int32_t* G) {
  if (ensureCapacity(h, std::max(n, h->opt.max_pop))) return -1;
  CUDA_OK(cudaSetDevice(h->device));
  if (ensureStage(h, sizeof(double) * 6 * (size_t)h->cap)) return -1;
  Bank& b = h->bank[h->cur];
  cudaStream_t st = h->stream;
  CUDA_OK(cudaMemcpyAsync(h->dStage, r, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(h->dStage + 3 * (size_t)h->cap, dir, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(b.w, w, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(b.G, G, sizeof(int) * n, cudaMemcpyHostToDevice, st));
  k_bank_unpack<<<gridFor(h, n, 256), 256, 0, st>>>(h->dStage, h->dStage + 3 * (size_t)h->cap, b, n);
  h->launches++;
  CUDA_OK(cudaStreamSynchronize(st));
  h->nCur = n; h->broodValid = false;
  return 0;
}

int sb_bank_download(sb_engine* h, int cap, int* n, double* r, double* dir, double* w, int32_t* G) {
  CUDA_OK(cudaSetDevice(h->device));
  int m = h->nCur;
  *n = m;
  if (m > cap) { h->err = "sb_bank_download: buffer too small"; return -1; }
  if (m == 0) return 0;
  if (ensureStage(h, sizeof(double) * 6 * (size_t)h->cap)) return -1;
  Bank& b = h->bank[h->cur];
  cudaStream_t st = h->stream;
  k_bank_pack<<<gridFor(h, m, 256), 256, 0, st>>>(b, h->dStage, h->dStage + 3 * (size_t)h->cap, m);
  h->launches++;
  CUDA_OK(cudaMemcpyAsync(r, h->dStage, sizeof(double) * 3 * m, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaMemcpyAsync(dir, h->dStage + 3 * (size_t)h->cap, sizeof(double) * 3 * m, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaMemcpyAsync(w, b.w, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaMemcpyAsync(G, b.G, sizeof(int) * m, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

static int checkDeviceError(sb_engine* h, int code) {
  if (code == 0) return 0;
  const char* msg = "unknown device error";
  switch (code) {
    case SB_ERR_BANK_OVERFLOW: msg = "Run out of space for particles (particleDungeon detain): fission bank overflow"; break;
    case SB_ERR_UNDEF_MAT: msg = "Particle is in undefined material"; break;
    case SB_ERR_OVERLAP_MAT: msg = "Particle is in overlapping cells"; break;
    case SB_ERR_SAMPLING: msg = "Sampling failed (scatter XS / chi normalisation or random number above 1)"; break;
    case SB_ERR_NEST: msg = "Failed to find material cell (nesting exceeded)"; break;
    case SB_ERR_MAT_SOURCE_VOID: msg = "materialSource: Nuclear data did not return neutron material (a sampled point lies in a void region)."; break;
    case SB_ERR_MAT_SOURCE: msg = "materialSource: Infinite loop in sampling source. Please check that defined volume contains source material."; break;
    case SB_ERR_PEER_TIMEOUT: msg = "peer exchange: a rank of the node did not post its cycle data in time"; break;
    case SB_ERR_BALANCE: msg = "loadBalancing: nearest-neighbour exchange cannot restore the shares of this distribution"; break;
    case SB_ERR_FILE_SOURCE: msg = "fileSource: neutron sampled from file source is outside of geometry or in undefined region"; break;
    case SB_ERR_SOURCE: msg = "fissionSource: failed to find a fissile material in 10000 attempts"; break;
    case SB_ERR_NORM: msg = "Normalisation failed!"; break;
    case SB_ERR_CE_ENERGY: msg = "Failed to find energy in the nuclide energy grids (particle energy outside the bounds of the CE data)"; break;
    case SB_ERR_CE_DATA: msg = "Continuous-energy reaction data: a table search or a rejection loop failed"; break;
  }
  h->err = msg;
  return -1;
}

int sb_source_generate(sb_engine* h, int n, uint64_t rng_state, int history_offset) {
  if (buildBlob(h)) return -1;
  if (ensureCapacity(h, std::max(n, h->opt.max_pop))) return -1;
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaMemsetAsync(h->dCd, 0, sizeof(CycleDev), h->stream));
  const double* b = h->bounds;
  if (h->ceMode) sbc::k_source_ce<<<gridFor(h, n, 128), 128, 0, h->stream>>>(h->M, h->dBlob, h->ceModel, h->bank[h->cur], n, rng_state, history_offset, b[0], b[1], b[2], b[3], b[4], b[5], h->dCd);
  else k_source<<<gridFor(h, n, 128), 128, 0, h->stream>>>(h->M, h->dBlob, h->bank[h->cur], n, rng_state, history_offset, b[0], b[1], b[2], b[3], b[4], b[5], h->dCd);
  h->launches++;
  CUDA_OK(cudaMemcpyAsync(h->hCd, h->dCd, sizeof(CycleDev), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaGetLastError());
  h->nCur = n; h->broodValid = false;
  return checkDeviceError(h, h->hCd->error);
}

int sb_set_fixed_source(sb_engine* h, int on, int buffer_size) {
  if (on && buffer_size < 1) { h->err = "sb_set_fixed_source: buffer size must be +ve"; return -1; }
  h->fixedSource = on != 0;
  if (on) h->stkCap = buffer_size;
  return 0;
}
int sb_source_point(sb_engine* h, int n, uint64_t rng_state, int history_offset, const sb_point_source* s) {
  if (!s || n < 1) { h->err = "sb_source_point: invalid arguments"; return -1; }
  if (buildBlob(h)) return -1;
  if (ensureCapacity(h, std::max(n, h->opt.max_pop))) return -1;
  CUDA_OK(cudaSetDevice(h->device));
  if (h->ceMode == (s->is_mg != 0)) { h->err = "sb_source_point: the source energy type (E / G) does not match the loaded nuclear data"; return -1; }
  const double* dProb = nullptr;
  if (s->is_mg && s->n_prob > 0) {
    if (s->n_prob != h->nG) { h->err = "Source energy group distribution must have as many entries as there are groups"; return -1; }
    if (ensureStage(h, sizeof(double) * 6 * (size_t)h->cap)) return -1;
    CUDA_OK(cudaMemcpyAsync(h->dStage, s->prob_g, sizeof(double) * s->n_prob, cudaMemcpyHostToDevice, h->stream));
    dProb = h->dStage;
  } else if (s->is_mg && (s->G < 1 || s->G > h->nG)) { h->err = "sb_source_point: source group outside the group structure"; return -1; }
  k_source_point<<<gridFor(h, n, 128), 128, 0, h->stream>>>(h->bank[h->cur], n, rng_state, history_offset, s->r[0], s->r[1], s->r[2], s->dir[0], s->dir[1], s->dir[2],
                                                             s->isotropic, s->is_mg, s->E, s->G, s->is_mg ? s->n_prob : 0, dProb);
  h->launches++;
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaGetLastError());
  h->nCur = n; h->broodValid = false;
  return 0;
}
int sb_geometry_bounds(sb_engine* h, double* b) { for (int i = 0; i < 6; ++i) b[i] = h->bounds[i]; return 0; }
int sb_source_material(sb_engine* h, int n, uint64_t rng_state, int history_offset, const sb_material_source* s) {
  if (!s || n < 1) { h->err = "sb_source_material: invalid arguments"; return -1; }
  if (h->ceMode == (s->is_mg != 0)) { h->err = "sb_source_material: the source data type (ce / mg) does not match the loaded nuclear data"; return -1; }
  if (s->mat_idx < 1 || s->mat_idx > h->nMat) { h->err = "sb_source_material: source material was not found in the material definitions"; return -1; }
  if (s->is_mg && (s->G < 1 || s->G > h->nG)) { h->err = "sb_source_material: source group outside the group structure"; return -1; }
  if (buildBlob(h)) return -1;
  if (ensureCapacity(h, std::max(n, h->opt.max_pop))) return -1;
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaMemsetAsync(h->dCd, 0, sizeof(CycleDev), h->stream));
  k_source_material<<<gridFor(h, n, 128), 128, 0, h->stream>>>(h->M, h->dBlob, h->bank[h->cur], n, rng_state, history_offset, *s, h->dCd);
  h->launches++;
  CUDA_OK(cudaMemcpyAsync(h->hCd, h->dCd, sizeof(CycleDev), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaGetLastError());
  h->nCur = n; h->broodValid = false;
  return checkDeviceError(h, h->hCd->error);
}
int sb_set_file_source(sb_engine* h, int64_t n_rows, const double* rows, int is_mg) {
  if (n_rows < 1 || !rows) { h->err = "sb_set_file_source: the source file holds no particles"; return -1; }
  if (h->ceMode == (is_mg != 0)) { h->err = "sb_set_file_source: source data type inconsistent with nuclear database"; return -1; }
  if (is_mg) for (int64_t i = 0; i < n_rows; ++i) {
    int g = (int)rows[10 * i + 7];
    if (g < 1 || g > h->nG) { h->err = "sb_set_file_source: a source particle has a group outside the group structure"; return -1; }
  }
  CUDA_OK(cudaSetDevice(h->device));
  cudaFree(h->dFileSrc); h->dFileSrc = nullptr;
  CUDA_OK(cudaMalloc(&h->dFileSrc, sizeof(double) * 10 * (size_t)n_rows));
  CUDA_OK(cudaMemcpy(h->dFileSrc, rows, sizeof(double) * 10 * (size_t)n_rows, cudaMemcpyHostToDevice));
  CUDA_OK(cudaDeviceSynchronize());
  h->nFileSrc = n_rows; h->fileSrcMG = is_mg != 0;
  return 0;
}
int sb_source_file(sb_engine* h, int n, uint64_t rng_state, int history_offset) {
  if (n < 1) { h->err = "sb_source_file: invalid arguments"; return -1; }
  if (!h->dFileSrc) { h->err = "sb_source_file: no file source has been set"; return -1; }
  if (buildBlob(h)) return -1;
  if (ensureCapacity(h, std::max(n, h->opt.max_pop))) return -1;
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaMemsetAsync(h->dCd, 0, sizeof(CycleDev), h->stream));
  k_source_file<<<gridFor(h, n, 128), 128, 0, h->stream>>>(h->M, h->dBlob, h->bank[h->cur], n, rng_state, history_offset, h->dFileSrc, h->nFileSrc,
                                                            h->fileSrcMG ? 1 : 0, h->dCd);
  h->launches++;
  CUDA_OK(cudaMemcpyAsync(h->hCd, h->dCd, sizeof(CycleDev), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaGetLastError());
  h->nCur = n; h->broodValid = false;
  return checkDeviceError(h, h->hCd->error);
}
// lanes x buffer_size entries of { r, dir, w, E } + G
static int ensureSecStack(sb_engine* h, size_t lanes, sbt::SecStack& out) {
  out = sbt::SecStack{nullptr, nullptr, h->stkCap, 0};
  if (!h->fixedSource) return 0;
  if (lanes > h->stkLanes || h->stkCap > h->stkAllocCap) {
    cudaFree(h->dStkD); cudaFree(h->dStkG); h->dStkD = nullptr; h->dStkG = nullptr;
    size_t L = std::max(lanes, h->stkLanes); int cap = std::max(h->stkCap, h->stkAllocCap);
    CUDA_OK(cudaMalloc(&h->dStkD, sizeof(double) * 8 * (size_t)cap * L));
    CUDA_OK(cudaMalloc(&h->dStkG, sizeof(int) * (size_t)cap * L));
    h->stkLanes = L; h->stkAllocCap = cap;
  }
  out.d = h->dStkD; out.G = h->dStkG; out.cap = h->stkCap; out.on = 1;
  return 0;
}

// transport of the whole bank + brood ordering + per-rank score sums (no cycle close yet)
static int cycleTransport(sb_engine* h, uint64_t rng_state, int history_offset, double k_eff, int phase) {
  if (phase < 0 || phase > 1) { h->err = "sb_run_cycle: phase must be 0 or 1"; return -1; }
  if (buildBlob(h)) return -1;
  if (h->nCur <= 0) { h->err = "sb_run_cycle: empty bank"; return -1; }
  CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const int n = h->nCur;
  Bank& in = h->bank[h->cur]; Bank& raw = h->bank[(h->cur + 1) % 3]; Bank& sorted = h->bank[(h->cur + 2) % 3];

  // reset the running-cycle record (the cumulative k sums stay)
  if (!h->dLoneCtl) {                                          // one queue slot per lane that can be resident (12 warps on every SM)
    h->loneCap = h->numSM * 512;
    CUDA_OK(cudaMalloc(&h->dLoneQ, sizeof(sbh::LoneRec) * (size_t)h->loneCap)); CUDA_OK(cudaMalloc(&h->dLoneCtl, 4 * sizeof(int)));
    CUDA_OK(cudaMalloc(&h->dLoneReady, sizeof(int) * (size_t)h->loneCap)); CUDA_OK(cudaMemsetAsync(h->dLoneReady, 0, sizeof(int) * (size_t)h->loneCap, st));
  }
  CUDA_OK(cudaMemsetAsync(h->dLoneCtl, 0, 4 * sizeof(int), st));
  k_cycle_begin<<<1, 1, 0, st>>>(h->dCd, h->dNcur, n);
  h->launches++;

  HistArgs a{};
  a.L = h->hot; a.L.oClerk[0] = h->hot.oClerk[phase]; a.L.nClerk[0] = h->hot.nClerk[phase]; a.L.oScoreMask[0] = h->hot.oScoreMask[phase];
  a.hot = h->dHot; a.blob = h->dBlob; a.seedTab = h->dSeedTab;
  a.n = n; a.in = in; a.out = raw; a.cap = h->cap;
  a.nsites = h->dNsites; a.hProd = h->dHProd; a.hAbs = h->dHAbs; a.hLeak = h->dHLeak; a.hScat = h->dHScat;
  a.bins = h->dBins[phase]; a.phase = phase;
  int impScores = (phase == 1) ? 1 : 0;                          // the active attachment clerk is keffImplicitClerk; a user clerk may ask for the scores in any phase
  for (int i = 0; i < h->userKeff[phase].n; ++i) if (h->userKeff[phase].kind[i] == SB_CLERK_KEFF_IMPLICIT) impScores = 1;
  a.impScores = impScores;
  a.rng0 = rng_state; a.histOffset = history_offset; a.k_eff = k_eff; a.cd = h->dCd;
#ifdef SB_PROFILE_ROUNDS
  { static long long* dProf = nullptr; if (!dProf) { cudaMalloc(&dProf, 8 * (36 * 148 * 384 + 12 * 148 * 16)); } cudaMemsetAsync(dProf, 0, 8 * (36 * 148 * 384 + 12 * 148 * 16), st); a.prof = dProf; h->dProfRounds = dProf; }
#endif
  a.refillMin = h->refillMin; a.maxSegMin = h->maxSegMin; a.loneMode = h->loneMode; a.cellCache = h->cellCache; a.laneMask = h->laneMask;
  // the last histories of every warp go on in k_lone (speculative batches; multiScatterMG and multiScatterP1MG).
  // A warp hands its histories over when it is down to T of them: as many as k_lone's warps (8 per SM) can take at once, counted over
  // the warps that will run (measured, profiles/README.md: T = 1 at 1e5 histories, 2 at 2e4, 8 at 2e3 per GPU)
  {
    const int warpsA = std::min(h->numSM * 12, (n + 31) / 32);
    int T = h->assist >= 0 ? h->assist : std::max(1, std::min(8, (int)(1.1 * 8 * h->numSM / std::max(1, warpsA))));
    a.assist = (T > 0 && h->loneMode && h->useSmem) ? std::min(T, 32) : 0;
  }
  a.loneQ = h->dLoneQ; a.loneCount = h->dLoneCtl; a.loneNext = h->dLoneCtl + 1; a.loneCap = h->loneCap;
  a.loneReady = h->dLoneReady; a.loneDone = h->dLoneCtl + 2; a.loneTag = ++h->loneTag;
  // one CTA of 12 warps per SM (168 registers per thread, no spills) unless told otherwise: at the populations of an
  // eigenvalue cycle the kernel's time is the chain of its longest history, not the number of resident warps
  int threads = 384;
  const int bps = h->opt.blocks_per_sm > 0 ? h->opt.blocks_per_sm : 1;
  if (const char* e = getenv("SB_HIST_THREADS")) threads = atoi(e);
  if (bps != 1 || !h->useSmem || (threads != 384 && threads != 448 && threads != 512)) threads = 256;
  int blocks = h->numSM * bps;
  int needBlocks = (n + threads - 1) / threads;
  if (needBlocks < blocks) blocks = needBlocks;
  a.loneWarps = blocks * (threads / 32);
  if (h->profiling) CUDA_OK(cudaEventRecord(h->evK0, st));
  bool useTrack = h->opt.tracking != SB_TRACK_DT || h->fixedSource;      // the DT-only kernel has no secondary buffer
  // (trackClerks score along surface-tracking segments only: under delta tracking the reference makes no path reports either)
  if (!h->ceMode && !h->fixedSource && h->opt.tracking == SB_TRACK_HT && !getenv("SB_FORCE_TRACK_KERNEL")) {
    // transportOperatorHT picks delta tracking when Sigma_t / Sigma_maj > 1 - cutoff (transportOperatorHT_class.f90:63-78). If that
    // holds for every (material, group) of the model -- and nothing is void -- the selector is a constant and the flights
    // are exactly deltaTracking's: run the delta-tracking kernel.
    bool always = true;
    for (int i = 0; i < h->g.n_graph; ++i) if (h->gi_gidx[i] == SB_VOID_MAT) always = false;
    for (int m = 0; m < h->nMat && always; ++m)
      for (int g = 0; g < h->nG; ++g) {
        double majorant_inv = 1.0 / std::fmax(h->majorant[g] + 0.0, h->collisionXS);
        double ratio = (h->xs[((size_t)m * h->nG + g) * 6] + 0.0) * majorant_inv;
        if (!(ratio > (1.0 - h->opt.ht_cutoff))) { always = false; break; }
      }
    if (always) useTrack = false;
  }
  if (h->ceMode) {                                            // continuous energy: sb_cehist.cuh
    sbc::CeArgs t{};
    t.M = h->M; t.blob = h->dBlob; t.ce = h->ceModel; t.seedTab = h->dSeedTab;
    t.n = n; t.in = in; t.out = raw; t.cap = h->cap;
    t.nsites = h->dNsites; t.hProd = h->dHProd; t.hAbs = h->dHAbs; t.hLeak = h->dHLeak; t.hScat = h->dHScat;
    t.bins = h->dBins[phase]; t.phase = phase; t.rng0 = rng_state; t.histOffset = history_offset; t.k_eff = k_eff; t.cd = h->dCd;
    t.tracking = h->opt.tracking; t.htCutoff = h->opt.ht_cutoff; t.stCache = h->opt.st_cache; t.impScores = impScores;
    t.needMacro = 0;
    for (const DClerk& k : h->clerks[phase]) { for (int i = 0; i < k.nResp; ++i) if (k.respMT[i] != 0) t.needMacro = 1; if (k.kind == SB_CLERK_TRACK) t.nTrackClerks++; }
    const char* cfg = getenv("SB_CE_KERNEL");                  // experiment switch: "async" = 128-thread CTAs without phase barriers
    if (h->fixedSource) {                                      // fixed source: private secondary buffers, lockstep kernel
      cfg = nullptr;
      const int threadsF = n < 400000 ? 512 : 1024;
      if (ensureSecStack(h, (size_t)std::min(h->numSM, (n + threadsF - 1) / threadsF) * threadsF, t.stk)) return -1;
    }
    if (t.nTrackClerks) cfg = nullptr;                             // path-length scores are made by the lane-resident kernel
    if (cfg && (!strcmp(cfg, "events") || !strcmp(cfg, "events1024"))) {   // event queues over slots in global memory (sb_ceevent.cuh)
      const int slotsPerCta = !strcmp(cfg, "events") ? 512 : 1024;
      const size_t need = (size_t)h->numSM * slotsPerCta;
      if (need > h->ceSlotCount) { cudaFree(h->dCeSlots); h->dCeSlots = nullptr; CUDA_OK(cudaMalloc(&h->dCeSlots, sbc::CeSlots::bytes(need))); h->ceSlotCount = need; }
      sbc::CeEventArgs ea{t, sbc::CeSlots{}};
      ea.S.carve(h->dCeSlots, h->ceSlotCount);
      if (slotsPerCta == 512) sbc::k_events_ce<512, 512><<<h->numSM, 512, 0, st>>>(ea);
      else sbc::k_events_ce<512, 1024><<<h->numSM, 512, 0, st>>>(ea);
    } else
    if (cfg && !strcmp(cfg, "async")) sbc::k_histories_ce<128, 4, false><<<std::min(h->numSM * 4, (n + 127) / 128), 128, 0, st>>>(t);
    else if (cfg && !strcmp(cfg, "sync256")) sbc::k_histories_ce<256, 2, true><<<std::min(h->numSM * 2, (n + 255) / 256), 256, 0, st>>>(t);
    else if (cfg && !strcmp(cfg, "sync768")) sbc::k_histories_ce<768, 1, true><<<std::min(h->numSM, (n + 767) / 768), 768, 0, st>>>(t);
    else if (cfg && !strcmp(cfg, "sync256x1")) sbc::k_histories_ce<256, 2, true><<<std::min(h->numSM, (n + 255) / 256), 256, 0, st>>>(t);
    else if (cfg && !strcmp(cfg, "sync128x1")) sbc::k_histories_ce<128, 4, false><<<std::min(h->numSM, (n + 127) / 128), 128, 0, st>>>(t);
    else if (cfg && !strcmp(cfg, "sync64x1")) sbc::k_histories_ce<128, 4, false><<<std::min(h->numSM, (n + 63) / 64), 64, 0, st>>>(t);
    else if ((cfg && !strcmp(cfg, "sync512")) || (!cfg && n < 400000)) sbc::k_histories_ce<512, 1, true><<<std::min(h->numSM, (n + 511) / 512), 512, 0, st>>>(t);
    else sbc::k_histories_ce<1024, 1, true><<<std::min(h->numSM, (n + 1023) / 1024), 1024, 0, st>>>(t);   // large banks: 32 warps per SM at 64 registers hide more latency (120 -> 108 ms at 1e6)
  } else
  if (useTrack) {                                             // surface / hybrid tracking: coordList-carrying kernel
    sbt::TrackArgs t{};
    t.M = h->M; t.blob = h->dBlob; t.useSmem = h->trackSmem; t.seedTab = h->dSeedTab;
    t.n = n; t.in = in; t.out = raw; t.cap = h->cap;
    t.nsites = h->dNsites; t.hProd = h->dHProd; t.hAbs = h->dHAbs; t.hLeak = h->dHLeak; t.hScat = h->dHScat;
    t.bins = h->dBins[phase]; t.phase = phase; t.rng0 = rng_state; t.histOffset = history_offset; t.k_eff = k_eff; t.cd = h->dCd;
    t.tracking = h->opt.tracking; t.htCutoff = h->opt.ht_cutoff; t.stCache = h->opt.st_cache; t.impScores = impScores;
    for (const DClerk& k : h->clerks[phase]) if (k.kind == SB_CLERK_TRACK) t.nTrackClerks++;
    const char* cfg = getenv("SB_TRACK_KERNEL");               // experiment switch: "async" = 128-thread CTAs without phase barriers
    if (h->fixedSource) { cfg = nullptr; if (ensureSecStack(h, (size_t)std::min(h->numSM, (n + 511) / 512) * 512, t.stk)) return -1; }
    if (cfg && !strcmp(cfg, "async")) sbt::k_histories_track<128, 4, false><<<std::min(h->numSM * 4, (n + 127) / 128), 128, h->trackSmem ? h->M.blobBytes : 0, st>>>(t);
    else sbt::k_histories_track<512, 1, true><<<std::min(h->numSM, (n + 511) / 512), 512, h->trackSmem ? h->M.blobBytes : 0, st>>>(t);
  } else
  {
    const int hotB = h->useSmem ? h->hot.bytes : 0;
    if (h->useSmem && bps == 1 && threads == 384 && a.assist > 0 && !h->hot.isP1) sbh::k_histories<true, 1, 384, false, false><<<blocks, 384, hotB + sbh::histScratchBytes(384), st>>>(a);
    else if (h->useSmem && bps == 1 && threads == 384 && a.assist > 0) sbh::k_histories<true, 1, 384, false, true><<<blocks, 384, hotB + sbh::histScratchBytes(384), st>>>(a);
    else if (h->useSmem && bps == 1 && threads == 384) sbh::k_histories<true, 1, 384><<<blocks, 384, hotB + sbh::histScratchBytes(384), st>>>(a);
    else if (h->useSmem && bps == 1 && threads == 448) sbh::k_histories<true, 1, 448><<<blocks, 448, hotB + sbh::histScratchBytes(448), st>>>(a);
    else if (h->useSmem && bps == 1 && threads == 512) sbh::k_histories<true, 1, 512><<<blocks, 512, hotB + sbh::histScratchBytes(512), st>>>(a);
    else if (h->useSmem && bps == 1) sbh::k_histories<true, 1><<<blocks, 256, hotB + sbh::histScratchBytes(256), st>>>(a);
    else if (h->useSmem && bps >= 3) sbh::k_histories<true, 3><<<blocks, 256, hotB + sbh::histScratchBytes(256), st>>>(a);
    else if (h->useSmem) sbh::k_histories<true, 2><<<blocks, 256, hotB + sbh::histScratchBytes(256), st>>>(a);
    else sbh::k_histories<false, 2><<<blocks, 256, sbh::histScratchBytes(256), st>>>(a);
    if (a.assist > 0) {                                            // the last histories of every warp, one warp each
      if (h->hot.isP1) pdlLaunchSmem(sbh::k_lone<128, true>, h->numSM * 2, 128, (size_t)(hotB + sbh::loneScratchBytes(128)), st, a);
      else pdlLaunchSmem(sbh::k_lone<128, false>, h->numSM * 2, 128, (size_t)(hotB + sbh::loneScratchBytes(128)), st, a);
      h->launches++;
    }
    pdlLaunch(k_finish_sites, gridFor(h, n, 128), 128, st, h->M, h->dBlob, raw, h->dCd, h->cap);     // the sites' directions and groups
    h->launches++;
  }
  if (h->profiling) CUDA_OK(cudaEventRecord(h->evK1, st));
  h->launches++;

  // brood offsets = exclusive scan of per-history site counts ; stable brood order
  int tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  pdlLaunch(k_scan_reduce, tiles, SCAN_BLOCK, st, h->dNsites, h->dNcur, h->dTile);
  pdlLaunch(k_scan_tiles, 1, 1024, st, h->dTile, h->dNcur, nullptr);
  pdlLaunch(k_scan_apply, tiles, SCAN_BLOCK, st, h->dNsites, h->dNcur, h->dTile, h->dOffsets);
  pdlLaunch(k_sort_sites, gridFor(h, 2LL * n, 256), 256, st, raw, sorted, h->dOffsets, h->dCd, h->cap);
  // deterministic reductions of this rank's scores
  pdlLaunch(k_reduce_hist, RED_BLOCKS, RED_THREADS, st, n, h->dHProd, h->dHAbs, h->dHLeak, h->dHScat, in.w, sorted.w, h->dCd, h->cap, h->dPartial);
  pdlLaunch(k_sum_partials, 1, 32, st, h->dPartial, h->dKsum, h->dCd, h->cap);
  h->launches += 6;
  h->phaseOpen = phase;
  return 0;
}

// cycle close from (possibly rank-reduced) score sums: k estimators, normalisation, scoreMemory%closeCycle
static int cycleCloseEnqueue(sb_engine* h, const double* dKsum) {
  const int phase = h->phaseOpen;
  if (phase < 0) { h->err = "sb_cycle_end: no cycle is open"; return -1; }
  CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  k_close_cycle_head<<<1, 32, 0, st>>>(dKsum, h->dKsum, h->dCd, phase, h->kNormNext, h->dBins[phase], h->normAddr[phase], h->normVal[phase],
                                       h->userKeff[phase], h->dCsum[phase], h->dCsum2[phase]);
  for (auto& sr : h->shannon[phase]) {                       // reportCycleEnd + closeCycle of the entropy clerks, on the un-normalised bank
    sr.cycle += 1;
    if (sr.cycle > sr.maxCycles) continue;
    Bank& sorted = h->bank[(h->cur + 2) % 3];
    pdlLaunch(k_shannon_score, gridFor(h, h->cap / 4 + 1, 256), 256, st, h->M, h->dBlob, phase, sr.clerk, sorted, h->dCd, h->cap, h->ceMode ? 1 : 0, h->dBins[phase]);
    pdlLaunch(k_shannon_close, 1, 256, st, sr.addr, sr.nBins, sr.cycle, h->dBins[phase], h->dCsum[phase], h->dCsum2[phase]);
    h->launches += 2;
  }
  int nb = std::max(1, h->nBins[phase]);
  pdlLaunch(k_close_cycle_bins, gridFor(h, nb, 256), 256, st, h->dBins[phase], h->dLast[phase], h->dCsum[phase], h->dCsum2[phase], h->nBins[phase], h->dCd);
  h->launches += 2;
  h->batchN[phase] += 1;
  return 0;
}
// read the cycle record back (one synchronisation) and fill the result
static int cycleFinish(sb_engine* h, sb_cycle_result* res) {
  cudaStream_t st = h->stream;
  CUDA_OK(cudaMemcpyAsync(h->hCd, h->dCd, sizeof(CycleDev), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  CUDA_OK(cudaGetLastError());
  const CycleDev& c = *h->hCd;
  if (h->profiling) {
    float ms = 0.f; CUDA_OK(cudaEventElapsedTime(&ms, h->evK0, h->evK1));
    h->msHistories += ms; h->nHistLaunches++; h->segProfiled += (long long)c.nSeg; h->scoreProfiled += (long long)c.nScore;
    if (h->peerStagesOpen) {
      CUDA_OK(cudaEventElapsedTime(&ms, h->evK1, h->evP1)); h->msPeerWait += ms;
      CUDA_OK(cudaEventElapsedTime(&ms, h->evP1, h->evP2)); h->msPeerTail += ms;
      h->peerStagesOpen = false;
    }
  }
  if (res) {
    res->n_start = c.nStart; res->n_sites = c.nSites; res->start_wgt = c.startWgt; res->end_wgt = c.endWgt;
    res->imp_prod = c.impProd; res->imp_abs = c.impAbs; res->scatter_prod = c.scatProd; res->ana_leak = c.anaLeak;
    res->k_analog = c.kAnalog; res->k_implicit = c.kImplicit; res->k_cum = c.kCum; res->k_cum_std = c.kCumStd;
    res->n_segments = (int64_t)c.nSeg; res->n_collisions = (int64_t)c.nColl; res->n_scores = (int64_t)c.nScore; res->error = c.error; res->max_history_segments = c.maxSeg; res->n_xs_terms = (int64_t)c.nXsTerms;
  }
  h->sortedReady = true; h->phaseOpen = -1;
  h->kCumLast = c.kCum;
  return checkDeviceError(h, c.error);
}
static int cycleClose(sb_engine* h, const double* dKsum, sb_cycle_result* res) {
  if (cycleCloseEnqueue(h, dKsum)) return -1;
  return cycleFinish(h, res);
}

int sb_run_cycle(sb_engine* h, uint64_t rng_state, int history_offset, double k_eff, int phase, sb_cycle_result* res) {
  if (cycleTransport(h, rng_state, history_offset, k_eff, phase)) return -1;
  return cycleClose(h, h->dKsum, res);
}

int sb_cycle_begin(sb_engine* h, uint64_t rng_state, int history_offset, double k_eff, int phase, double* dev_sums, int32_t* n_sites) {
  if (cycleTransport(h, rng_state, history_offset, k_eff, phase)) return -1;
  if (dev_sums) CUDA_OK(cudaMemcpyAsync(dev_sums, h->dKsum, 8 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  CUDA_OK(cudaMemcpyAsync(h->hCd, h->dCd, sizeof(CycleDev), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaGetLastError());
  if (n_sites) *n_sites = std::min(h->hCd->nSites, h->cap);
  return checkDeviceError(h, h->hCd->error);
}
int sb_cycle_end(sb_engine* h, const double* dev_sums, sb_cycle_result* res) {
  return cycleClose(h, dev_sums ? dev_sums : h->dKsum, res);
}

// normSize_Repr kernels. nSitesHost: the bank size if the host knows it, -1 if the cycle record has not been read back
// yet (fused cycle): launch sizes then come from the bank capacity, the kernels read the true size on the device.
static int resampleEnqueue(sb_engine* h, int totPop, uint64_t rng_state, int nGlobal, int offLocal, int check, int nSitesHost) {
  CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  Bank& sorted = h->bank[(h->cur + 2) % 3]; Bank& dst = h->bank[(h->cur + 1) % 3];
  const int nSites = nSitesHost >= 0 ? nSitesHost : h->cap;
  const int nG = nGlobal < 0 ? nSites : nGlobal;
  if (nG <= 0) { h->err = "sb_resample: the fission bank is empty"; return -1; }
  unsigned long long* rn = h->dRn;
  if (nGlobal >= 0) {                                  // the stream over all ranks' banks
    if ((size_t)nG > h->rnGlobalCap) { cudaFree(h->dRnGlobal); h->rnGlobalCap = (size_t)nG + nG / 4 + 1024; CUDA_OK(cudaMalloc(&h->dRnGlobal, sizeof(unsigned long long) * h->rnGlobalCap)); }
    rn = h->dRnGlobal;
  }
  pdlLaunch(k_norm_setup, 1, 1, st, h->dNd, h->dCd, h->cap, totPop, nGlobal, offLocal, check);
  int g = gridFor(h, nG, 256), gl = gridFor(h, std::max(1, nSites), 256);
  pdlLaunch(k_rn_generate, gridFor(h, (nG + RN_CHUNK - 1) / RN_CHUNK, 128), 128, st, rn, h->dNd, rng_state);
  pdlLaunch(k_zero_int, gridFor(h, SEL_BINS, 256), 256, st, h->dHist, SEL_BINS);
  pdlLaunch(k_sel_hist, g, 256, st, rn, h->dNd, h->dHist);
  pdlLaunch(k_sel_find_bin, 1, 1024, st, h->dHist, h->dCd, h->dNd);
  pdlLaunch(k_sel_collect, g, 256, st, rn, h->dCd, h->dNd, h->dCand);
  pdlLaunch(k_sel_threshold, 1, 1024, st, h->dCand, h->dCd);
  pdlLaunch(k_norm_flags, gl, 256, st, rn, h->dCd, h->dNd, h->dFlag);
  int tiles = (std::max(1, nSites) + SCAN_TILE - 1) / SCAN_TILE;
  pdlLaunch(k_scan_reduce, tiles, SCAN_BLOCK, st, h->dFlag, &h->dNd->nLocal, h->dTile);
  pdlLaunch(k_scan_tiles, 1, 1024, st, h->dTile, &h->dNd->nLocal, nullptr);
  pdlLaunch(k_scan_apply, tiles, SCAN_BLOCK, st, h->dFlag, &h->dNd->nLocal, h->dTile, h->dFlagOff);
  pdlLaunch(k_norm_scatter, gl, 256, st, sorted, dst, h->dFlag, h->dFlagOff, h->dOffsets, h->dNsites, h->dNd, h->cap, h->dCd);
  pdlLaunch(k_norm_count, 1, 1, st, h->dFlag, h->dFlagOff, h->dNd, h->dCd);
  h->launches += 13;
  return 0;
}
static int resampleFinish(sb_engine* h, int32_t* newLocal, bool readBack) {
  if (readBack) {
    CUDA_OK(cudaMemcpyAsync(h->hCd, h->dCd, sizeof(CycleDev), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    CUDA_OK(cudaGetLastError());
  }
  if (checkDeviceError(h, h->hCd->error)) return -1;
  h->cur = (h->cur + 1) % 3;
  h->nCur = h->hCd->nNew; h->broodValid = true;
  if (newLocal) *newLocal = h->nCur;
  h->kNormNext = h->kCumLast;       // self%nextCycle%k_eff = k_new (eigenPhysicsPackage_class.f90:306)
  h->sortedReady = false;
  return 0;
}
static int resampleImpl(sb_engine* h, int totPop, uint64_t rng_state, int nGlobal, int offLocal, int check, int32_t* newLocal) {
  if (!h->sortedReady) { h->err = "sb_resample: no cycle has been run"; return -1; }
  if (resampleEnqueue(h, totPop, rng_state, nGlobal, offLocal, check, std::min(h->hCd->nSites, h->cap))) return -1;
  return resampleFinish(h, newLocal, true);
}

// the whole cycle with one host synchronisation: transport, cycle close, normSize_Repr (single rank).
// rng_state_resample = pRNG state after the stride(totalPop + 1) that follows the history loop.
int sb_run_cycle_resample(sb_engine* h, uint64_t rng_state, int history_offset, double k_eff, int phase, int tot_pop, uint64_t rng_state_resample, sb_cycle_result* res) {
  if (2 * tot_pop > h->cap && h->cap > 0) { h->err = "sb_resample: target population exceeds the bank capacity"; return -1; }
  if (cycleTransport(h, rng_state, history_offset, k_eff, phase)) return -1;
  if (cycleCloseEnqueue(h, h->dKsum)) return -1;
  if (resampleEnqueue(h, tot_pop, rng_state_resample, -1, 0, 1, -1)) return -1;
  if (cycleFinish(h, res)) return -1;                      // one synchronisation; the record now holds nNew as well
  if (h->hCd->nSites <= 0) { h->err = "sb_resample: the fission bank is empty"; return -1; }
  return resampleFinish(h, nullptr, false);
}

// The whole cycle of a caller that keeps its dungeons in (page-locked) host memory, with ONE host synchronisation: upload of this
// cycle's bank, transport, cycle close, normSize_Repr, and the read-back of the normalised bank and of the cycle's BIN column are
// all enqueued on the engine's stream; the sizes of the copies do not depend on anything the device computes (single rank:
// normSize_Repr returns exactly tot_pop sites). Host arrays must stay untouched until the call returns; they need not be
// page-locked (the copies are then staged by the driver), the timing figures of bench.py use page-locked ones.
// E_or_null: continuous-energy banks carry E (G is then not read / written).
int sb_run_cycle_resample_host(sb_engine* h, int n, const double* r, const double* dir, const double* w, const int32_t* G, const double* E,
                               uint64_t rng_state, int history_offset, double k_eff, int phase, int tot_pop, uint64_t rng_state_resample,
                               int* n_out, double* r_out, double* dir_out, double* w_out, int32_t* G_out, double* E_out, double* bins_out,
                               sb_cycle_result* res) {
  if (n < 1 || !r || !dir || !w || (!G && !E)) { h->err = "sb_run_cycle_resample_host: invalid bank arguments"; return -1; }
  if (ensureCapacity(h, std::max(std::max(n, tot_pop), h->opt.max_pop))) return -1;
  if (2 * tot_pop > h->cap) { h->err = "sb_resample: target population exceeds the bank capacity"; return -1; }
  CUDA_OK(cudaSetDevice(h->device));
  if (ensureStage(h, sizeof(double) * 6 * (size_t)h->cap)) return -1;
  cudaStream_t st = h->stream;
  {
    Bank& b = h->bank[h->cur];
    CUDA_OK(cudaMemcpyAsync(h->dStage, r, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(h->dStage + 3 * (size_t)h->cap, dir, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(b.w, w, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    if (E) { CUDA_OK(cudaMemcpyAsync(b.E, E, sizeof(double) * n, cudaMemcpyHostToDevice, st)); CUDA_OK(cudaMemsetAsync(b.G, 0, sizeof(int) * n, st)); }
    else CUDA_OK(cudaMemcpyAsync(b.G, G, sizeof(int) * n, cudaMemcpyHostToDevice, st));
    k_bank_unpack<<<gridFor(h, n, 256), 256, 0, st>>>(h->dStage, h->dStage + 3 * (size_t)h->cap, b, n);
    h->launches++;
    h->nCur = n; h->broodValid = false;
  }
  if (cycleTransport(h, rng_state, history_offset, k_eff, phase)) return -1;
  if (cycleCloseEnqueue(h, h->dKsum)) return -1;
  if (resampleEnqueue(h, tot_pop, rng_state_resample, -1, 0, 1, -1)) return -1;
  {                                                                   // the normalised bank (tot_pop sites) and the BIN column of the cycle
    Bank& nb = h->bank[(h->cur + 1) % 3];
    const int m = tot_pop;
    k_bank_pack<<<gridFor(h, m, 256), 256, 0, st>>>(nb, h->dStage, h->dStage + 3 * (size_t)h->cap, m);
    h->launches++;
    CUDA_OK(cudaMemcpyAsync(r_out, h->dStage, sizeof(double) * 3 * m, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(dir_out, h->dStage + 3 * (size_t)h->cap, sizeof(double) * 3 * m, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(w_out, nb.w, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    if (E_out) CUDA_OK(cudaMemcpyAsync(E_out, nb.E, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    else if (G_out) CUDA_OK(cudaMemcpyAsync(G_out, nb.G, sizeof(int) * m, cudaMemcpyDeviceToHost, st));
    if (bins_out && h->nBins[phase] > 0) CUDA_OK(cudaMemcpyAsync(bins_out, h->dLast[phase], sizeof(double) * h->nBins[phase], cudaMemcpyDeviceToHost, st));
  }
  if (cycleFinish(h, res)) return -1;                      // the one synchronisation
  if (h->hCd->nSites <= 0) { h->err = "sb_resample: the fission bank is empty"; return -1; }
  if (resampleFinish(h, nullptr, false)) return -1;
  if (h->nCur != tot_pop) { h->err = "sb_run_cycle_resample_host: normSize_Repr did not return the target population"; return -1; }
  if (n_out) *n_out = h->nCur;
  return 0;
}

int sb_resample(sb_engine* h, int totPop, uint64_t rng_state) {
  if (2 * totPop > h->cap) { h->err = "sb_resample: target population exceeds the bank capacity"; return -1; }
  return resampleImpl(h, totPop, rng_state, -1, 0, 1, nullptr);
}

int sb_resample_ranked(sb_engine* h, int tot_pop, uint64_t master_rng_state, int n_ranks, int rank, const int32_t* pop_sizes, int32_t* new_local_pop) {
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks || !pop_sizes) { h->err = "sb_resample_ranked: invalid rank arguments"; return -1; }
  long long tot = 0, off = 0;
  for (int i = 0; i < n_ranks; ++i) { if (i < rank) off += pop_sizes[i]; tot += pop_sizes[i]; }
  if (tot > 2000000000LL) { h->err = "sb_resample_ranked: more than 2^31 sites"; return -1; }
  if (pop_sizes[rank] != std::min(h->hCd->nSites, h->cap)) { h->err = "sb_resample_ranked: pop_sizes[rank] is not this rank's bank size"; return -1; }
  return resampleImpl(h, tot_pop, master_rng_state, (int)tot, (int)off, 0, new_local_pop);
}

// sb_cycle_end + sb_resample_ranked with one synchronisation. host_sums: the 6 score sums reduced over ranks (host memory);
// new_sizes[n_ranks]: every rank's bank size after normSize_Repr (computed here, no second all-gather needed)
int sb_cycle_end_resample_ranked(sb_engine* h, const double* host_sums, int tot_pop, uint64_t master_rng_state, int n_ranks, int rank,
                                 const int32_t* pop_sizes, int32_t* new_sizes, sb_cycle_result* res) {
  if (n_ranks < 1 || n_ranks > 64 || rank < 0 || rank >= n_ranks || !pop_sizes) { h->err = "sb_cycle_end_resample_ranked: invalid rank arguments (at most 64 ranks)"; return -1; }
  CUDA_OK(cudaSetDevice(h->device));
  long long tot = 0, off = 0; RankOffs ro{}; ro.n = n_ranks;
  for (int i = 0; i < n_ranks; ++i) { ro.off[i] = (int)tot; if (i < rank) off += pop_sizes[i]; tot += pop_sizes[i]; }
  ro.off[n_ranks] = (int)tot;
  if (tot > 2000000000LL || tot <= 0) { h->err = "sb_cycle_end_resample_ranked: invalid total number of sites"; return -1; }
  if (!h->dRankCounts) { CUDA_OK(cudaMalloc(&h->dRankCounts, 64 * sizeof(int))); CUDA_OK(cudaMallocHost(&h->hRankCounts, 64 * sizeof(int))); }
  cudaStream_t st = h->stream;
  if (!h->dKsumRed) CUDA_OK(cudaMalloc(&h->dKsumRed, 8 * sizeof(double)));
  CUDA_OK(cudaMemcpyAsync(h->dKsumRed, host_sums, 6 * sizeof(double), cudaMemcpyHostToDevice, st));     // (the rank's own sums stay in dKsum: user k-eff clerks)
  if (cycleCloseEnqueue(h, h->dKsumRed)) return -1;
  if (resampleEnqueue(h, tot_pop, master_rng_state, (int)tot, (int)off, 0, pop_sizes[rank])) return -1;
  CUDA_OK(cudaMemsetAsync(h->dRankCounts, 0, 64 * sizeof(int), st));
  pdlLaunch(k_norm_rank_counts, gridFor(h, tot, 256), 256, st, h->dRnGlobal, h->dCd, h->dNd, ro, h->dRankCounts);
  h->launches++;
  CUDA_OK(cudaMemcpyAsync(h->hRankCounts, h->dRankCounts, 64 * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (cycleFinish(h, res)) return -1;
  if (pop_sizes[rank] != std::min(h->hCd->nSites, h->cap)) { h->err = "sb_cycle_end_resample_ranked: pop_sizes[rank] is not this rank's bank size"; return -1; }
  if (resampleFinish(h, nullptr, false)) return -1;
  const long long excess = tot - tot_pop;
  const long long nCopies = excess < 0 ? (-excess) / tot : 0;
  for (int i = 0; i < n_ranks; ++i)
    new_sizes[i] = excess > 0 ? h->hRankCounts[i] : (excess == 0 ? pop_sizes[i] : (int)(pop_sizes[i] * (nCopies + 1) + h->hRankCounts[i]));
  if (new_sizes[rank] != h->nCur) { h->err = "sb_cycle_end_resample_ranked: inconsistent bank size after normalisation"; return -1; }
  return 0;
}

// ---- ranks of one node over peer memory -------------------------------------------------------------------------------------
static size_t peerRegionBytes(int cap) { return (sizeof(PeerBox) + 255) / 256 * 256 + 4 * peerStageBytes(cap); }
int sb_bank_capacity(sb_engine* h) {
  if (h->opt.max_pop < 1) { h->err = "sb_bank_capacity: set the options (max_pop) first"; return -1; }
  if (ensureCapacity(h, h->opt.max_pop)) return -1;
  return h->cap;
}
int sb_peer_create(sb_engine* h, int n_ranks, int rank, int stage_cap, void* ipc_handle) {
  if (n_ranks < 1 || n_ranks > PEER_MAX || rank < 0 || rank >= n_ranks || !ipc_handle) { h->err = "sb_peer_create: invalid rank arguments (at most 64 ranks)"; return -1; }
  if (h->peerRegion) { h->err = "sb_peer_create: the peer region of this engine exists already"; return -1; }
  if (h->opt.max_pop < 1) { h->err = "sb_peer_create: set the options (max_pop) first"; return -1; }
  if (ensureCapacity(h, h->opt.max_pop)) return -1;
  if (stage_cap < h->cap) { h->err = "sb_peer_create: stage_cap must be at least the largest bank capacity of the ranks (sb_bank_capacity)"; return -1; }
  CUDA_OK(cudaSetDevice(h->device));
  h->peerRanks = n_ranks; h->peerRank = rank; h->peerCap = stage_cap;
  const size_t bytes = peerRegionBytes(stage_cap);
  CUDA_OK(cudaMalloc(&h->peerRegion, bytes));
  CUDA_OK(cudaMemset(h->peerRegion, 0, (sizeof(PeerBox) + 255) / 256 * 256));
  CUDA_OK(cudaMalloc(&h->dPlan, sizeof(PeerPlan))); CUDA_OK(cudaMallocHost(&h->hPlan, sizeof(PeerPlan)));
  CUDA_OK(cudaMalloc(&h->dKsumTot, 8 * sizeof(double)));
  if (!h->dRankCounts) { CUDA_OK(cudaMalloc(&h->dRankCounts, 64 * sizeof(int))); CUDA_OK(cudaMallocHost(&h->hRankCounts, 64 * sizeof(int))); }
  const size_t nG = (size_t)n_ranks * (size_t)stage_cap;                 // the stream over all ranks' banks, upper bound
  if (nG > h->rnGlobalCap) { cudaFree(h->dRnGlobal); h->rnGlobalCap = nG; CUDA_OK(cudaMalloc(&h->dRnGlobal, sizeof(unsigned long long) * nG)); }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t hd;
  CUDA_OK(cudaIpcGetMemHandle(&hd, h->peerRegion));
  memcpy(ipc_handle, &hd, sizeof(hd));
  CUDA_OK(cudaDeviceSynchronize());
  return 0;
}
int sb_peer_attach(sb_engine* h, const void* ipc_handles, const int32_t* caps) {
  if (!h->peerRegion) { h->err = "sb_peer_attach: sb_peer_create first"; return -1; }
  CUDA_OK(cudaSetDevice(h->device));
  const size_t boxBytes = (sizeof(PeerBox) + 255) / 256 * 256;
  for (int r = 0; r < h->peerRanks; ++r) {
    if (caps && caps[r] != h->peerCap) { h->err = "sb_peer_attach: every rank must have created its region with the same stage_cap"; return -1; }
    char* base = h->peerRegion;
    if (r != h->peerRank) {
      cudaIpcMemHandle_t hd; memcpy(&hd, (const char*)ipc_handles + 64 * (size_t)r, 64);
      void* ptr = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) { h->err = std::string("sb_peer_attach: cudaIpcOpenMemHandle failed for rank ") + std::to_string(r) + ": " + cudaGetErrorString(e); cudaGetLastError(); return -1; }
      h->peerOpened[r] = ptr; base = (char*)ptr;
    }
    h->peerPtrs.box[r] = (PeerBox*)base; h->peerPtrs.stage[r] = base + boxBytes;
  }
  h->peerSeq = 0;
  return 0;
}
int sb_peer_capacity(sb_engine* h) { return h->peerCap; }
int sb_peer_set_timeout(sb_engine* h, double seconds) { if (!(seconds > 0.0)) { h->err = "sb_peer_set_timeout: must be +ve"; return -1; } h->peerTimeoutS = seconds; return 0; }

// One cycle of one rank with the other ranks' data arriving through peer memory: transport, all-to-all of the score sums and bank
// sizes, cycle close, normSize_Repr over the global stream, loadBalancing with the two neighbours - one host synchronisation.
int sb_run_cycle_ranked_peer(sb_engine* h, uint64_t rng_state, int history_offset, double k_eff, int phase, int tot_pop, uint64_t master_rng_resample,
                             int32_t* final_sizes, sb_cycle_result* res) {
  if (!h->peerPtrs.box[h->peerRank]) { h->err = "sb_run_cycle_ranked_peer: sb_peer_attach first"; return -1; }
  if (h->cap > h->peerCap) { h->err = "sb_run_cycle_ranked_peer: the bank capacity grew beyond the stage capacity of sb_peer_create"; return -1; }
  if (cycleTransport(h, rng_state, history_offset, k_eff, phase)) return -1;
  cudaStream_t st = h->stream;
  const unsigned long long seq = ++h->peerSeq, tmo = (unsigned long long)(h->peerTimeoutS * 1.0e9);
  const int par = (int)(seq & 1ULL), nr = h->peerRanks;
  pdlLaunch(k_peer_post_wait, 1, PEER_MAX, st, h->peerPtrs, nr, h->peerRank, seq, h->dKsum, h->dKsumTot, h->dPlan, h->dCd, h->cap, tmo);
  if (h->profiling) CUDA_OK(cudaEventRecord(h->evP1, st));
  if (cycleCloseEnqueue(h, h->dKsumTot)) return -1;
  {  // normSize_Repr with the global sizes taken from the plan on the device (resampleEnqueue with host-known sizes otherwise)
    Bank& sorted = h->bank[(h->cur + 2) % 3]; Bank& dst = h->bank[(h->cur + 1) % 3];
    const long long nGmax = (long long)nr * h->peerCap; const int nSites = h->cap;
    unsigned long long* rn = h->dRnGlobal;
    pdlLaunch(k_norm_setup_plan, 1, 1, st, h->dNd, h->dCd, h->cap, tot_pop, h->dPlan);
    int g = gridFor(h, nGmax, 256), gl = gridFor(h, nSites, 256);
    pdlLaunch(k_rn_generate, gridFor(h, (nGmax + RN_CHUNK - 1) / RN_CHUNK, 128), 128, st, rn, h->dNd, master_rng_resample);
    pdlLaunch(k_zero_int, gridFor(h, SEL_BINS, 256), 256, st, h->dHist, SEL_BINS);
    pdlLaunch(k_sel_hist, g, 256, st, rn, h->dNd, h->dHist);
    pdlLaunch(k_sel_find_bin, 1, 1024, st, h->dHist, h->dCd, h->dNd);
    pdlLaunch(k_sel_collect, g, 256, st, rn, h->dCd, h->dNd, h->dCand);
    pdlLaunch(k_sel_threshold, 1, 1024, st, h->dCand, h->dCd);
    pdlLaunch(k_norm_flags, gl, 256, st, rn, h->dCd, h->dNd, h->dFlag);
    int tiles = (nSites + SCAN_TILE - 1) / SCAN_TILE;
    pdlLaunch(k_scan_reduce, tiles, SCAN_BLOCK, st, h->dFlag, &h->dNd->nLocal, h->dTile);
    pdlLaunch(k_scan_tiles, 1, 1024, st, h->dTile, &h->dNd->nLocal, nullptr);
    pdlLaunch(k_scan_apply, tiles, SCAN_BLOCK, st, h->dFlag, &h->dNd->nLocal, h->dTile, h->dFlagOff);
    pdlLaunch(k_norm_scatter, gl, 256, st, sorted, dst, h->dFlag, h->dFlagOff, h->dOffsets, h->dNsites, h->dNd, h->cap, h->dCd);
    pdlLaunch(k_norm_count, 1, 1, st, h->dFlag, h->dFlagOff, h->dNd, h->dCd);
    CUDA_OK(cudaMemsetAsync(h->dRankCounts, 0, 64 * sizeof(int), st));
    pdlLaunch(k_norm_rank_counts_plan, g, 256, st, rn, h->dCd, h->dNd, h->dPlan, h->dRankCounts);
    pdlLaunch(k_peer_plan, 1, 1, st, h->dPlan, h->dRankCounts, h->dNd, h->dCd, h->cap, h->peerCap);
    // loadBalancing: push to the neighbours, flag, wait for what they push, rebuild
    pdlLaunch(k_peer_push, gridFor(h, h->cap / 8 + 1, 256), 256, st, h->peerPtrs, h->dPlan, dst, h->peerCap, par);
    pdlLaunch(k_peer_flag_sites, 1, 32, st, h->peerPtrs, h->dPlan, seq);
    pdlLaunch(k_peer_wait_sites, 1, 32, st, h->peerPtrs, h->dPlan, seq, h->dCd, tmo);
    pdlLaunch(k_peer_splice, gl, 256, st, h->peerPtrs, h->dPlan, dst, sorted, h->peerCap, par, h->dCd);
    h->launches += 20;
  }
  if (h->profiling) { CUDA_OK(cudaEventRecord(h->evP2, st)); h->peerStagesOpen = true; }
  CUDA_OK(cudaMemcpyAsync(h->hPlan, h->dPlan, sizeof(PeerPlan), cudaMemcpyDeviceToHost, st));
  if (cycleFinish(h, res)) return -1;                      // the one synchronisation
  if (h->hCd->nSites <= 0) { h->err = "sb_resample: the fission bank is empty"; return -1; }
  h->cur = (h->cur + 2) % 3;
  h->nCur = h->hPlan->finalLocal; h->broodValid = true;
  h->kNormNext = h->kCumLast;
  h->sortedReady = false;
  if (final_sizes) for (int i = 0; i < nr; ++i) final_sizes[i] = h->hPlan->finalSizes[i];
  return 0;
}

// loadBalancing (particleDungeon_class.f90:607-698): sites leave from / arrive at the two ends of the bank
size_t sb_site_buffer_bytes(int k) { return (size_t)k * (8 * sizeof(double) + 2 * sizeof(int32_t)) + 8; }
int sb_bank_export(sb_engine* h, int k_front, void* dev_buf_front, int k_back, void* dev_buf_back) {
  CUDA_OK(cudaSetDevice(h->device));
  if (k_front < 0 || k_back < 0 || k_front + k_back > h->nCur) { h->err = "sb_bank_export: more sites requested than the bank holds"; return -1; }
  Bank& b = h->bank[h->cur];
  if (k_front > 0) { k_bank_pack_range<<<gridFor(h, k_front, 256), 256, 0, h->stream>>>(b, 0, k_front, (double*)dev_buf_front); h->launches++; }
  if (k_back > 0) { k_bank_pack_range<<<gridFor(h, k_back, 256), 256, 0, h->stream>>>(b, h->nCur - k_back, k_back, (double*)dev_buf_back); h->launches++; }
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaGetLastError());
  return 0;
}
int sb_bank_splice(sb_engine* h, int drop_front, int drop_back, int add_front, const void* dev_buf_front, int add_back, const void* dev_buf_back) {
  CUDA_OK(cudaSetDevice(h->device));
  int keep = h->nCur - drop_front - drop_back;
  if (drop_front < 0 || drop_back < 0 || add_front < 0 || add_back < 0 || keep < 0) { h->err = "sb_bank_splice: invalid counts"; return -1; }
  int nNew = add_front + keep + add_back;
  if (nNew > h->cap) { h->err = "Run out of space for particles (loadBalancing)"; return -1; }
  if (drop_front == 0 && drop_back == 0 && add_front == 0 && add_back == 0) return 0;
  Bank& src = h->bank[h->cur]; Bank& dst = h->bank[(h->cur + 1) % 3];
  cudaStream_t st = h->stream;
  if (add_front > 0) { k_bank_unpack_range<<<gridFor(h, add_front, 256), 256, 0, st>>>(dst, 0, add_front, (const double*)dev_buf_front); h->launches++; }
  if (keep > 0) { k_bank_copy_range<<<gridFor(h, keep, 256), 256, 0, st>>>(src, drop_front, keep, dst, add_front); h->launches++; }
  if (add_back > 0) { k_bank_unpack_range<<<gridFor(h, add_back, 256), 256, 0, st>>>(dst, add_front + keep, add_back, (const double*)dev_buf_back); h->launches++; }
  CUDA_OK(cudaStreamSynchronize(st));
  CUDA_OK(cudaGetLastError());
  h->cur = (h->cur + 1) % 3;
  h->nCur = nNew;
  return 0;
}

int64_t sb_tally_size(sb_engine* h, int phase) { return (phase < 0 || phase > 1) ? -1 : h->nBins[phase]; }
int sb_tally_read(sb_engine* h, int phase, double* csum, double* csum2, int32_t* batch_n) {
  if (phase < 0 || phase > 1) { h->err = "phase must be 0 or 1"; return -1; }
  CUDA_OK(cudaSetDevice(h->device));
  if (h->nBins[phase] > 0 && h->dCsum[phase]) {
    CUDA_OK(cudaMemcpy(csum, h->dCsum[phase], sizeof(double) * h->nBins[phase], cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(csum2, h->dCsum2[phase], sizeof(double) * h->nBins[phase], cudaMemcpyDeviceToHost));
  }
  *batch_n = h->batchN[phase];
  return 0;
}
int sb_tally_last_bins(sb_engine* h, int phase, double* bins) {
  if (phase < 0 || phase > 1) { h->err = "phase must be 0 or 1"; return -1; }
  CUDA_OK(cudaSetDevice(h->device));
  if (h->nBins[phase] > 0 && h->dLast[phase]) CUDA_OK(cudaMemcpy(bins, h->dLast[phase], sizeof(double) * h->nBins[phase], cudaMemcpyDeviceToHost));
  return 0;
}

// ---- measurement -----------------------------------------------------------------------------------
int sb_profile_enable(sb_engine* h, int on) { h->profiling = on != 0; h->msPeerWait = 0.0; h->msPeerTail = 0.0; h->msHistories = 0.0; h->nHistLaunches = 0; h->segProfiled = 0; h->scoreProfiled = 0; return 0; }
int sb_profile_read(sb_engine* h, double* ms_histories, int64_t* n_launches, int64_t* n_segments, int64_t* n_scores) {
  *ms_histories = h->msHistories; *n_launches = h->nHistLaunches; *n_segments = h->segProfiled; *n_scores = h->scoreProfiled; return 0;
}
int sb_profile_peer_stages(sb_engine* h, double* ms_wait, double* ms_tail) { *ms_wait = h->msPeerWait; *ms_tail = h->msPeerTail; return 0; }
int sb_timer_begin(sb_engine* h) { CUDA_OK(cudaSetDevice(h->device)); CUDA_OK(cudaStreamSynchronize(h->stream)); CUDA_OK(cudaEventRecord(h->evT0, h->stream)); return 0; }
int sb_timer_end(sb_engine* h, double* ms) {
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaEventRecord(h->evT1, h->stream)); CUDA_OK(cudaEventSynchronize(h->evT1));
  float f = 0.f; CUDA_OK(cudaEventElapsedTime(&f, h->evT0, h->evT1)); *ms = f; return 0;
}
void* sb_pinned_alloc(size_t bytes) { void* p = nullptr; if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr; return p; }
void sb_pinned_free(void* p) { if (p) cudaFreeHost(p); }
// writes `bytes` of device memory (> L2) so that the next timed step starts with a cold L2
int sb_flush_l2(sb_engine* h, size_t bytes) {
  CUDA_OK(cudaSetDevice(h->device));
  if (ensureStage(h, bytes)) return -1;
  CUDA_OK(cudaMemsetAsync(h->dStage, 0, bytes, h->stream));
  return 0;
}

// ---- continuous-energy lookup ----------------------------------------------------------------------
static int ceLaunch(sb_engine* h, int64_t n, const double* dE, const int* dMat, double* dT, double* dM, double* dJ, int* dIdx, int probeNuc, bool sorted = false) {
  if (!h->ce.loaded) { h->err = "continuous-energy data has not been loaded (sb_load_ce_data)"; return -1; }
  if (!h->dCeErr) { CUDA_OK(cudaMalloc(&h->dCeErr, sizeof(int))); CUDA_OK(cudaEventCreate(&h->evC0)); CUDA_OK(cudaEventCreate(&h->evC1)); }
  CUDA_OK(cudaMemsetAsync(h->dCeErr, 0, sizeof(int), h->stream));
  int blocks = (int)std::min<long long>((n + 255) / 256, (long long)h->numSM * 8);
  if (blocks < 1) blocks = 1;
  CUDA_OK(cudaEventRecord(h->evC0, h->stream));
  const int* perm = nullptr;
  if (sorted && n > 1) {                                     // bin the lookups by (material, energy): counting sort on the engine's stream
    if (n > 0x7fffffffLL) { h->err = "sb_ce_lookup_sorted_device: at most 2^31 - 1 lookups per call"; return -1; }
    long long kLo, kHi; { double a = h->ce.dev.eMin, b = h->ce.dev.eMax; memcpy(&kLo, &a, 8); memcpy(&kHi, &b, 8); }
    kLo >>= (52 - sbce::SORT_MBITS); kHi >>= (52 - sbce::SORT_MBITS);
    const int nEb = (int)(kHi - kLo + 1), nBin = nEb * h->ce.dev.nMat;
    if ((size_t)n > h->cePermCap) { cudaFree(h->dCePerm); h->dCePerm = nullptr; CUDA_OK(cudaMalloc(&h->dCePerm, sizeof(int) * (size_t)n)); h->cePermCap = (size_t)n; }
    if ((size_t)nBin + 1 > h->ceBinCap) {
      cudaFree(h->dCeHist); cudaFree(h->dCeCursor); cudaFree(h->dCeTile); cudaFree(h->dCeNbin); h->dCeHist = h->dCeCursor = h->dCeTile = h->dCeNbin = nullptr;
      h->ceBinCap = (size_t)nBin + 1;
      CUDA_OK(cudaMalloc(&h->dCeHist, sizeof(int) * h->ceBinCap)); CUDA_OK(cudaMalloc(&h->dCeCursor, sizeof(int) * h->ceBinCap));
      CUDA_OK(cudaMalloc(&h->dCeTile, sizeof(int) * (h->ceBinCap / SCAN_TILE + 2))); CUDA_OK(cudaMalloc(&h->dCeNbin, sizeof(int)));
    }
    CUDA_OK(cudaMemsetAsync(h->dCeHist, 0, sizeof(int) * (size_t)nBin, h->stream));
    CUDA_OK(cudaMemcpyAsync(h->dCeNbin, &nBin, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    const int gs = (int)std::min<long long>((n + 255) / 256, (long long)h->numSM * 8);
    sbce::k_ce_sort_hist<<<gs, 256, 0, h->stream>>>(h->ce.dev, n, dE, dMat, nEb, kLo, h->dCeHist);
    const int tiles = (nBin + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_reduce<<<tiles, SCAN_BLOCK, 0, h->stream>>>(h->dCeHist, h->dCeNbin, h->dCeTile);
    k_scan_tiles<<<1, 1024, 0, h->stream>>>(h->dCeTile, h->dCeNbin, nullptr);
    k_scan_apply<<<tiles, SCAN_BLOCK, 0, h->stream>>>(h->dCeHist, h->dCeNbin, h->dCeTile, h->dCeCursor);
    sbce::k_ce_sort_scatter<<<gs, 256, 0, h->stream>>>(h->ce.dev, n, dE, dMat, nEb, kLo, h->dCeCursor, h->dCePerm);
    h->launches += 5;
    perm = h->dCePerm;
  }
  if (!h->ce.dev.idxTab && dT && !dM && !dJ && !dIdx)       // total cross sections of a library without the union table: the lean kernel
    sbce::k_ce_total_hashed<<<(int)std::min<long long>((n + 255) / 256, (long long)h->numSM * 6), 256, 0, h->stream>>>(h->ce.dev, n, dE, dMat, dT, h->dCeErr, perm);
  else
    sbce::k_ce_lookup<<<blocks, 256, 0, h->stream>>>(h->ce.dev, n, dE, dMat, dT, dM, dJ, dIdx, probeNuc, h->dCeErr, perm);
  CUDA_OK(cudaEventRecord(h->evC1, h->stream));
  h->launches++;
  int e = 0;
  CUDA_OK(cudaMemcpyAsync(&e, h->dCeErr, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventElapsedTime(&h->ceLastMs, h->evC0, h->evC1));
  if (e == 1) { h->err = "Failed to find energy in the nuclide energy grids (energy outside the bounds of the data)"; return -1; }
  if (e == 2) { h->err = "Invalid material index in continuous-energy lookup"; return -1; }
  return 0;
}
int sb_load_ce_data(sb_engine* h, const sb_ce_flat* d) {
  CUDA_OK(cudaSetDevice(h->device));
  h->err.clear();
  if (sbce::ceBuild(h->ce, d, h->err)) return -1;
  CUDA_OK(cudaDeviceSynchronize());                       // ceBuild uploads on the null stream
  sbce::k_ce_majorant<<<std::max(1, std::min((h->ce.dev.nUnion + 127) / 128, h->numSM * 8)), 128, 0, h->stream>>>(h->ce.dev, (double*)h->ce.dev.uMaj);
  h->launches++;
  CUDA_OK(cudaMemcpyAsync(h->ce.uMaj.data(), h->ce.dev.uMaj, sizeof(double) * h->ce.uMaj.size(), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaGetLastError());
  return 0;
}
// aceNeutronDatabase%init + activate for the eigenvalue driver (aceNeutronDatabase_class.f90:873-1163,1292-1328)
int sb_load_ce_model(sb_engine* h, const sb_ce_model* m) {
  CUDA_OK(cudaSetDevice(h->device));
  h->err.clear();
  if (!m || m->n_nuc < 1 || m->n_mat < 1 || !m->cards) { h->err = "sb_load_ce_model: invalid sizes"; return -1; }
  for (void* p : h->ceAllocs) cudaFree(p);
  h->ceAllocs.clear(); h->ceCards.assign(m->n_nuc, sbk::CardOut());
  try {
    for (int n = 0; n < m->n_nuc; ++n) sbk::ceProcessCard(m->cards[n], m->energy_per_fission, h->ceCards[n]);
  } catch (const std::exception& e) { h->err = e.what(); return -1; }
  // concatenated grids / main data for the lookup structures, concatenated tape + directory for the reactions
  std::vector<int> gsize(m->n_nuc), rows(m->n_nuc); std::vector<double> grid, data, tape; std::vector<sbk::CeNucRec> recs(m->n_nuc); std::vector<sbk::CeMtRec> mts;
  for (int n = 0; n < m->n_nuc; ++n) {
    sbk::CardOut& c = h->ceCards[n];
    gsize[n] = (int)c.grid.size(); rows[n] = c.rec.rows;
    grid.insert(grid.end(), c.grid.begin(), c.grid.end()); data.insert(data.end(), c.main.begin(), c.main.end());
    c.rec.base = (int)tape.size() - 1; c.rec.mtFirst = (int)mts.size();
    tape.insert(tape.end(), c.tape.begin(), c.tape.end());
    mts.insert(mts.end(), c.mt.begin(), c.mt.end());
    recs[n] = c.rec;
  }
  sb_ce_flat f{}; f.n_nuc = m->n_nuc; f.grid_size = gsize.data(); f.rows = rows.data(); f.grid = grid.data(); f.data = data.data();
  f.n_mat = m->n_mat; f.mat_off = m->mat_off; f.mat_nuc = m->mat_nuc; f.mat_dens = m->mat_dens;
  std::vector<int> active(m->active_mats, m->active_mats + std::max(0, m->n_active));
  if (active.empty()) { h->err = "sb_load_ce_model: no active material"; return -1; }
  if (sbce::ceBuild(h->ce, &f, h->err, &active)) return -1;
  CUDA_OK(cudaDeviceSynchronize());                       // ceBuild uploads on the null stream; the engine stream does not wait for it
  sbce::k_ce_majorant<<<std::max(1, std::min((h->ce.dev.nUnion + 127) / 128, h->numSM * 8)), 128, 0, h->stream>>>(h->ce.dev, (double*)h->ce.dev.uMaj);
  h->launches++;
  CUDA_OK(cudaMemcpyAsync(h->ce.uMaj.data(), h->ce.dev.uMaj, sizeof(double) * h->ce.uMaj.size(), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaGetLastError());
  sbc::CeModelDev& D = h->ceModel;
  D.xs = h->ce.dev;
  D.tape = ceModelUpload(h, tape); D.nuc = ceModelUpload(h, recs); D.mt = ceModelUpload(h, mts);
  if (!D.tape || !D.nuc || !D.mt) return -1;
  CUDA_OK(cudaDeviceSynchronize());
  D.minE = m->min_energy; D.maxE = m->max_energy; D.threshE = m->thresh_energy; D.threshA = m->thresh_mass; D.sourceE = m->source_energy;
  D.eLo = h->ce.dev.eMin; D.eHi = h->ce.dev.eMax;
  if (!(D.minE >= 0.0) || !(D.maxE >= 0.0) || D.minE >= D.maxE || D.threshE < 0 || D.threshA < 0) { h->err = "sb_load_ce_model: invalid neutronCEstd settings (minEnergy / maxEnergy / thresholds)"; return -1; }
  // the generic model blob carries the material count, fissile flags and collisionXS; no multigroup tables
  h->nMat = m->n_mat; h->nG = 0; h->isP1 = 0; h->collisionXS = m->collision_xs;
  h->xs.clear(); h->P0.clear(); h->prod.clear(); h->P1.clear(); h->chi.clear(); h->majorant.clear();
  h->fissile.assign(m->n_mat, 0);
  for (int i = 0; i < m->n_mat; ++i) for (int k = m->mat_off[i]; k < m->mat_off[i + 1]; ++k) if (recs[m->mat_nuc[k] - 1].fissile) h->fissile[i] = 1;
  h->haveData = true; h->ceMode = true; h->blobDirty = true;
  return 0;
}
int sb_ce_nuclide_info(sb_engine* h, int nuc_idx, int32_t* grid_size, int32_t* rows, int32_t* n_mt) {
  if (nuc_idx < 1 || nuc_idx > (int)h->ceCards.size()) { h->err = "sb_ce_nuclide_info: invalid nuclide index"; return -1; }
  const sbk::CardOut& c = h->ceCards[nuc_idx - 1];
  *grid_size = (int)c.grid.size(); *rows = c.rec.rows; *n_mt = c.rec.nMT;
  return 0;
}
int sb_ce_nuclide_data(sb_engine* h, int nuc_idx, double* grid, double* main_data, int32_t* mt_list) {
  if (nuc_idx < 1 || nuc_idx > (int)h->ceCards.size()) { h->err = "sb_ce_nuclide_data: invalid nuclide index"; return -1; }
  const sbk::CardOut& c = h->ceCards[nuc_idx - 1];
  std::copy(c.grid.begin(), c.grid.end(), grid); std::copy(c.main.begin(), c.main.end(), main_data);
  for (size_t i = 0; i < c.mt.size(); ++i) mt_list[i] = c.mt[i].MT;
  return 0;
}
int sb_bank_upload_ce(sb_engine* h, int n, const double* r, const double* dir, const double* w, const double* E) {
  if (ensureCapacity(h, std::max(n, h->opt.max_pop))) return -1;
  CUDA_OK(cudaSetDevice(h->device));
  if (ensureStage(h, sizeof(double) * 6 * (size_t)h->cap)) return -1;
  Bank& b = h->bank[h->cur];
  cudaStream_t st = h->stream;
  CUDA_OK(cudaMemcpyAsync(h->dStage, r, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(h->dStage + 3 * (size_t)h->cap, dir, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(b.w, w, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(b.E, E, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemsetAsync(b.G, 0, sizeof(int) * n, st));
  k_bank_unpack<<<gridFor(h, n, 256), 256, 0, st>>>(h->dStage, h->dStage + 3 * (size_t)h->cap, b, n);
  h->launches++;
  CUDA_OK(cudaStreamSynchronize(st));
  h->nCur = n; h->broodValid = false;
  return 0;
}
int sb_bank_download_ce(sb_engine* h, int cap, int* n, double* r, double* dir, double* w, double* E) {
  CUDA_OK(cudaSetDevice(h->device));
  int m = h->nCur;
  *n = m;
  if (m > cap) { h->err = "sb_bank_download: buffer too small"; return -1; }
  if (m == 0) return 0;
  if (ensureStage(h, sizeof(double) * 6 * (size_t)h->cap)) return -1;
  Bank& b = h->bank[h->cur];
  cudaStream_t st = h->stream;
  k_bank_pack<<<gridFor(h, m, 256), 256, 0, st>>>(b, h->dStage, h->dStage + 3 * (size_t)h->cap, m);
  h->launches++;
  CUDA_OK(cudaMemcpyAsync(r, h->dStage, sizeof(double) * 3 * m, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaMemcpyAsync(dir, h->dStage + 3 * (size_t)h->cap, sizeof(double) * 3 * m, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaMemcpyAsync(w, b.w, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaMemcpyAsync(E, b.E, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}
int sb_ce_union_size(sb_engine* h) { return h->ce.loaded ? h->ce.dev.nUnion : 0; }
int sb_ce_union(sb_engine* h, double* grid, double* majorant) {
  if (!h->ce.loaded) { h->err = "continuous-energy data has not been loaded"; return -1; }
  std::copy(h->ce.uGrid.begin(), h->ce.uGrid.end(), grid); std::copy(h->ce.uMaj.begin(), h->ce.uMaj.end(), majorant);
  return 0;
}
int sb_ce_lookup_device(sb_engine* h, int64_t n, const double* dE, const int32_t* dMat, double* dTotal, double* dMacro, double* dMajorant) {
  CUDA_OK(cudaSetDevice(h->device));
  return ceLaunch(h, n, dE, dMat, dTotal, dMacro, dMajorant, nullptr, 0);
}
int sb_ce_lookup_sorted_device(sb_engine* h, int64_t n, const double* dE, const int32_t* dMat, double* dTotal, double* dMacro, double* dMajorant) {
  CUDA_OK(cudaSetDevice(h->device));
  if (!dMat) { h->err = "sb_ce_lookup_sorted_device: material indices are required"; return -1; }
  return ceLaunch(h, n, dE, dMat, dTotal, dMacro, dMajorant, nullptr, 0, true);
}
int sb_ce_last_kernel_ms(sb_engine* h, double* ms) { *ms = h->ceLastMs; return 0; }
int sb_ce_memory(sb_engine* h, int64_t* raw_bytes, int64_t* index_bytes, int32_t* has_union_table) {
  if (!h->ce.loaded) { h->err = "continuous-energy data has not been loaded (sb_load_ce_data)"; return -1; }
  *raw_bytes = h->ce.rawBytes; *index_bytes = h->ce.indexBytes; *has_union_table = h->ce.dev.idxTab ? 1 : 0;
  return 0;
}
int sb_ce_lookup(sb_engine* h, int64_t n, const double* E, const int32_t* mat, double* total, double* macro, double* majorant) {
  CUDA_OK(cudaSetDevice(h->device));
  if (n <= 0) return 0;
  if (!h->ce.loaded) { h->err = "continuous-energy data has not been loaded (sb_load_ce_data)"; return -1; }
  if ((total || macro) && !mat) { h->err = "sb_ce_lookup: material indices are required for total / macro"; return -1; }
  // staging: E | mat | total | macro | majorant
  size_t need = sizeof(double) * (size_t)n * (1 + 1 + 1 + 8 + 1);
  if (ensureStage(h, need)) return -1;
  double* dE = h->dStage; int* dMat = (int*)(dE + n); double* dT = dE + 2 * n; double* dM = dE + 3 * n; double* dJ = dE + 11 * n;
  if (!h->dCeErr) { CUDA_OK(cudaMalloc(&h->dCeErr, sizeof(int))); CUDA_OK(cudaEventCreate(&h->evC0)); CUDA_OK(cudaEventCreate(&h->evC1)); }
  // Three-stage pipeline over chunks: host->device copies, lookups and device->host copies run on their own streams (both copy
  // engines busy, PCIe in both directions at once); with page-locked host arrays the call is bounded by the slower direction.
  if (!h->ceStreamIn) {
    CUDA_OK(cudaStreamCreateWithFlags(&h->ceStreamIn, cudaStreamNonBlocking)); CUDA_OK(cudaStreamCreateWithFlags(&h->ceStreamOut, cudaStreamNonBlocking));
    for (int i = 0; i < CE_PIPE; ++i) { CUDA_OK(cudaEventCreateWithFlags(&h->ceEvIn[i], cudaEventDisableTiming)); CUDA_OK(cudaEventCreateWithFlags(&h->ceEvK[i], cudaEventDisableTiming)); }
  }
  cudaStream_t st = h->stream, sIn = h->ceStreamIn, sOut = h->ceStreamOut;
  CUDA_OK(cudaMemsetAsync(h->dCeErr, 0, sizeof(int), st));
  CUDA_OK(cudaEventRecord(h->ceEvK[0], st));
  CUDA_OK(cudaStreamWaitEvent(sIn, h->ceEvK[0], 0));                 // the staging buffer may still be in use by earlier work of the main stream
  const int64_t chunk = std::max<int64_t>(1 << 18, (n + CE_PIPE - 1) / CE_PIPE);
  int c = 0;
  for (int64_t o = 0; o < n; o += chunk, ++c) {
    const int64_t m = std::min(chunk, n - o);
    CUDA_OK(cudaMemcpyAsync(dE + o, E + o, sizeof(double) * m, cudaMemcpyHostToDevice, sIn));
    if (mat) CUDA_OK(cudaMemcpyAsync(dMat + o, mat + o, sizeof(int) * m, cudaMemcpyHostToDevice, sIn));
    CUDA_OK(cudaEventRecord(h->ceEvIn[c], sIn));
    CUDA_OK(cudaStreamWaitEvent(st, h->ceEvIn[c], 0));
    int blocks = (int)std::min<long long>((m + 255) / 256, (long long)h->numSM * 8);
    if (!h->ce.dev.idxTab && total && !macro && !majorant && mat)
      sbce::k_ce_total_hashed<<<(int)std::min<long long>((m + 255) / 256, (long long)h->numSM * 6), 256, 0, st>>>(h->ce.dev, m, dE + o, dMat + o, dT + o, h->dCeErr, nullptr);
    else
      sbce::k_ce_lookup<<<blocks, 256, 0, st>>>(h->ce.dev, m, dE + o, mat ? dMat + o : nullptr, total ? dT + o : nullptr, macro ? dM + 8 * o : nullptr,
                                                majorant ? dJ + o : nullptr, nullptr, 0, h->dCeErr, nullptr);
    h->launches++;
    CUDA_OK(cudaEventRecord(h->ceEvK[c], st));
    CUDA_OK(cudaStreamWaitEvent(sOut, h->ceEvK[c], 0));
    if (total) CUDA_OK(cudaMemcpyAsync(total + o, dT + o, sizeof(double) * m, cudaMemcpyDeviceToHost, sOut));
    if (macro) CUDA_OK(cudaMemcpyAsync(macro + 8 * o, dM + 8 * o, sizeof(double) * 8 * m, cudaMemcpyDeviceToHost, sOut));
    if (majorant) CUDA_OK(cudaMemcpyAsync(majorant + o, dJ + o, sizeof(double) * m, cudaMemcpyDeviceToHost, sOut));
  }
  int e = 0;
  CUDA_OK(cudaMemcpyAsync(&e, h->dCeErr, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  CUDA_OK(cudaStreamSynchronize(sOut));
  CUDA_OK(cudaGetLastError());
  if (e == 1) { h->err = "Failed to find energy in the nuclide energy grids (energy outside the bounds of the data)"; return -1; }
  if (e == 2) { h->err = "Invalid material index in continuous-energy lookup"; return -1; }
  return 0;
}
int sb_ce_nuclide_index(sb_engine* h, int nuc_idx, int64_t n, const double* E, int32_t* idx) {
  CUDA_OK(cudaSetDevice(h->device));
  if (!h->ce.loaded || nuc_idx < 1 || nuc_idx > h->ce.nNuc) { h->err = "sb_ce_nuclide_index: invalid nuclide index"; return -1; }
  if (ensureStage(h, sizeof(double) * 2 * (size_t)n)) return -1;
  double* dE = h->dStage; int* dI = (int*)(dE + n);
  CUDA_OK(cudaMemcpyAsync(dE, E, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  if (ceLaunch(h, n, dE, nullptr, nullptr, nullptr, nullptr, dI, nuc_idx)) return -1;
  CUDA_OK(cudaMemcpy(idx, dI, sizeof(int) * n, cudaMemcpyDeviceToHost));
  return 0;
}

// ---- batch queries -------------------------------------------------------------------------------
int sb_geom_query(sb_engine* h, int64_t n, double* r, double* dir, const double* dist, int32_t* mat, int32_t* uid) {
  if (!h->haveData) {                        // geometry-only use: give the blob a dummy 1-group material table
    sb_mg_flat d{}; double z6[6] = {1, 0, 1, 0, 0, 0}, one = 1.0, zero = 0.0; int f = 0;
    d.n_mat = 1; d.n_g = 1; d.data = z6; d.P0 = &zero; d.prod = &one; d.P1 = nullptr; d.chi = &zero; d.fissile = &f; d.majorant = &one; d.collision_xs = 0.0;
    if (sb_load_mg_data(h, &d)) return -1;
  }
  if (buildBlob(h)) return -1;
  CUDA_OK(cudaSetDevice(h->device));
  double *dr, *du, *dd = nullptr; int *dm, *dq;
  CUDA_OK(cudaMalloc(&dr, sizeof(double) * 3 * n)); CUDA_OK(cudaMalloc(&du, sizeof(double) * 3 * n));
  CUDA_OK(cudaMalloc(&dm, sizeof(int) * n)); CUDA_OK(cudaMalloc(&dq, sizeof(int) * n));
  CUDA_OK(cudaMemcpy(dr, r, sizeof(double) * 3 * n, cudaMemcpyHostToDevice)); CUDA_OK(cudaMemcpy(du, dir, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
  if (dist) { CUDA_OK(cudaMalloc(&dd, sizeof(double) * n)); CUDA_OK(cudaMemcpy(dd, dist, sizeof(double) * n, cudaMemcpyHostToDevice)); }
  k_geom_query<<<gridFor(h, n, 128), 128, 0, h->stream>>>(h->M, h->dBlob, n, dr, du, dd, dm, dq);
  h->launches++;
  CUDA_OK(cudaStreamSynchronize(h->stream)); CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpy(r, dr, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost)); CUDA_OK(cudaMemcpy(dir, du, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(mat, dm, sizeof(int) * n, cudaMemcpyDeviceToHost)); CUDA_OK(cudaMemcpy(uid, dq, sizeof(int) * n, cudaMemcpyDeviceToHost));
  cudaFree(dr); cudaFree(du); cudaFree(dm); cudaFree(dq); cudaFree(dd);
  return 0;
}

int sb_mg_query(sb_engine* h, int64_t n, const int32_t* mat, const int32_t* G, double* total, double* majorant) {
  if (buildBlob(h)) return -1;
  CUDA_OK(cudaSetDevice(h->device));
  int *dm, *dg; double *dt, *dj;
  CUDA_OK(cudaMalloc(&dm, sizeof(int) * n)); CUDA_OK(cudaMalloc(&dg, sizeof(int) * n));
  CUDA_OK(cudaMalloc(&dt, sizeof(double) * n)); CUDA_OK(cudaMalloc(&dj, sizeof(double) * n));
  CUDA_OK(cudaMemcpy(dm, mat, sizeof(int) * n, cudaMemcpyHostToDevice)); CUDA_OK(cudaMemcpy(dg, G, sizeof(int) * n, cudaMemcpyHostToDevice));
  k_mg_query<<<gridFor(h, n, 256), 256, 0, h->stream>>>(h->M, h->dBlob, n, dm, dg, dt, dj);
  h->launches++;
  CUDA_OK(cudaStreamSynchronize(h->stream)); CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpy(total, dt, sizeof(double) * n, cudaMemcpyDeviceToHost)); CUDA_OK(cudaMemcpy(majorant, dj, sizeof(double) * n, cudaMemcpyDeviceToHost));
  cudaFree(dm); cudaFree(dg); cudaFree(dt); cudaFree(dj);
  return 0;
}

int sb_rng_query(int64_t n, const uint64_t* state, const int64_t* skip, uint64_t* out_state, double* out_real) {
  unsigned long long *ds, *dout; long long* dk; double* dr;
  if (cudaMalloc(&ds, 8 * n) != cudaSuccess) { g_globalErr = "sb_rng_query: no CUDA device / allocation failed"; return -1; }
  cudaMalloc(&dk, 8 * n); cudaMalloc(&dout, 8 * n); cudaMalloc(&dr, 8 * n);
  cudaMemcpy(ds, state, 8 * n, cudaMemcpyHostToDevice); cudaMemcpy(dk, skip, 8 * n, cudaMemcpyHostToDevice);
  k_rng_query<<<(int)std::min<long long>((n + 255) / 256, 1184), 256>>>(n, ds, dk, dout, dr);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(out_state, dout, 8 * n, cudaMemcpyDeviceToHost); cudaMemcpy(out_real, dr, 8 * n, cudaMemcpyDeviceToHost);
  cudaFree(ds); cudaFree(dk); cudaFree(dout); cudaFree(dr);
  if (e != cudaSuccess) { g_globalErr = cudaGetErrorString(e); return -1; }
  return 0;
}

int sb_math_query(int64_t n, const double* x, double* lg, double* sn, double* cs) {
  double *dx, *dl, *dsn, *dcs;
  if (cudaMalloc(&dx, 8 * n) != cudaSuccess) { g_globalErr = "sb_math_query: no CUDA device / allocation failed"; return -1; }
  cudaMalloc(&dl, 8 * n); cudaMalloc(&dsn, 8 * n); cudaMalloc(&dcs, 8 * n);
  cudaMemcpy(dx, x, 8 * n, cudaMemcpyHostToDevice);
  k_math_query<<<(int)std::min<long long>((n + 255) / 256, 1184), 256>>>(n, dx, dl, dsn, dcs);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(lg, dl, 8 * n, cudaMemcpyDeviceToHost); cudaMemcpy(sn, dsn, 8 * n, cudaMemcpyDeviceToHost); cudaMemcpy(cs, dcs, 8 * n, cudaMemcpyDeviceToHost);
  cudaFree(dx); cudaFree(dl); cudaFree(dsn); cudaFree(dcs);
  if (e != cudaSuccess) { g_globalErr = cudaGetErrorString(e); return -1; }
  return 0;
}

#ifdef SB_PROFILE_ROUNDS
int sb_profile_rounds(sb_engine* h, long long* out) { cudaStreamSynchronize(h->stream); return cudaMemcpy(out, h->dProfRounds, 8 * (36 * 148 * 384 + 12 * 148 * 16), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1; }
#endif
int sb_fastmath_check(int64_t n, uint64_t seed, int exp_span, int64_t* mismatches) {
  unsigned long long* d = nullptr;
  if (n < 1 || exp_span < 0 || exp_span > 500 || !mismatches) { g_globalErr = "sb_fastmath_check: invalid arguments"; return -1; }
  if (cudaMalloc(&d, 8) != cudaSuccess) { g_globalErr = "sb_fastmath_check: no CUDA device / allocation failed"; return -1; }
  cudaMemset(d, 0, 8);
  k_fastmath_check<<<1184, 256>>>(n, seed & RNG_MASK, exp_span, d);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long hbad = 0; cudaMemcpy(&hbad, d, 8, cudaMemcpyDeviceToHost); cudaFree(d);
  if (e != cudaSuccess) { g_globalErr = cudaGetErrorString(e); return -1; }
  *mismatches = (int64_t)hbad;
  return 0;
}

}  // extern "C"
