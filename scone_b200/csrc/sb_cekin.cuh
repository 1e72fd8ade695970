// Continuous-energy reaction data on the device: the "tape".
//
// The engine keeps, per nuclide, the XSS array of its ACE card as it is (aceCard_class.f90:1454-1534 holds the same
// array) and samples outgoing angles and energies by reading the ACE blocks in place: the blocks are already flat
// arrays (energy grid, locators, { x, pdf, cdf } tables contiguous), which is what a coalesced device layout wants,
// and no pointer-chasing object tree has to be rebuilt.  The host walks every block once at load time (ceProcessCard):
// it validates what the reference validates at build time (sorted grids, interpolation flags, normalised CDFs),
// applies the one change the reference makes to the data (cdf(N) = 1, tabularPdf_class.f90:270), refuses what is
// not supported (correlated laws, 32-equiprobable-bin pdfs, ENDF interpolation other than histogram / lin-lin), and
// builds the directory records (CeNucRec / CeMtRec) with the positions the kernels start from.
//
// What the device functions restate:
//   NuclearData/NuclearDataStructures/pdf/tabularPdf_class.f90:49-92                tabularPdf%sample
//   NuclearData/NuclearDataStructures/endfTable/endfTable_class.f90:162-198         endfTable%at
//   NuclearData/emissionENDF/angleLawENDF/tabularAngle_class.f90:62-79              tabularAngle%sample
//   NuclearData/emissionENDF/energyLawENDF/{levelScattering,contTabularEnergy,maxwellSpectrum,evaporationSpectrum,
//       multipleEnergyLaws}_class.f90 ; NuclearDataStructures/pdf/maxwellEnergyPdf_class.f90:30-50
//   NuclearData/emissionENDF/releaseLawENDF/{polynomialRelease,tabularRelease}_class.f90
//   NuclearData/Reactions/uncorrelatedReactionCE/{elasticNeutronScatter,neutronScatter,fissionCE}_class.f90 sampleOut
//   NuclearData/ceNeutronData/aceDatabase/aceNeutronNuclide_class.f90:135-165,737-948 invertInelastic, init
// Positions are 1-based ACE positions of the card; tape index = base + position.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/scone_b200.h"
#include "sb_math.h"

#if defined(__CUDACC__)
#define SBK_HD __host__ __device__ inline
#else
#define SBK_HD inline
#endif

namespace sbk {

constexpr double PI = 3.14159265358979323846264338327950288;
constexpr double TWO_PI = 6.283185307179586476925286766559;
constexpr double SQRT_PI = 1.77245385090551602729816748334115;
constexpr double MIN_E = 1.0E-11;                    // universalVariables.f90 MINIMUM_ENERGY
constexpr double HUGE_D = 1.7976931348623157e308;

enum { KERR_SEARCH = 1, KERR_LAW = 2, KERR_REJECT = 3, KERR_INTERP = 4 };

struct CeNucRec {
  double awr, kT;
  int fissile, rows, base;
  int elAng;                       // position of the elastic angular block (0 = isotropic)
  int andPos, dlwPos;              // JXS(9), JXS(11)
  int nMT, mtFirst;                // inelastic MT records [mtFirst, mtFirst + nMT) in invertInelastic order
  int nuTotPos, nuDelPos;          // position of LNU of the total / delayed nu-bar (0 = none)
  int fisLawPos;                   // position of LNW of the prompt fission energy law
  int nPrec, precPos, precLocPos, precRootPos;   // NXS(8), JXS(25), JXS(26), JXS(27)
  int pad;
};
struct CeMtRec { int MT, firstIdx, xsPos, nXs, cmFrame, TY, relPos, angPos, lawPos, pad; };

struct Tape {
  const double* t; int base;
  SBK_HD double operator()(int pos) const {
#if defined(__CUDA_ARCH__)
    return __ldg(t + base + pos);
#else
    return t[base + pos];
#endif
  }
  SBK_HD int i(int pos) const {
    double v = (*this)(pos);
    return (int)(v < 0.0 ? v - 0.5 : v + 0.5);
  }
};

// genericProcedures.f90:132-166 binarySearch on tape[p0 .. p0+N-1]; 1-based result, <= 0 on failure.
// (linearSearchFloor, :186-199, returns the same index for a sorted array: the largest idx <= N-1 with a(idx) <= value)
SBK_HD int tapeSearch(const Tape& T, int p0, int N, double v) {
  if (N < 1 || v < T(p0) || v > T(p0 + N - 1)) return -1;
  int bottom = 1, top = N;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int it = 0; it < 70; ++it) {
    int idx = (top + bottom) / 2;
    if (bottom == idx) return idx;
    if (T(p0 + idx - 1) <= v) bottom = idx; else top = idx;
  }
  return -2;
}
SBK_HD double interp(double x0, double x1, double y0, double y1, double x) {          // genericProcedures.f90:778-786
  double f = (x - x0) / (x1 - x0);
  return y1 * f + (1.0 - f) * y0;
}
// endfTable%at for a table stored as NR, [bounds(NR), flags(NR)], N, x(N), y(N) at `pos`; *next = position after the table
SBK_HD double tapeTableAt(const Tape& T, int pos, double x, int* err, int* next = nullptr) {
  const int NR = T.i(pos);
  const int pb = pos + 1, pN = pos + 1 + 2 * NR;
  const int N = T.i(pN);
  const int px = pN + 1, py = px + N;
  if (next) *next = py + N;
  int idx = tapeSearch(T, px, N, x);
  if (idx <= 0) { *err = KERR_SEARCH; return 0.0; }
  const double x0 = T(px + idx - 1), x1 = T(px + idx), y0 = T(py + idx - 1), y1 = T(py + idx);
  int flag = 2;
  if (NR == 1) flag = T.i(pb + 1);
  else if (NR > 1) {                                                                   // linearCeilingIdxOpen on bounds for idx + 1
    int b = 0; while (b + 1 < NR && T.i(pb + b) < idx + 1) ++b;
    flag = T.i(pb + NR + b);
  }
  if (flag == 1) return y0;
  if (flag == 2) return interp(x0, x1, y0, y1, x);
  *err = KERR_INTERP; return 0.0;
}
SBK_HD bool tapeTableHas(const Tape& T, int pos, double x) {                            // endfTable%hasX
  const int NR = T.i(pos); const int pN = pos + 1 + 2 * NR; const int N = T.i(pN);
  return !(x > T(pN + N) || x < T(pN + 1));
}
// release law at a nu block (LNU at pos): polynomialRelease / tabularRelease
SBK_HD double tapeNu(const Tape& T, int pos, double E, int* err) {
  const int LNU = T.i(pos);
  if (LNU == 1) {
    const int N = T.i(pos + 1);
    double r = 0.0;
    for (int i = N; i >= 1; --i) r = r * E + T(pos + 1 + i);
    return r;
  }
  return tapeTableAt(T, pos + 1, E, err);
}
SBK_HD bool tapeNuHas(const Tape& T, int pos, double E) { return T.i(pos) == 1 ? true : tapeTableHas(T, pos + 1, E); }

}  // namespace sbk

// =====================================================================================================================
// device-only sampling
// =====================================================================================================================
#if defined(__CUDACC__)
namespace sbk {

// rng: the lane's LCG state, advanced through sbh::rngGet.  log / sincos / the table functions are kept out of line: one
// copy each in the kernel (instruction-cache footprint, see sb_track.cuh)
#define SBK_RNG(rng) sbh::rngGet(rng)
__device__ __noinline__ double kLog(double x) { return -sbd::negLogHot(x); }
__device__ __noinline__ void kSinCos(double x, double* s, double* c) { sbd::sincosHot(x, s, c); }
__device__ __noinline__ double tapeTableAtNI(const Tape& T, int pos, double x, int* err, int* next) { return tapeTableAt(T, pos, x, err, next); }
__device__ __noinline__ double tapeNuNI(const Tape& T, int pos, double E, int* err) { return tapeNu(T, pos, E, err); }

// tabularPdf%sample with x at px, pdf at px+NP, cdf at px+2NP
__device__ __noinline__ double tapePdfSample(const Tape& T, int px, int NP, int flag, double r, int* err) {
  int idx = tapeSearch(T, px + 2 * NP, NP, r);
  if (idx <= 0) { *err = KERR_SEARCH; return T(px); }
  idx = min(idx, NP - 1);
  const double ci = T(px + 2 * NP + idx - 1), pi = T(px + NP + idx - 1), x0 = T(px + idx - 1);
  if (flag == 1) return x0 + (r - ci) / pi;
  const double f = (T(px + NP + idx) - pi) / (T(px + idx) - x0);
  const double disc = pi * pi + 2 * f * (r - ci);
  if (f == 0.0 || disc < 0.0) return x0 + (r - ci) / pi;
  return x0 + (sqrt(disc) - pi) / f;
}
// angleLawENDF%sample: angPos = 0 -> isotropic, else tabularAngle block
__device__ __noinline__ double tapeSampleMu(const Tape& T, int angPos, int andPos, double E, uint64_t& rng, int* err) {
  if (angPos == 0) return 2.0 * SBK_RNG(rng) - 1.0;
  const int NE = T.i(angPos);
  int idx = tapeSearch(T, angPos + 1, NE, E);
  if (idx <= 0) { *err = KERR_SEARCH; idx = 1; }
  const double e0 = T(angPos + idx), e1 = T(angPos + idx + 1);
  const double eps = (E - e0) / (e1 - e0);
  const double r = SBK_RNG(rng);
  const int k = (r < eps) ? idx + 1 : idx;
  const int LC = T.i(angPos + NE + k);
  if (LC == 0) return 2.0 * SBK_RNG(rng) - 1.0;                      // isotropicMu
  const int q = andPos + (LC < 0 ? -LC : LC) - 1;
  const int flag = T.i(q), NP = T.i(q + 1);
  const double rr = SBK_RNG(rng);
  return tapePdfSample(T, q + 2, NP, flag, rr, err);                  // tabularMu
}
// one ENDF energy law with its data at position q (root = JXS(11) or JXS(27))
__device__ __noinline__ double tapeSampleLaw(const Tape& T, int LAW, int q, int root, double E_in, uint64_t& rng, int* err) {
  if (LAW == 3) return T(q + 1) * (E_in - T(q));                     // levelScattering: LDAT2 * (E_in - LDAT1)
  if (LAW == 4) {                                                    // contTabularEnergy%sample
    const int NR = T.i(q);
    const int pN = q + 1 + 2 * NR;
    const int NE = T.i(pN);
    const int pe = pN + 1, pl = pe + NE;
    int idx = tapeSearch(T, pe, NE, E_in);
    if (idx <= 0) { *err = KERR_SEARCH; idx = 1; }
    int flag = 2;
    if (NR > 0) {
      int b = -1;
      for (int k = 0; k < NR; ++k) if (T.i(q + 1 + k) >= idx) { b = k; break; }
      if (b < 0) { *err = KERR_SEARCH; b = NR - 1; }
      flag = T.i(q + 1 + NR + b);
    }
    const int t0 = root + T.i(pl + idx - 1) - 1;
    const int f0 = T.i(t0), n0 = T.i(t0 + 1);
    if (flag == 1) return tapePdfSample(T, t0 + 2, n0, f0, SBK_RNG(rng), err);
    if (flag != 2) { *err = KERR_INTERP; return E_in; }
    const int t1 = root + T.i(pl + idx) - 1;
    const int f1 = T.i(t1), n1 = T.i(t1 + 1);
    const double E_min_low = T(t0 + 2), E_max_low = T(t0 + 1 + n0), E_min_up = T(t1 + 2), E_max_up = T(t1 + 1 + n1);
    const double e0 = T(pe + idx - 1), e1 = T(pe + idx);
    const double eps = (E_in - e0) / (e1 - e0);
    const double E_min = E_min_low * (1.0 - eps) + eps * E_min_up;
    const double E_max = E_max_low * (1.0 - eps) + eps * E_max_up;
    const double r = SBK_RNG(rng);
    double E_out, factor;
    if (r < eps) {
      E_out = tapePdfSample(T, t1 + 2, n1, f1, SBK_RNG(rng), err);
      factor = (E_out - E_min_up) / (E_max_up - E_min_up);
    } else {
      E_out = tapePdfSample(T, t0 + 2, n0, f0, SBK_RNG(rng), err);
      factor = (E_out - E_min_low) / (E_max_low - E_min_low);
    }
    return E_min * (1.0 - factor) + factor * E_max;
  }
  if (LAW == 7 || LAW == 9) {
    int next = 0;
    const double Tn = tapeTableAtNI(T, q, E_in, err, &next);
    const double U = T(next);
    if (LAW == 7) {                                                  // maxwellSpectrum%sample + maxwellEnergyPdf sample_Johnk
      for (int it = 0; it < 1000; ++it) {
        const double r1 = SBK_RNG(rng), r2 = SBK_RNG(rng), r3 = SBK_RNG(rng);
        double s, c; kSinCos(0.5 * PI * r1, &s, &c);
        const double beta = c * c;
        const double gamma05 = -kLog(r2) * beta;
        const double E_out = (-kLog(r3) + gamma05) * Tn;
        if (E_out < E_in - U) return E_out;
      }
      *err = KERR_REJECT; return 0.0;
    }
    // evaporationSpectrum%sample: an unbounded rejection loop in the reference; just above a threshold the acceptance
    // probability ~ ((E_in - U) / T)^2 / 2 can be 1e-6 and below, so the bound is generous (a lane that reaches it flags an error)
    for (long long it = 0; it < 2000000000LL; ++it) {
      const double r1 = SBK_RNG(rng), r2 = SBK_RNG(rng);
      const double E_out = -Tn * kLog(r1 * r2);
      if (E_out <= E_in - U) return E_out;
    }
    *err = KERR_REJECT; return 0.0;
  }
  *err = KERR_LAW; return E_in;
}
// energyLawENDF%sample for the law chain that starts at lawPos (LNW, LAW, IDAT, NR, ..., NE, E(NE), P(NE))
__device__ __noinline__ double tapeSampleEnergy(const Tape& T, int lawPos, int root, double E_in, uint64_t& rng, int* err) {
  int LNW = T.i(lawPos);
  if (LNW == 0) return tapeSampleLaw(T, T.i(lawPos + 1), root + T.i(lawPos + 2) - 1, root, E_in, rng, err);
  double r = SBK_RNG(rng);                                           // multipleEnergyLaws%sample
  int p = lawPos;
  for (int it = 0; it < 100; ++it) {
    LNW = T.i(p);
    const int LAW = T.i(p + 1), IDAT = T.i(p + 2);
    const int NR = T.i(p + 3); const int pN = p + 4 + 2 * NR; const int N = T.i(pN);
    double E = E_in;
    E = fmax(E, T(pN + 1));
    E = fmin(E, T(pN + N));
    const double prob = tapeTableAtNI(T, p + 3, E, err, nullptr);
    if (r < prob) return tapeSampleLaw(T, LAW, root + IDAT - 1, root, E_in, rng, err);
    r = r - prob;
    if (LNW == 0) break;
    p = root + LNW - 1;
  }
  *err = KERR_LAW; return E_in;
}
// fissionCE%sampleOut (fissionCE_class.f90:187-235)
__device__ __noinline__ void tapeSampleFission(const Tape& T, const CeNucRec& n, double E_in, uint64_t& rng, double& mu, double& phi, double& E_out, int* err) {
  mu = 2.0 * SBK_RNG(rng) - 1.0;
  phi = TWO_PI * SBK_RNG(rng);
  double p_del = 0.0;
  if (n.nPrec > 0) {
    double del = 0.0;
    if (n.nuDelPos != 0 && tapeNuHas(T, n.nuDelPos, E_in)) del = tapeNuNI(T, n.nuDelPos, E_in, err);
    p_del = del / tapeNuNI(T, n.nuTotPos, E_in, err);
  }
  const double r1 = SBK_RNG(rng);
  if (r1 > p_del) { E_out = tapeSampleEnergy(T, n.fisLawPos, n.dlwPos, E_in, rng, err); return; }
  double r2 = SBK_RNG(rng);
  int p = n.precPos, g = 1;
  for (; g <= n.nPrec; ++g) {                                        // precursor block: DEC, NR, [..], NE, E(NE), P(NE)
    int next = 0;
    r2 = r2 - tapeTableAtNI(T, p + 1, E_in, err, &next);
    if (r2 < 0.0) break;
    p = next;
  }
  if (g > n.nPrec) g = n.nPrec;
  const int locc = T.i(n.precLocPos + g - 1);
  E_out = tapeSampleEnergy(T, n.precRootPos + locc - 1, n.precRootPos, E_in, rng, err);
}

}  // namespace sbk
#endif

// =====================================================================================================================
// host: one ACE card -> nuclide main data (aceNeutronNuclide%init) + directory records + validated tape
// =====================================================================================================================
namespace sbk {

struct CardOut {
  CeNucRec rec{}; std::vector<CeMtRec> mt;
  std::vector<double> grid, main;          // eGrid(N), mainData(rows, N) in Fortran storage order
  std::vector<double> tape;                // XSS with cdf(N) = 1 applied
};

struct CardWalker {
  std::vector<double>& X; const int* NXS; const int* JXS; std::string& err; std::string zaid;
  double x(int p) const { if (p < 1 || p > (int)X.size()) throw std::runtime_error(zaid + ": ACE position outside the XSS array"); return X[p - 1]; }
  int xi(int p, const char* what) const {
    double v = x(p); double r = std::nearbyint(v);
    if (std::fabs(v - r) > 1.0e-9) throw std::runtime_error(zaid + ": expected an integer in the ACE data (" + what + ")");   // aceCard real2Int
    return (int)r;
  }
  void sortedAsc(int p, int N, const char* what) const { for (int i = 1; i < N; ++i) if (x(p + i) < x(p + i - 1)) throw std::runtime_error(zaid + ": " + what + " is not sorted ascending"); }
  // endfTable: returns the position after the table
  int table(int pos, const char* what) {
    int NR = xi(pos, what);
    if (NR < 0) throw std::runtime_error(zaid + ": -ve number of interpolation regions");
    for (int k = 0; k < NR; ++k) { int fl = xi(pos + 1 + NR + k, what); if (fl != 1 && fl != 2) throw std::runtime_error(zaid + ": " + what + ": ENDF interpolation flag " + std::to_string(fl) + " is not supported on the device (histogram and lin-lin only)"); }
    int pN = pos + 1 + 2 * NR; int N = xi(pN, what);
    if (N < 2) throw std::runtime_error(zaid + ": " + what + ": table with fewer than 2 points");
    sortedAsc(pN + 1, N, what);
    if (NR > 0 && xi(pos + NR, what) != N) throw std::runtime_error(zaid + ": " + what + ": incomplete interpolation scheme");
    return pN + 1 + 2 * N;
  }
  void pdf(int px, int NP, int flag, const char* what) {                                   // tabularPdf initCdf
    if (flag != 1 && flag != 2) throw std::runtime_error(zaid + ": " + what + ": unrecognised interpolation flag of a tabular pdf");
    sortedAsc(px, NP, what); sortedAsc(px + 2 * NP, NP, what);
    for (int i = 0; i < NP; ++i) if (x(px + NP + i) < 0.0) throw std::runtime_error(zaid + ": " + what + ": pdf contains -ve values");
    if (std::fabs(x(px + 2 * NP)) > 1.0e-6) throw std::runtime_error(zaid + ": " + what + ": CDF does not begin with 0");
    if (std::fabs(x(px + 3 * NP - 1) - 1.0) > 1.0e-6) throw std::runtime_error(zaid + ": " + what + ": CDF does not end with 1");
    X[px + 3 * NP - 2] = 1.0;                                                              // self % cdf(size(cdf)) = ONE
  }
  void angle(int angPos) {                                                                 // tabularAngle%init
    if (angPos == 0) return;
    int NE = xi(angPos, "angular block NE");
    sortedAsc(angPos + 1, NE, "angular energy grid");
    for (int k = 1; k <= NE; ++k) {
      int LC = xi(angPos + NE + k, "angular locator");
      if (LC == 0) continue;
      if (LC > 0) throw std::runtime_error(zaid + ": 32 equiprobable bin angular distributions are not supported");
      int q = JXS[8] + (-LC) - 1;
      int flag = xi(q, "angular table flag"), NP = xi(q + 1, "angular table size");
      if (x(q + 2) != -1.0 || x(q + 1 + NP) != 1.0) throw std::runtime_error(zaid + ": mu grid does not begin with -1 and end with 1");
      pdf(q + 2, NP, flag, "angular table");
    }
  }
  void law(int LAW, int q, int root) {                                                     // buildENDFLaw
    if (LAW == 3) { double l2 = x(q + 1); if (l2 < 0.0 || l2 >= 1.0) throw std::runtime_error(zaid + ": level scattering LDAT2 outside [0, 1)"); return; }
    if (LAW == 4) {
      int NR = xi(q, "law 4 NR");
      for (int k = 0; k < NR; ++k) { int fl = xi(q + 1 + NR + k, "law 4 flag"); if (fl != 1 && fl != 2) throw std::runtime_error(zaid + ": continuous tabular law with interpolation other than histogram / lin-lin"); }
      int pN = q + 1 + 2 * NR; int NE = xi(pN, "law 4 NE");
      sortedAsc(pN + 1, NE, "law 4 energy grid");
      if (NR > 0 && xi(q + NR, "law 4 bounds") != NE) throw std::runtime_error(zaid + ": law 4: incomplete interpolation scheme");
      for (int k = 0; k < NE; ++k) {
        int t = root + xi(pN + 1 + NE + k, "law 4 locator") - 1;
        int INTT = xi(t, "law 4 INTT"), NP = xi(t + 1, "law 4 NP");
        if (INTT > 10) throw std::runtime_error(zaid + ": law 4 with discrete photon lines (INTT > 10) is not supported");
        pdf(t + 2, NP, INTT, "law 4 outgoing energy table");
      }
      return;
    }
    if (LAW == 7 || LAW == 9) { table(q, "law 7/9 temperature table"); return; }
    throw std::runtime_error(zaid + ": energy law " + std::to_string(LAW) + " is not supported (3, 4, 7, 9 are)");
  }
  void lawChain(int lawPos, int root) {                                                    // new_energyLawENDF
    int p = lawPos;
    for (int it = 0; it < 101; ++it) {
      int LNW = xi(p, "LNW"), LAW = xi(p + 1, "LAW"), IDAT = xi(p + 2, "IDAT");
      table(p + 3, "law applicability table");
      law(LAW, root + IDAT - 1, root);
      if (LNW == 0) return;
      p = root + LNW - 1;
    }
    throw std::runtime_error(zaid + ": energy law chain does not terminate");
  }
  void nu(int pos) {
    int LNU = xi(pos, "LNU");
    if (LNU == 1) { (void)xi(pos + 1, "NC"); return; }
    if (LNU == 2) { table(pos + 1, "nu-bar table"); return; }
    throw std::runtime_error(zaid + ": unrecognised LNU (not 1 or 2)");
  }
};

// aceNeutronNuclide%init + the reaction objects it builds; throws std::runtime_error with the reference's message where it has one
inline void ceProcessCard(const sb_ace_card& c, double H235, CardOut& out) {
  const int* NXS = c.nxs; const int* JXS = c.jxs;
  if (c.n_xss < 1 || NXS[0] != c.n_xss) throw std::runtime_error(std::string(c.zaid) + ": NXS(1) does not match the length of XSS");
  out.tape.assign(c.xss, c.xss + c.n_xss);
  std::string err;
  CardWalker W{out.tape, NXS, JXS, err, c.zaid ? c.zaid : "?"};
  const int Ngrid = NXS[2], NMT = NXS[3], NMTs = NXS[4];
  const bool fissile = JXS[1] != 0, hasFIS = JXS[20] != 0;
  const int rows = fissile ? 8 : 4;
  CeNucRec& R = out.rec;
  R.awr = c.aw; R.kT = c.tz; R.fissile = fissile; R.rows = rows; R.andPos = JXS[8]; R.dlwPos = JXS[10];
  // MT table (aceCard setMTdata)
  struct MTr { int MT, TY, XSp, N_xs, IE, LOCB, LOCC; double Q; bool isCapture, CM; };
  std::vector<MTr> mt(NMT);
  for (int i = 0; i < NMT; ++i) {
    MTr& m = mt[i];
    m.MT = W.xi(JXS[2] + i, "MT"); m.Q = W.x(JXS[3] + i); m.TY = W.xi(JXS[4] + i, "TY");
    m.CM = false; m.isCapture = false;
    if (m.TY < 0) { m.TY = -m.TY; m.CM = true; } else if (m.TY == 0) m.isCapture = true;
    m.XSp = W.xi(JXS[5] + i, "LOCA") + JXS[6];
    m.N_xs = W.xi(m.XSp, "NE of MT"); m.IE = W.xi(m.XSp - 1, "IE of MT");
    m.XSp += 1; m.LOCB = 0; m.LOCC = 0;
  }
  for (int i = 0; i < NMTs; ++i) { mt[i].LOCB = W.xi(JXS[7] + 1 + i, "LOCB"); mt[i].LOCC = W.xi(JXS[9] + i, "LOCC"); }
  auto rec = [&](int MT) -> const MTr& { for (auto& m : mt) if (m.MT == MT) return m; throw std::runtime_error("Given MT is not present in ACE card"); };
  // main data
  out.grid.assign(out.tape.begin() + JXS[0] - 1, out.tape.begin() + JXS[0] - 1 + Ngrid);
  out.main.assign((size_t)rows * Ngrid, 0.0);
  auto md = [&](int row, int j) -> double& { return out.main[(size_t)(j - 1) * rows + (row - 1)]; };
  for (int j = 1; j <= Ngrid; ++j) {
    md(1, j) = W.x(JXS[0] + Ngrid + j - 1); md(2, j) = W.x(JXS[0] + 3 * Ngrid + j - 1); md(4, j) = W.x(JXS[0] + 2 * Ngrid + j - 1);
  }
  // elastic scattering (elasticNeutronScatter buildFromACE)
  { int LOCB = W.xi(JXS[7], "elastic LOCB"); R.elAng = (LOCB == 0) ? 0 : JXS[8]; W.angle(R.elAng); }
  if (fissile) {
    int bottom;
    if (hasFIS) {
      int p = JXS[20]; int IE = W.xi(p, "FIS IE"), NE = W.xi(p + 1, "FIS NE"); bottom = IE;
      for (int k = 0; k < NE; ++k) md(5, IE + k) = W.x(p + 2 + k);
    } else {
      bottom = Ngrid + 1; bool any = false;
      for (auto& m : mt) if (m.TY == 19) {
        any = true; bottom = std::min(bottom, m.IE);
        for (int k = 0; k < m.N_xs; ++k) md(5, m.IE + k) = md(5, m.IE + k) + W.x(m.XSp + k);
      }
      if (!any) throw std::runtime_error(std::string(c.zaid) + " seems to have NU data but no fission reactions among its MT numbers");
    }
    // fissionCE buildFromACE
    int KNU = W.xi(JXS[1], "KNU");
    int promptNUp = 0, totalNUp = 0, delayNUp = JXS[23] > 0 ? JXS[23] : 0;
    if (KNU > 0) totalNUp = JXS[1];
    else if (KNU < 0) { promptNUp = JXS[1] + 1; totalNUp = JXS[1] + std::abs(KNU) + 1; }
    else throw std::runtime_error("KNU is equal to 0");
    const bool onlyOneNu = totalNUp != 0 && promptNUp == 0, withDelayed = delayNUp != 0;
    if (withDelayed && onlyOneNu) throw std::runtime_error("Prompt/Total Nu is given with delayed data. Which one is which? " + std::string(c.zaid));
    R.nuTotPos = totalNUp; W.nu(totalNUp);
    const MTr& fm = rec(hasFIS ? 18 : 19);
    R.fisLawPos = JXS[10] + fm.LOCC - 1; W.lawChain(R.fisLawPos, JXS[10]);
    const double Q = fm.Q;
    if (withDelayed) {
      R.nuDelPos = delayNUp; W.nu(delayNUp);
      R.nPrec = NXS[7];
      if (R.nPrec == 0) throw std::runtime_error("Has delayed neutrons but not precursors. WTF? " + std::string(c.zaid));
      if (JXS[24] == 0) throw std::runtime_error("Missing Fission Data. Cannot Locate Precursor PDF. JXS(25) == 0");
      R.precPos = JXS[24]; R.precLocPos = JXS[25]; R.precRootPos = JXS[26];
      int p = R.precPos;
      for (int g = 0; g < R.nPrec; ++g) p = W.table(p + 1, "precursor probability table");
      for (int g = 0; g < R.nPrec; ++g) W.lawChain(JXS[26] + W.xi(JXS[25] + g, "precursor LOCC") - 1, JXS[26]);
    }
    const double H_Q = H235 / 193.406;
    const Tape T{out.tape.data(), -1};
    for (int i = bottom; i <= Ngrid; ++i) {
      int e = 0;
      const double E = out.grid[i - 1];
      const double nuT = tapeNu(T, totalNUp, E, &e);
      double nuD = 0.0;
      if (withDelayed && tapeNuHas(T, delayNUp, E)) nuD = tapeNu(T, delayNUp, E, &e);
      if (e) throw std::runtime_error(std::string(c.zaid) + ": nu-bar table does not cover the energy grid");
      md(6, i) = md(5, i) * nuT;
      md(7, i) = md(5, i) * Q * H_Q;
      md(8, i) = md(5, i) * (nuT - nuD);
    }
  }
  // scattering MTs in the order invertInelastic walks them (stack pop = reverse card order)
  std::vector<int> scat;
  for (auto& m : mt) if (m.TY != 19 && !m.isCapture && m.MT != 4) scat.push_back(m.MT);
  for (int s = (int)scat.size() - 1; s >= 0; --s) {
    const MTr& m = rec(scat[s]);
    CeMtRec r{}; r.MT = m.MT; r.firstIdx = m.IE; r.xsPos = m.XSp; r.nXs = m.N_xs; r.cmFrame = m.CM ? 1 : 0; r.TY = m.TY;
    if (m.TY > 100) { r.relPos = JXS[10] + m.TY - 101; W.table(r.relPos, "energy dependent neutron yield"); }
    if (m.LOCB == -1) throw std::runtime_error(std::string(c.zaid) + ": MT " + std::to_string(m.MT) + " has a correlated angle-energy law (not supported)");
    if (m.LOCB < -1) throw std::runtime_error("For some reason LOCB is -ve and diffrent from unINIT. WTF?");
    r.angPos = (m.LOCB == 0) ? 0 : JXS[8] + m.LOCB - 1; W.angle(r.angPos);
    r.lawPos = JXS[10] + m.LOCC - 1; W.lawChain(r.lawPos, JXS[10]);
    out.mt.push_back(r);
    const int bottom = m.IE, top = bottom + m.N_xs;
    for (int j = 1; j <= Ngrid; ++j) if (j >= bottom && j <= top && j - bottom < m.N_xs) md(3, j) = md(3, j) + W.x(m.XSp + j - bottom);
  }
  R.nMT = (int)out.mt.size();
  const int K = fissile ? 5 : 4;
  for (int j = 1; j <= Ngrid; ++j) { double s = 0.0; for (int r = 2; r <= K; ++r) s = s + md(r, j); md(1, j) = s; }
}

}  // namespace sbk
