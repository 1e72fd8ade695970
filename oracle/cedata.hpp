// ORACLE (test infrastructure, not product code).
// CPU restatement of SCONE's continuous-energy cross-section path: ACE card -> aceNeutronNuclide main data ->
// nuclide / material / majorant lookups.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs use it.
//
//   NuclearData/DataDecks/ACE/aceCard_class.f90:255-300,356-500,748-830,1290-1500   card layout, MT table, FIS/NU
//   NuclearData/ceNeutronData/aceDatabase/aceNeutronNuclide_class.f90:342-453,737-948  search, totalXS, microXSs, init
//   NuclearData/ceNeutronData/aceDatabase/aceNeutronDatabase_class.f90:346-394,509-644,1044-1051,1330-1621
//                                              updateMajorantXS, updateTotalMatXS, updateMacroXSs, eBounds, initMajorant
//   NuclearData/emissionENDF/releaseLawENDF/{polynomialRelease,tabularRelease}_class.f90, releaseLawENDFfactory_func.f90
//   NuclearData/NuclearDataStructures/endfTable/endfTable_class.f90:162-198 ; SharedModules/genericProcedures.f90:132-199,778-815
//   NuclearData/Reactions/uncorrelatedReactionCE/fissionCE_class.f90:187-235,360-400
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc_ce {

struct CeError : std::runtime_error { using std::runtime_error::runtime_error; };

constexpr int TOTAL_XS = 1, ESCATTER_XS = 2, IESCATTER_XS = 3, CAPTURE_XS = 4, FISSION_XS = 5, NU_FISSION = 6, KAPPA_XS = 7, PROMPT_NU_FISSION = 8;
constexpr int NON_FISSILE_SIZE = 4, FISSILE_SIZE = 8;
constexpr int N_N_ELASTIC = 2, N_N_INELASTIC = 4, N_FISSION = 18, N_f = 19;
constexpr double H235 = 202.27, Q235 = 193.406;

// genericProcedures.f90:132-166 ; 1-based result, <= 0 on failure
inline int binarySearch(const std::vector<double>& a, double value) {
  int bottom = 1, top = (int)a.size();
  if (top < 1 || value < a[bottom - 1] || value > a[top - 1]) return -1;
  for (int i = 0; i < 70; ++i) {
    int idx = (top + bottom) / 2;
    if (bottom == idx) return idx;
    if (a[idx - 1] <= value) bottom = idx; else top = idx;
  }
  return -2;
}
// genericProcedures.f90:186-199 linearFloorIdxClosed_Real
inline int linearFloor(const std::vector<double>& a, double value) {
  int N = (int)a.size();
  if (value > a[N - 1] || value < a[0]) return -1;
  for (int idx = N - 1; idx >= 1; --idx) if (a[idx - 1] <= value) return idx;
  return -1;
}
inline double interpolate(double xMin, double xMax, double yMin, double yMax, double x) {   // :778-786
  double f = (x - xMin) / (xMax - xMin);
  return yMax * f + (1 - f) * yMin;
}
inline double endfInterpolate(double x0, double x1, double y0, double y1, double x, int flag) {   // :791-815
  switch (flag) {
    case 1: return y0;
    case 2: return interpolate(x0, x1, y0, y1, x);
    case 3: return interpolate(std::log(x0), std::log(x1), y0, y1, std::log(x));
    case 4: return std::exp(interpolate(x0, x1, std::log(y0), std::log(y1), x));
    case 5: return std::exp(interpolate(std::log(x0), std::log(x1), std::log(y0), std::log(y1), std::log(x)));
    default: throw CeError("Unknown ENDF interpolation number");
  }
}

// releaseLawENDF: polynomial (LNU = 1) or tabular (LNU = 2)
struct Release {
  bool present = false, poly = false;
  std::vector<double> coeffs, x, y; std::vector<int> bounds, inter;
  double at(double E) const {
    if (poly) { double r = 0.0; for (int i = (int)coeffs.size(); i >= 1; --i) r = r * E + coeffs[i - 1]; return r; }
    int idx = linearFloor(x, E);
    if (idx < 0) throw CeError("endfTable at: search of grid failed");
    double x0 = x[idx - 1], x1 = x[idx], y0 = y[idx - 1], y1 = y[idx];
    if (bounds.empty()) return interpolate(x0, x1, y0, y1, E);
    if (bounds.size() == 1) return endfInterpolate(x0, x1, y0, y1, E, inter[0]);
    size_t b = 0; while (b + 1 < bounds.size() && bounds[b] < idx + 1) ++b;      // linearCeilingIdxOpen
    return endfInterpolate(x0, x1, y0, y1, E, inter[b]);
  }
  bool hasEnergy(double E) const { return poly ? true : (E >= x.front() && E <= x.back()); }
};

// ---------------------------------------------------------------------------------------------------------
struct AceCard {
  std::string ZAID; double AW = 0, TZ = 0;
  int NXS[16], JXS[32];
  std::vector<double> XSS;            // 1-based access through xss()
  struct MTrec { int MT = 0; double Q = 0; int TY = 0; bool isCapture = false; int XSp = 0, N_xs = 0, IE = 0; };
  std::vector<MTrec> mt;
  bool isFiss = false, hasFIS = false; int fissIE = 0, fissNE = 0, fissXSp = 0, promptNUp = 0, totalNUp = 0, delayNUp = 0;
  int head = 0;

  double xss(int i) const { return XSS.at(i - 1); }
  static int r2i(double r) { return (int)std::lround(r); }

  void readFromFile(const std::string& path, int lineNum) {          // aceCard_class.f90 readFromFile
    std::ifstream f(path);
    if (!f) throw CeError("Cannot open ACE file: " + path);
    std::string line;
    for (int i = 1; i < lineNum; ++i) if (!std::getline(f, line)) throw CeError("ACE file shorter than the requested line");
    if (!std::getline(f, line)) throw CeError("ACE header missing");
    ZAID = line.substr(0, 10);
    AW = std::stod(line.substr(10, 12)); TZ = std::stod(line.substr(22, 12));
    std::getline(f, line);
    for (int i = 0; i < 4; ++i) std::getline(f, line);
    for (int i = 0; i < 16; ++i) f >> NXS[i];
    for (int i = 0; i < 32; ++i) f >> JXS[i];
    XSS.resize(NXS[0]);
    for (int i = 0; i < NXS[0]; ++i) { if (!(f >> XSS[i])) throw CeError("ACE XSS array is truncated"); }
    setMTdata(); setFissionData();
  }
  int gridSize() const { return NXS[2]; }
  std::vector<double> ESZ(int block) const {                          // 0 grid, 1 total, 2 absorption, 3 elastic, 4 heating
    int N = NXS[2], ptr = JXS[0] + block * N;
    return std::vector<double>(XSS.begin() + ptr - 1, XSS.begin() + ptr - 1 + N);
  }
  void setMTdata() {                                                   // :1290-1396
    int NMT = NXS[3];
    mt.assign(NMT, MTrec());
    for (int i = 0; i < NMT; ++i) {
      MTrec& m = mt[i];
      m.MT = r2i(xss(JXS[2] + i)); m.Q = xss(JXS[3] + i); m.TY = r2i(xss(JXS[4] + i));
      if (m.TY < 0) { m.TY = -m.TY; m.isCapture = false; } else if (m.TY == 0) m.isCapture = true; else m.isCapture = false;
      m.XSp = r2i(xss(JXS[5] + i)) + JXS[6];
      m.N_xs = r2i(xss(m.XSp)); m.IE = r2i(xss(m.XSp - 1));
      m.XSp += 1;
    }
  }
  void setFissionData() {                                              // :1398-1447
    isFiss = JXS[1] != 0; hasFIS = JXS[20] != 0;
    promptNUp = totalNUp = delayNUp = 0;
    if (!isFiss) return;
    if (hasFIS) { int p = JXS[20]; fissIE = r2i(xss(p)); fissNE = r2i(xss(p + 1)); fissXSp = p + 2; }
    int KNU = r2i(xss(JXS[1]));
    if (KNU > 0) totalNUp = JXS[1];
    else if (KNU < 0) { promptNUp = JXS[1] + 1; totalNUp = JXS[1] + std::abs(KNU) + 1; }
    else throw CeError("KNU is equal to 0");
    if (JXS[23] > 0) delayNUp = JXS[23];
  }
  const MTrec& rec(int MT) const { for (auto& m : mt) if (m.MT == MT) return m; throw CeError("Given MT is not present in ACE card"); }
  std::vector<double> xsMT(int MT) const { const MTrec& m = rec(MT); return std::vector<double>(XSS.begin() + m.XSp - 1, XSS.begin() + m.XSp - 1 + m.N_xs); }
  // read head
  int readInt() { return r2i(xss(head++)); }
  std::vector<double> readReals(int N) { std::vector<double> r(XSS.begin() + head - 1, XSS.begin() + head - 1 + N); head += N; return r; }
  std::vector<int> readInts(int N) { std::vector<int> r(N); for (int i = 0; i < N; ++i) r[i] = readInt(); return r; }
  Release readNu(int ptr) {                                            // releaseLawENDFfactory_func.f90 allocateNu
    Release R; R.present = true; head = ptr;
    int LNU = readInt();
    if (LNU == 1) { R.poly = true; int N = readInt(); R.coeffs = readReals(N); }
    else if (LNU == 2) {
      int NR = readInt();
      if (NR != 0) { R.bounds = readInts(NR); R.inter = readInts(NR); }
      int N = readInt(); R.x = readReals(N); R.y = readReals(N);
    } else throw CeError("Unrecoginised LNU. Not 1 or 2");
    return R;
  }
};

// ---------------------------------------------------------------------------------------------------------
struct Nuclide {                                                       // aceNeutronNuclide
  std::string ZAID; double mass = 0, kT = 0; bool fissile = false;
  std::vector<double> eGrid;
  int rows = NON_FISSILE_SIZE;
  std::vector<double> main;                                            // mainData(rows, N), column-major: main[(j-1)*rows + (row-1)]
  Release nuTotal, nuDelayed; double Q = 0;

  double& md(int row, int j) { return main[(size_t)(j - 1) * rows + (row - 1)]; }
  double md(int row, int j) const { return main[(size_t)(j - 1) * rows + (row - 1)]; }
  int N() const { return (int)eGrid.size(); }

  double release(double E) const { return nuTotal.at(E); }
  double releaseDelayed(double E) const { if (!nuDelayed.present) return 0.0; if (!nuDelayed.hasEnergy(E)) return 0.0; return nuDelayed.at(E); }
  double releasePrompt(double E) const { return release(E) - releaseDelayed(E); }

  void init(AceCard& ACE, bool withDatabase = true) {                   // aceNeutronNuclide_class.f90:737-948
    ZAID = ACE.ZAID; fissile = ACE.isFiss; mass = ACE.AW; kT = ACE.TZ;
    int Ngrid = ACE.gridSize();
    rows = fissile ? FISSILE_SIZE : NON_FISSILE_SIZE;
    main.assign((size_t)rows * Ngrid, 0.0);
    eGrid = ACE.ESZ(0);
    { auto t = ACE.ESZ(1), e = ACE.ESZ(3), a = ACE.ESZ(2);
      for (int j = 1; j <= Ngrid; ++j) { md(TOTAL_XS, j) = t[j - 1]; md(ESCATTER_XS, j) = e[j - 1]; md(CAPTURE_XS, j) = a[j - 1]; } }
    if (fissile) {
      int bottom;
      std::vector<int> fissMTs; for (auto& m : ACE.mt) if (m.TY == 19) fissMTs.push_back(m.MT);
      if (ACE.hasFIS) {
        int N0 = ACE.fissIE; bottom = N0;
        for (int k = 0; k < ACE.fissNE; ++k) md(FISSION_XS, N0 + k) = ACE.xss(ACE.fissXSp + k);
      } else {
        if (fissMTs.empty()) throw CeError(ACE.ZAID + " seems to have NU data but no fission reactions among its MT numbers");
        bottom = Ngrid + 1;
        for (int MT : fissMTs) {
          const auto& m = ACE.rec(MT); bottom = std::min(bottom, m.IE);
          auto xs = ACE.xsMT(MT);
          for (int k = 0; k < m.N_xs; ++k) md(FISSION_XS, m.IE + k) = md(FISSION_XS, m.IE + k) + xs[k];
        }
      }
      // fissionCE init: total nu, delayed nu, Q of MT 18 (FIS block) or 19
      bool onlyOneNu = ACE.totalNUp != 0 && ACE.promptNUp == 0, withDelayed = ACE.delayNUp != 0;
      if (withDelayed && onlyOneNu) throw CeError("Prompt/Total Nu is given with delayed data. Which one is which?");
      nuTotal = ACE.readNu(ACE.totalNUp);
      Q = ACE.rec(ACE.hasFIS ? N_FISSION : N_f).Q;
      if (withDelayed) nuDelayed = ACE.readNu(ACE.delayNUp);
      double H_Q = withDatabase ? H235 / Q235 : 1.0;
      for (int i = bottom; i <= Ngrid; ++i) {
        md(NU_FISSION, i) = md(FISSION_XS, i) * release(eGrid[i - 1]);
        md(KAPPA_XS, i) = md(FISSION_XS, i) * Q * H_Q;
        md(PROMPT_NU_FISSION, i) = md(FISSION_XS, i) * releasePrompt(eGrid[i - 1]);
      }
    }
    // scattering MTs (secondary particles, not fission, not MT 4); a stack: processed in reverse card order
    std::vector<int> scatterMT;
    for (auto& m : ACE.mt) if (m.TY != 19 && !m.isCapture && m.MT != N_N_INELASTIC) scatterMT.push_back(m.MT);
    for (int s = (int)scatterMT.size() - 1; s >= 0; --s) {
      const auto& m = ACE.rec(scatterMT[s]);
      auto xs = ACE.xsMT(m.MT);
      int bottom = m.IE, top = bottom + (int)xs.size();
      for (int j = 1; j <= Ngrid; ++j)
        if (j >= bottom && j <= top && j - bottom < (int)xs.size()) md(IESCATTER_XS, j) = md(IESCATTER_XS, j) + xs[j - bottom];
    }
    // total = sum of the main channels
    int K = fissile ? FISSION_XS : CAPTURE_XS;
    for (int j = 1; j <= Ngrid; ++j) { double s = 0.0; for (int r = ESCATTER_XS; r <= K; ++r) s = s + md(r, j); md(TOTAL_XS, j) = s; }
  }
  void fromArrays(int n, int nrows, const double* grid, const double* data) {
    eGrid.assign(grid, grid + n); rows = nrows; fissile = nrows == FISSILE_SIZE; main.assign(data, data + (size_t)n * nrows);
  }
  void search(int& idx, double& f, double E) const {                    // :342-359
    idx = binarySearch(eGrid, E);
    if (idx <= 0) throw CeError("Failed to find energy for nuclide " + ZAID);
    double E_top = eGrid[idx], E_low = eGrid[idx - 1];
    f = (E - E_low) / (E_top - E_low);
  }
  double totalXS(int idx, double f) const { return md(TOTAL_XS, idx + 1) * f + (1.0 - f) * md(TOTAL_XS, idx); }   // :377-385
  void microXSs(double out[8], int idx, double f) const {              // :418-445 ; out = total, el, inel, capture, fission, nuFission, kappa, promptNu
    for (int r = 1; r <= 8; ++r) out[r - 1] = 0.0;
    for (int r = 1; r <= rows; ++r) out[r - 1] = md(r, idx + 1) * f + (1.0 - f) * md(r, idx);
  }
};

struct Material { std::vector<int> nuclides; std::vector<double> dens; };   // ceNeutronMaterial: nucIdx (1-based), atomic densities

struct Database {                                                      // aceNeutronDatabase (no TMS / URR / S(a,b): not in the bundled data)
  std::vector<Nuclide> nuclides; std::vector<Material> materials; std::vector<int> activeMat;
  double eBounds[2] = {0, 0};
  std::vector<double> eGridUnion, majorant;

  void finalise() {                                                    // :1044-1051
    eBounds[0] = nuclides[0].eGrid.front(); eBounds[1] = nuclides[0].eGrid.back();
    for (auto& n : nuclides) { eBounds[0] = std::max(eBounds[0], n.eGrid.front()); eBounds[1] = std::min(eBounds[1], n.eGrid.back()); }
    activeMat.clear(); for (size_t i = 0; i < materials.size(); ++i) activeMat.push_back((int)i + 1);
  }
  double totalMatXS(double E, int matIdx) const {                      // updateTotalMatXS :509-571
    const Material& m = materials.at(matIdx - 1);
    double tot = 0.0;
    for (size_t i = 0; i < m.nuclides.size(); ++i) {
      const Nuclide& n = nuclides[m.nuclides[i] - 1];
      int idx; double f; n.search(idx, f, E);
      tot = tot + m.dens[i] * n.totalXS(idx, f);
    }
    return tot * 1.0;
  }
  void macroXSs(double out[8], double E, int matIdx) const {           // updateMacroXSs :579-644 ; neutronMacroXSs%add
    const Material& m = materials.at(matIdx - 1);
    for (int r = 0; r < 8; ++r) out[r] = 0.0;
    for (size_t i = 0; i < m.nuclides.size(); ++i) {
      const Nuclide& n = nuclides[m.nuclides[i] - 1];
      int idx; double f; n.search(idx, f, E);
      double mic[8]; n.microXSs(mic, idx, f);
      double d = m.dens[i] * 1.0;
      for (int r = 0; r < 8; ++r) out[r] = out[r] + d * mic[r];
    }
  }
  void initMajorant() {                                                // :1330-1621 (sorted merge of the nuclide grids inside eBounds, duplicates removed)
    std::vector<char> used(nuclides.size(), 0);
    for (int mi : activeMat) for (int n : materials[mi - 1].nuclides) used[n - 1] = 1;
    std::vector<double> tmp;
    for (size_t n = 0; n < nuclides.size(); ++n) if (used[n]) for (double e : nuclides[n].eGrid) if (!(e < eBounds[0] || e > eBounds[1])) tmp.push_back(e);
    std::sort(tmp.begin(), tmp.end());
    tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
    eGridUnion = tmp;
    majorant.assign(eGridUnion.size(), 0.0);
    for (size_t i = 0; i < eGridUnion.size(); ++i) {
      double E = eGridUnion[i];
      if (E < eBounds[0]) E = eBounds[0];
      if (E > eBounds[1]) E = eBounds[1];
      double maj = 0.0;
      for (int mi : activeMat) maj = std::max(maj, totalMatXS(E, mi));
      majorant[i] = maj * (1.0 + 1.0e-06);
    }
  }
  double majorantXS(double E) const {                                  // updateMajorantXS :346-394
    int idx = binarySearch(eGridUnion, E);
    if (idx <= 0) throw CeError("Failed to find energy in unionised majorant grid");
    double E_top = eGridUnion[idx], E_low = eGridUnion[idx - 1];
    double f = (E - E_low) / (E_top - E_low);
    return majorant[idx] * f + (1.0 - f) * majorant[idx - 1];
  }
};

}  // namespace orc_ce
