// ORACLE (test infrastructure, not product code).
// CPU restatement of SCONE's ACE card reader and the search / interpolation helpers the CE data uses.
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs use it.
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
  struct MTrec { int MT = 0; double Q = 0; int TY = 0; bool isCapture = false, CMframe = false; int XSp = 0, N_xs = 0, IE = 0, LOCB = -17, LOCC = -17; };
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
  void fromArrays(const std::string& zaid, double aw, double tz, const int* nxs, const int* jxs, const double* xss_, long n) {
    ZAID = zaid; AW = aw; TZ = tz;
    for (int i = 0; i < 16; ++i) NXS[i] = nxs[i];
    for (int i = 0; i < 32; ++i) JXS[i] = jxs[i];
    XSS.assign(xss_, xss_ + n);
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
      if (m.TY < 0) { m.TY = -m.TY; m.CMframe = true; m.isCapture = false; } else if (m.TY == 0) m.isCapture = true; else { m.CMframe = false; m.isCapture = false; }
      m.XSp = r2i(xss(JXS[5] + i)) + JXS[6];
      m.N_xs = r2i(xss(m.XSp)); m.IE = r2i(xss(m.XSp - 1));
      m.XSp += 1;
    }
    int NMTs = NXS[4];                                                 // reactions with secondary neutrons: LOCB (LAND block) and LOCC (LDLW block)
    for (int i = 0; i < NMTs; ++i) {
      mt[i].LOCB = r2i(xss(JXS[7] + 1 + i));
      if (mt[i].LOCB < -1) throw CeError("setMTdata: LOCB is -ve and different from -1");
      mt[i].LOCC = r2i(xss(JXS[9] + i));
    }
  }
  int LOCB(int MT) const {                                             // :595-618, elastic: :722-735
    if (MT == N_N_ELASTIC) return r2i(xss(JXS[7]));
    const MTrec& m = rec(MT);
    if (m.isCapture) throw CeError("LOCBforMT: MT reaction is capture. No LOCB data.");
    return m.LOCB;
  }
  int LOCC(int MT) const {
    if (MT == N_N_ELASTIC) throw CeError("LOCCforMT: elastic scattering has no energy law");
    return rec(MT).LOCC;
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
  double readReal() { return xss(head++); }
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

}  // namespace orc_ce
