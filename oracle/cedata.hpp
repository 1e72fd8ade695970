// ORACLE (test infrastructure, not product code).
// CPU restatement of SCONE's continuous-energy cross-section path: ACE card -> aceNeutronNuclide main data ->
// nuclide / material / majorant lookups.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs use it.
//
//   NuclearData/ceNeutronData/aceDatabase/aceNeutronNuclide_class.f90:342-453,737-948  search, totalXS, microXSs, init
//   NuclearData/ceNeutronData/aceDatabase/aceNeutronDatabase_class.f90:346-394,509-644,1044-1051,1330-1621
//                                              updateMajorantXS, updateTotalMatXS, updateMacroXSs, eBounds, initMajorant
//   NuclearData/Reactions/uncorrelatedReactionCE/fissionCE_class.f90:187-235,360-400
#pragma once
#include "ceace.hpp"
#include "cereact.hpp"

namespace orc_ce {

// ---------------------------------------------------------------------------------------------------------
struct Nuclide {                                                       // aceNeutronNuclide
  std::string ZAID; double mass = 0, kT = 0; bool fissile = false;
  std::vector<double> eGrid;
  int rows = NON_FISSILE_SIZE;
  std::vector<double> main;                                            // mainData(rows, N), column-major: main[(j-1)*rows + (row-1)]
  Release nuTotal, nuDelayed; double Q = 0;
  // reactions (aceNeutronNuclide_class.f90:95-110): elastic scattering, fission, MT reactions in the order of MTdata
  bool hasReactions = false;
  ElasticScatter elastic; FissionCE fission;
  struct MTdata { int MT = 0, firstIdx = 0; std::vector<double> xs; NeutronScatter kin; };
  std::vector<MTdata> mtData; int nMTinelastic = 0;

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
    elastic.init(ACE);
    hasReactions = true;
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
      fission.init(ACE, ACE.hasFIS ? N_FISSION : N_f);
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
    nMTinelastic = (int)scatterMT.size();
    mtData.assign(nMTinelastic, MTdata());
    for (int s = (int)scatterMT.size() - 1; s >= 0; --s) {
      const auto& m = ACE.rec(scatterMT[s]);
      auto xs = ACE.xsMT(m.MT);
      { MTdata& d = mtData[nMTinelastic - 1 - s]; d.MT = m.MT; d.firstIdx = m.IE; d.xs = xs; d.kin.init(ACE, m.MT); }
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
  // invertInelastic (:135-165): the MT reactions in stack-pop order
  int invertInelastic(double E, RNG& rand, int* which = nullptr) const {
    int idx; double f; search(idx, f, E);
    double XS = md(IESCATTER_XS, idx + 1) * f + (1.0 - f) * md(IESCATTER_XS, idx);
    XS = XS * rand.get();
    for (int i = 0; i < nMTinelastic; ++i) {
      int idxT = idx - mtData[i].firstIdx + 1;
      if (idxT < 1) continue;
      double topXS = mtData[i].xs.at(idxT), bottomXS = mtData[i].xs.at(idxT - 1);
      XS = XS - topXS * f - (1.0 - f) * bottomXS;
      if (XS <= 0.0) { if (which) *which = i; return mtData[i].MT; }
    }
    throw CeError("Failed to invert inelastic scattering in nuclide " + ZAID);
  }
  double scatterXS(int idx, double f) const { return md(ESCATTER_XS, idx + 1) * f + (1.0 - f) * md(ESCATTER_XS, idx); }
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
