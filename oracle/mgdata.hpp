// ORACLE (test infrastructure, not product code).
// CPU restatement of SCONE's multigroup nuclear data:
//   NuclearData/materialMenu_mod.f90:161-196                          (matIdx = order in materials{})
//   NuclearData/mgNeutronData/baseMgNeutron/baseMgNeutronMaterial_class.f90:112-291
//   NuclearData/mgNeutronData/baseMgNeutron/baseMgNeutronDatabase_class.f90:95-236,343-503
//   NuclearData/Reactions/reactionMG/multiScatterMG_class.f90:199-351
//   NuclearData/Reactions/reactionMG/multiScatterP1MG_class.f90:69-128
//   NuclearData/Reactions/reactionMG/fissionMG_class.f90:183-274
//   NuclearData/xsPackages/neutronXsPackages_class.f90:143-190,211-250
//   SharedModules/legendrePoly_func.f90:35-93
#pragma once
#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "geom.hpp"
#include "rng.hpp"

namespace orc {

// SharedModules/endfConstants.f90:116-129
constexpr int macroTotal = -1, macroDisappearance = -2, macroEscatter = -3, macroIEscatter = -4,
              macroFission = -6, macroNuFission = -7, macroPromptNuFission = -8, macroDelayedNuFission = -9,
              macroKappaFission = -80, macroAllScatter = -20, macroAbsorbtion = -21, macroNonElastic = -22;

struct MacroXSs {
  double total = 0, elasticScatter = 0, inelasticScatter = 0, capture = 0, fission = 0, nuFission = 0,
         kappaXS = 0, promptNuFission = 0;
  double get(int MT) const {
    switch (MT) {
      case macroTotal: return total;
      case macroDisappearance: return capture;
      case macroEscatter: return elasticScatter;
      case macroNonElastic: return inelasticScatter + fission + capture;
      case macroIEscatter: return inelasticScatter;
      case macroAllScatter: return elasticScatter + inelasticScatter;
      case macroFission: return fission;
      case macroNuFission: return nuFission;
      case macroKappaFission: return kappaXS;
      case macroPromptNuFission: return promptNuFission;
      case macroDelayedNuFission: return nuFission - promptNuFission;
      case macroAbsorbtion: return fission + capture;
      default: return 0.0;
    }
  }
  int invert(double r) const {
    int C = 1;
    double xs = total * r - elasticScatter;
    if (xs > 0.0) C += 1;
    xs = xs - inelasticScatter;
    if (xs > 0.0) C += 1;
    xs = xs - capture;
    if (xs > 0.0) C += 1;
    switch (C) {
      case 1: return macroEscatter;
      case 2: return macroIEscatter;
      case 3: return macroDisappearance;
      case 4: return macroFission;
    }
    return INT32_MAX;
  }
};

inline double sampleLegendreP1(double P1, RNG& rand) {
  const int UNIFORM = 1, LIN = 2, DELTA = 3;
  double P1_loc = std::fabs(P1), threshold;
  int Low, Top;
  if (P1_loc < 1.0) { threshold = P1_loc; Top = LIN; Low = UNIFORM; }
  else if (P1_loc <= 3.0) { threshold = 0.5 * (P1_loc - 1.0); Top = DELTA; Low = LIN; }
  else throw FatalError("sampleLegendre_P1", "P1 must have absolute value < 3.0");
  int exec = (rand.get() < threshold) ? Top : Low;
  double x;
  if (exec == UNIFORM) x = 2.0 * rand.get() - 1.0;
  else if (exec == LIN) x = 2.0 * std::sqrt(rand.get()) - 1.0;
  else x = 1.0;
  if (P1 < 0.0) x = -x;
  return x;
}

struct MgMaterial {
  std::string name;
  int nG = 0;
  bool fissile = false, isP1 = false;
  // data rows, group-indexed (0-based): TOTAL, IESCATTER, CAPTURE, FISSION, NU_FISSION, KAPPA
  std::vector<double> total, iescatter, capture, fission, nuFission, kappaFission;
  // scattering: P0[gOut + nG*gIn] (Fortran P0(G_out,G_in) column-major; file order is read
  // straight into that storage, so the n-th number in the file is (gOut = n % nG, gIn = n / nG))
  std::vector<double> P0, prod, P1, scatterXSs;
  std::vector<double> nu, chi;

  void init(const Dict& d, const std::string& scatterKey) {
    nG = d.getInt("numberOfGroups");
    if (nG < 1) throw FatalError("init (baseMgNeutronMaterial)", "Number of groups is invalid");
    fissile = d.isPresent("fission");
    if (scatterKey == "P0") isP1 = false;
    else if (scatterKey == "P1") isP1 = true;
    else throw FatalError("init (baseMgNeutronMaterial)", "scatterKey must be P0 or P1");
    // multiScatterMG buildFromDict
    P0 = d.getRealArray("P0");
    if ((int)P0.size() != nG * nG) throw FatalError("buildFromDict (multiScatterMG)", "Invalid size of P0");
    prod = d.getRealArray("scatteringMultiplicity");
    if ((int)prod.size() != nG * nG) throw FatalError("buildFromDict (multiScatterMG)", "Invalid size of scatteringMultiplicity");
    scatterXSs.assign(nG, 0.0);
    for (int gi = 0; gi < nG; ++gi) { double s = 0.0; for (int go = 0; go < nG; ++go) s += P0[go + nG * gi]; scatterXSs[gi] = s; }
    if (isP1) {
      P1 = d.getRealArray("P1");
      if ((int)P1.size() != nG * nG) throw FatalError("buildFromDict (multiScatterP1MG)", "Invalid size of P1");
      for (int i = 0; i < nG * nG; ++i) { if (P0[i] != 0.0) P1[i] = P1[i] / P0[i] * 3.0; else P1[i] = 0.0; }
    }
    capture = d.getRealArray("capture");
    if ((int)capture.size() != nG) throw FatalError("init (baseMgNeutronMaterial)", "Capture XSs have wrong size");
    iescatter = scatterXSs;
    if (fissile) {
      nu = d.getRealArray("nu");
      if ((int)nu.size() != nG) throw FatalError("buildFromDict (fissionMG)", "Invalid number of values of nu");
      chi = d.getRealArray("chi");
      if ((int)chi.size() != nG) throw FatalError("buildFromDict (fissionMG)", "Invalid number of values of chi");
      double S = 0.0; for (double c : chi) S += c;
      if (std::fabs(S - 1.0) > 0.01 * FP_REL_TOL) for (double& c : chi) c = c / S;
      fission = d.getRealArray("fission");
      if ((int)fission.size() != nG) throw FatalError("init (baseMgNeutronMaterial)", "Fission XSs have wrong size");
      nuFission.resize(nG); kappaFission.resize(nG);
      std::vector<double> kappa;
      if (d.isPresent("kappa")) {
        kappa = d.getRealArray("kappa");
        if ((int)kappa.size() != nG) throw FatalError("init (baseMgNeutronMaterial)", "Kappa vector has wrong size");
      } else kappa.assign(nG, (double)202.27f);   // KAPPA_DEFAULT = 202.27 is a default-REAL literal (fissionMG_class.f90:64)
      for (int g = 0; g < nG; ++g) { nuFission[g] = nu[g] * fission[g]; kappaFission[g] = kappa[g] * fission[g]; }
    }
    total.resize(nG);
    for (int g = 0; g < nG; ++g) {
      total[g] = iescatter[g] + capture[g];
      if (fissile) total[g] = total[g] + fission[g];
    }
  }

  void getMacroXSs(MacroXSs& x, int G) const {           // G 1-based
    if (G < 1 || G > nG) throw FatalError("getMacroXSs (baseMgNeutronMaterial)", "Invalid group number");
    int g = G - 1;
    x = MacroXSs();
    x.total = total[g]; x.elasticScatter = 0.0; x.inelasticScatter = iescatter[g]; x.capture = capture[g];
    if (fissile) { x.fission = fission[g]; x.nuFission = nuFission[g]; x.kappaXS = kappaFission[g]; }
  }
  int sampleGout(int G_in, RNG& rand) const {
    double rem = rand.get() * scatterXSs[G_in - 1];
    for (int go = 1; go <= nG; ++go) {
      rem = rem - P0[(go - 1) + nG * (G_in - 1)];
      if (rem < 0.0) return go;
    }
    throw FatalError("sampleGout (multiScatterMG)", "Sampling failed. Wrong scatter XS or random number above 1?");
  }
  void scatterSampleOut(double& mu, double& phi, int& G_out, int G_in, RNG& rand) const {
    G_out = sampleGout(G_in, rand);
    if (isP1) mu = sampleLegendreP1(P1[(G_out - 1) + nG * (G_in - 1)], rand);
    else mu = 2.0 * rand.get() - 1.0;
    phi = TWO_PI * rand.get();
  }
  double production(int G_in, int G_out) const { return prod[(G_out - 1) + nG * (G_in - 1)]; }
  void fissionSampleOut(double& mu, double& phi, int& G_out, RNG& rand) const {
    mu = 2.0 * rand.get() - 1.0;
    phi = TWO_PI * rand.get();
    double rem = rand.get();
    for (G_out = 1; G_out <= nG; ++G_out) {
      rem = rem - chi[G_out - 1];
      if (rem < 0.0) return;
    }
    throw FatalError("sampleOut (fissionMG)", "Sampling failed. Unnormalised CHI or rand above 1?!");
  }
};

struct MgDatabase {
  std::vector<MgMaterial> mats;
  std::map<std::string, int> nameMap;      // includes void / outside / overlap
  std::vector<int> activeMats;
  std::vector<double> majorant;
  double collisionXS = 0.0;
  int nG = 0;

  // materialMenu init: names only (needed before geometry is built)
  static std::map<std::string, int> materialMenu(const Dict& nuclearData, std::vector<std::string>* names = nullptr) {
    std::map<std::string, int> m;
    int i = 0;
    for (auto& n : nuclearData.getDict("materials").keys("dict")) { m[n] = ++i; if (names) names->push_back(n); }
    m["void"] = VOID_MAT; m["outside"] = OUTSIDE_MAT; m["overlap"] = OVERLAP_MAT;
    return m;
  }

  void init(const Dict& nuclearData, const std::string& handleName, const std::string& baseDir) {
    const Dict& h = nuclearData.getDict("handles").getDict(handleName);
    if (h.getWord("type") != "baseMgNeutronDatabase") throw FatalError("ndReg", "oracle supports baseMgNeutronDatabase for MG");
    if (h.isPresent("avgDist")) {
      double t = h.getReal("avgDist");
      if (t <= 0.0) throw FatalError("init (baseMgNeutronDatabase)", "Must have a finite, positive minimum average collision distance");
      collisionXS = 1.0 / t;
    }
    std::string scatterKey = h.getWord("PN");
    std::vector<std::string> names;
    nameMap = materialMenu(nuclearData, &names);
    const Dict& md = nuclearData.getDict("materials");
    for (auto& n : names) {
      std::string path = md.getDict(n).getWord("xsFile");
      if (!path.empty() && path[0] != '/') path = baseDir + "/" + path;
      MgMaterial m; m.name = n;
      m.init(Dict::fromFile(path), scatterKey);
      mats.push_back(std::move(m));
    }
    nG = mats.at(0).nG;
    for (auto& m : mats) if (m.nG != nG) throw FatalError("init (baseMgNeutronDatabase)", "Inconsistent # of groups in materials");
  }
  void activate(const std::vector<int>& active) { activeMats = active; initMajorant(); }
  void initMajorant() {
    majorant.assign(nG, 0.0);
    for (int g = 0; g < nG; ++g) {
      double xs = 0.0;
      for (int idx : activeMats) xs = std::max(xs, mats.at(idx - 1).total[g]);
      majorant[g] = xs * 1.0;
    }
  }
  // alpha absorption is zero for eigenvalue calculations (particle_class.f90:503-513, alpha = 0)
  double getTotalMatXS(int G, int matIdx) const {
    if (matIdx < 1 || matIdx > (int)mats.size()) throw FatalError("getTotalMatXS", "Particle is in an undefined material");
    return mats[matIdx - 1].total.at(G - 1) + 0.0;
  }
  double getTrackMatXS(int G, int matIdx) const {
    if (matIdx == VOID_MAT) return 0.0;
    return getTotalMatXS(G, matIdx);
  }
  double getMajorantXS(int G) const {
    if (G < 1 || G > nG) throw FatalError("getMajorantXS", "Invalid group number");
    return majorant[G - 1] + 0.0;
  }
};

}  // namespace orc
