// ORACLE (test infrastructure, not product code).
// CPU restatement of SCONE's continuous-energy reaction kinematics: tabular pdfs, angle and energy laws, release laws,
// elastic / inelastic / fission reactions built from an ACE card.  Each piece cites what it follows.
//
//   NuclearData/NuclearDataStructures/pdf/tabularPdf_class.f90:49-92,190-260          tabularPdf sample / init
//   NuclearData/emissionENDF/angleLawENDF/tabularAngle_class.f90:40-95 ; muEndfPdf/{tabularMu,isotropicMu}_class.f90
//   NuclearData/emissionENDF/angleLawENDF/angleLawENDFfactory_func.f90:25-60
//   NuclearData/emissionENDF/energyLawENDF/energyLawENDFfactory_func.f90:35-150
//   NuclearData/emissionENDF/energyLawENDF/{levelScattering,contTabularEnergy,maxwellSpectrum,evaporationSpectrum,
//                                           multipleEnergyLaws}_class.f90 ; energyEndfPdf/tabularEnergy_class.f90
//   NuclearData/NuclearDataStructures/pdf/maxwellEnergyPdf_class.f90:30-50
//   NuclearData/emissionENDF/releaseLawENDF/releaseLawENDFfactory_func.f90:25-50
//   NuclearData/Reactions/uncorrelatedReactionCE/{elasticNeutronScatter,neutronScatter,fissionCE}_class.f90
#pragma once
#include <limits>

#include "ceace.hpp"
#include "mathmode.hpp"
#include "rng.hpp"

namespace orc_ce {
using orc::RNG;

constexpr double PI = 3.14159265358979323846264338327950288;          // numPrecision.f90: 4*atan(1)
constexpr double TWO_PI = 6.283185307179586476925286766559;
constexpr double SQRT_PI = 1.77245385090551602729816748334115;        // sqrt(PI), correctly rounded
constexpr double MINIMUM_ENERGY = 1.0E-11, MAXIMUM_ENERGY = 20.0;     // universalVariables.f90:28-29
constexpr double shakesPerS = 1.0e+8;
constexpr double HUGE_LAMBDA = std::numeric_limits<double>::max();    // huge(lambda)

// ---- tabularPdf ------------------------------------------------------------------------------------------
struct TabularPdf {
  std::vector<double> x, pdf, cdf; int flag = 2;                      // 1 histogram, 2 lin-lin
  void initCdf(const std::vector<double>& x_, const std::vector<double>& p_, const std::vector<double>& c_, int flag_) {   // :240-275
    if (x_.size() != p_.size() || x_.size() != c_.size()) throw CeError("tabularPdf: sizes differ");
    for (size_t i = 1; i < x_.size(); ++i) if (x_[i] < x_[i - 1]) throw CeError("tabularPdf: x grid is not sorted");
    for (size_t i = 1; i < c_.size(); ++i) if (c_[i] < c_[i - 1]) throw CeError("tabularPdf: CDF is not sorted");
    for (double v : p_) if (v < 0.0) throw CeError("tabularPdf: PDF contains -ve values");
    if (std::fabs(c_.front()) > 1.0e-6) throw CeError("tabularPdf: CDF does not begin with 0");
    if (std::fabs(c_.back() - 1.0) > 1.0e-6) throw CeError("tabularPdf: CDF does not end with 1");
    if (flag_ != 1 && flag_ != 2) throw CeError("tabularPdf: unrecognised interpolation flag");
    x = x_; pdf = p_; cdf = c_; cdf.back() = 1.0; flag = flag_;
  }
  double sample(double r) const {                                     // :49-92
    int idx = linearFloor(cdf, r);
    if (idx <= 0) throw CeError("tabularPdf sample: search failed");
    idx = std::min(idx, (int)x.size() - 1);
    double ci = cdf[idx - 1], pi = pdf[idx - 1];
    if (flag == 1) return x[idx - 1] + (r - ci) / pi;
    double f = (pdf[idx] - pdf[idx - 1]) / (x[idx] - x[idx - 1]);
    if (f == 0.0 || (pi * pi + 2 * f * (r - ci)) < 0.0) return x[idx - 1] + (r - ci) / pi;
    double delta = std::sqrt(pi * pi + 2 * f * (r - ci));
    return x[idx - 1] + (delta - pi) / f;
  }
  double xMin() const { return x.front(); }
  double xMax() const { return x.back(); }
};

// ---- angle laws --------------------------------------------------------------------------------------------
struct AngleLaw {
  enum Kind { ISOTROPIC, TABULAR } kind = ISOTROPIC;
  std::vector<double> eGrid;
  std::vector<int> muKind;                                            // per energy point: 0 isotropicMu, 2 tabularMu
  std::vector<TabularPdf> muPdf;
  // new_angleLawENDF for a reaction with LOCB > 0 / tabularAngle%init: head is at the angular block of the MT
  void initTabular(AceCard& ACE) {
    kind = TABULAR;
    int N = ACE.readInt();
    eGrid = ACE.readReals(N);
    std::vector<int> muLoc = ACE.readInts(N);
    for (size_t i = 1; i < eGrid.size(); ++i) if (eGrid[i] < eGrid[i - 1]) throw CeError("tabularAngle: eGrid is not sorted ascending");
    muKind.assign(N, 0); muPdf.assign(N, TabularPdf());
    for (int i = 0; i < N; ++i) {
      if (muLoc[i] == 0) muKind[i] = 0;
      else if (muLoc[i] > 0) throw CeError("tabularAngle: 32 equiprobable bin mu pdf is not supported by the oracle (the reference indexes boundaries(0))");
      else {
        ACE.head = ACE.JXS[8] + std::abs(muLoc[i]) - 1;              // setToAnglePdf
        int inter = ACE.readInt(); int M = ACE.readInt();
        auto mu = ACE.readReals(M), pdf = ACE.readReals(M), cdf = ACE.readReals(M);
        if (mu.front() != -1.0 || mu.back() != 1.0) throw CeError("tabularMu: mu does not begin with -1 and end with 1");
        muPdf[i].initCdf(mu, pdf, cdf, inter); muKind[i] = 2;
      }
    }
  }
  double sampleMuPdf(int i, RNG& rand) const {                        // muEndfPdf%sample (1-based i)
    if (muKind[i - 1] == 0) return 2.0 * rand.get() - 1.0;
    double r = rand.get();
    return muPdf[i - 1].sample(r);
  }
  double sample(double E, RNG& rand) const {
    if (kind == ISOTROPIC) return 2.0 * rand.get() - 1.0;             // isotropicAngle -> isotropicMu
    int idx = binarySearch(eGrid, E);                                 // tabularAngle_class.f90:62-79
    if (idx <= 0) throw CeError("tabularAngle sample: energy search failed");
    double eps = (E - eGrid[idx - 1]) / (eGrid[idx] - eGrid[idx - 1]);
    double r = rand.get();
    if (r < eps) return sampleMuPdf(idx + 1, rand);
    return sampleMuPdf(idx, rand);
  }
};

// ---- endfTable (Release with LNU = 2 layout is the same object) -------------------------------------------
struct EndfTable {
  std::vector<double> x, y; std::vector<int> bounds, inter;
  void read(AceCard& ACE) {                                           // NR, [bounds, flags], N, x(N), y(N)
    int NR = ACE.readInt();
    if (NR != 0) { bounds = ACE.readInts(NR); inter = ACE.readInts(NR); }
    int N = ACE.readInt(); x = ACE.readReals(N); y = ACE.readReals(N);
  }
  double at(double v) const {                                         // endfTable_class.f90:162-198
    int idx = linearFloor(x, v);
    if (idx < 0) throw CeError("endfTable at: search of grid failed");
    double x0 = x[idx - 1], x1 = x[idx], y0 = y[idx - 1], y1 = y[idx];
    if (bounds.empty()) return interpolate(x0, x1, y0, y1, v);
    if (bounds.size() == 1) return endfInterpolate(x0, x1, y0, y1, v, inter[0]);
    size_t b = 0; while (b + 1 < bounds.size() && bounds[b] < idx + 1) ++b;
    return endfInterpolate(x0, x1, y0, y1, v, inter[b]);
  }
};

// ---- energy laws -------------------------------------------------------------------------------------------
struct EnergyLaw {
  enum Kind { NONE, LEVEL, CONT_TAB, MAXWELL, EVAPORATION, MULTI } kind = NONE;
  // level scattering
  double LDAT1 = 0.0, LDAT2 = 0.0;
  // continuous tabular
  std::vector<double> eGrid; std::vector<TabularPdf> ePdfs; std::vector<int> interBounds, interFlags;
  // Maxwell / evaporation
  EndfTable T_of_E; double U = 0.0;
  // multiple laws
  struct Sub { EndfTable prob; double E_min = 0, E_max = 0; std::shared_ptr<EnergyLaw> law; };
  std::vector<Sub> subs;

  static constexpr int maxIter = 1000;                                // maxwellSpectrum_class.f90 maxIter? (see build)

  double sample(double E_in, RNG& rand) const {
    switch (kind) {
      case NONE: return E_in;                                         // noEnergy%sample
      case LEVEL: return LDAT2 * (E_in - LDAT1);                      // levelScattering_class.f90:35-42
      case CONT_TAB: return sampleContTab(E_in, rand);
      case MAXWELL: {                                                 // maxwellSpectrum_class.f90:40-56 + maxwellEnergyPdf sample_Johnk
        double T = T_of_E.at(E_in);
        for (int i = 0; i < maxIter; ++i) {
          double r1 = rand.get(), r2 = rand.get(), r3 = rand.get();
          double cosine = orc::mcos(0.5 * PI * r1);
          double beta = cosine * cosine;
          double gamma05 = -orc::mlog(r2) * beta;
          double E_out = (-orc::mlog(r3) + gamma05) * T;
          if (E_out < E_in - U) return E_out;
        }
        throw CeError("maxwellSpectrum: sampling failed to be accepted after maxIter iterations");
      }
      case EVAPORATION: {                                             // evaporationSpectrum_class.f90:36-52
        double T = T_of_E.at(E_in);
        for (;;) {
          double r1 = rand.get(), r2 = rand.get();
          double E_out = -T * orc::mlog(r1 * r2);
          if (E_out <= E_in - U) return E_out;
        }
      }
      case MULTI: {                                                   // multipleEnergyLaws_class.f90:52-75
        double r = rand.get();
        for (auto& s : subs) {
          double E = E_in;
          E = std::max(E, s.E_min);
          E = std::min(E, s.E_max);
          double prob = s.prob.at(E);
          if (r < prob) return s.law->sample(E_in, rand);
          r = r - prob;
        }
        throw CeError("multipleEnergyLaws: failed to sample an energy law");
      }
    }
    return 0.0;
  }
  double sampleContTab(double E_in, RNG& rand) const {                // contTabularEnergy_class.f90:40-90
    int idx = binarySearch(eGrid, E_in);
    if (idx <= 0) throw CeError("contTabularEnergy sample: energy search failed");
    int flag = 2;
    if (!interBounds.empty()) {
      int ii = -1;
      for (size_t k = 0; k < interBounds.size(); ++k) if (interBounds[k] >= idx) { ii = (int)k; break; }   // linearCeilingIdxOpen
      if (ii < 0) throw CeError("contTabularEnergy: failed interpolation region search");
      flag = interFlags[ii];
    }
    if (flag == 2) {
      double E_min_low = ePdfs[idx - 1].xMin(), E_max_low = ePdfs[idx - 1].xMax();
      double E_min_up = ePdfs[idx].xMin(), E_max_up = ePdfs[idx].xMax();
      double eps = (E_in - eGrid[idx - 1]) / (eGrid[idx] - eGrid[idx - 1]);
      double E_min = E_min_low * (1.0 - eps) + eps * E_min_up;
      double E_max = E_max_low * (1.0 - eps) + eps * E_max_up;
      double r = rand.get();
      double E_out, factor;
      if (r < eps) {
        double rr = rand.get();
        E_out = ePdfs[idx].sample(rr);
        factor = (E_out - E_min_up) / (E_max_up - E_min_up);
      } else {
        double rr = rand.get();
        E_out = ePdfs[idx - 1].sample(rr);
        factor = (E_out - E_min_low) / (E_max_low - E_min_low);
      }
      return E_min * (1.0 - factor) + factor * E_max;
    } else if (flag == 1) {
      double rr = rand.get();
      return ePdfs[idx - 1].sample(rr);
    }
    throw CeError("contTabularEnergy: unsupported interpolation flag");
  }

  // buildENDFLaw: head is set relative to root by the caller (energyLawENDFfactory_func.f90:118-150)
  void buildLaw(int LAW, int root, int offset, AceCard& ACE) {
    ACE.head = root + offset - 1;                                     // setRelativeTo
    switch (LAW) {
      case 4: {                                                       // contTabularEnergy from ACE (:140-175)
        kind = CONT_TAB;
        int NR = ACE.readInt();
        if (NR < 0) throw CeError("contTabularEnergy: -ve number of interpolation regions");
        auto b = ACE.readInts(NR), fl = ACE.readInts(NR);
        int N = ACE.readInt();
        eGrid = ACE.readReals(N);
        auto loc = ACE.readInts(N);
        ePdfs.assign(N, TabularPdf());
        for (int i = 0; i < N; ++i) {
          ACE.head = root + loc[i] - 1;
          int INTT = ACE.readInt();                                   // tabularEnergy init_fromACE
          if (INTT > 10) throw CeError("tabularEnergy: INTT > 10, discrete photon lines are not implemented");
          int NP = ACE.readInt();
          auto e = ACE.readReals(NP), p = ACE.readReals(NP), c = ACE.readReals(NP);
          for (double v : e) if (v < 0.0) throw CeError("tabularEnergy: E contains -ve values");
          ePdfs[i].initCdf(e, p, c, INTT);
        }
        if (NR > 0) {
          if (b.back() != N) throw CeError("contTabularEnergy: incomplete interpolation scheme");
          interBounds = b; interFlags = fl;
        }
        break;
      }
      case 7: case 9: {                                               // maxwellSpectrum / evaporationSpectrum from ACE
        kind = (LAW == 7) ? MAXWELL : EVAPORATION;
        T_of_E.read(ACE);
        U = ACE.readReal();
        break;
      }
      case 3: {                                                       // levelScattering from ACE
        kind = LEVEL;
        LDAT1 = ACE.readReal(); LDAT2 = ACE.readReal();
        if (LDAT2 < 0.0) throw CeError("levelScattering: LDAT2 is -ve");
        if (LDAT2 >= 1.0) throw CeError("levelScattering: LDAT2 is >= 1.0");
        break;
      }
      default: throw CeError("Energy law type is not recognised or yet supported: " + std::to_string(LAW));
    }
  }
  // new_energyLawENDF (energyLawENDFfactory_func.f90:35-116)
  void build(AceCard& ACE, int MT, bool delayed) {
    int root, LOCC;
    if (delayed) { root = ACE.JXS[26]; LOCC = AceCard::r2i(ACE.xss(ACE.JXS[25] + MT - 1)); }
    else {
      if (MT == N_N_ELASTIC || ACE.rec(MT).isCapture) { kind = NONE; return; }
      root = ACE.JXS[10]; LOCC = ACE.LOCC(MT);
    }
    ACE.head = root + LOCC - 1;
    int LNW = ACE.readInt();
    if (LNW == 0) {
      int LAW = ACE.readInt(); int loc = ACE.readInt();
      buildLaw(LAW, root, loc, ACE);
      return;
    }
    int N = 1;
    while (LNW != 0) { N += 1; ACE.head = root + LNW - 1; LNW = ACE.readInt(); if (N > 100) throw CeError("new_energyLawENDF: infinite loop"); }
    kind = MULTI; subs.clear();
    ACE.head = root + LOCC - 1;
    LNW = ACE.readInt();
    for (int i = 0; i < N; ++i) {
      int LAW = ACE.readInt(); int loc = ACE.readInt();
      Sub s;
      s.prob.read(ACE);
      s.law = std::make_shared<EnergyLaw>();
      s.law->buildLaw(LAW, root, loc, ACE);
      for (double v : s.prob.x) if (v < 0.0) throw CeError("multipleEnergyLaws: -ve entries in eGrid");
      for (double v : s.prob.y) if (v < 0.0) throw CeError("multipleEnergyLaws: -ve entries in pdf");
      s.E_min = s.prob.x.front(); s.E_max = s.prob.x.back();
      subs.push_back(s);
      if (LNW != 0) { ACE.head = root + LNW - 1; LNW = ACE.readInt(); }
    }
    if (LNW != 0) throw CeError("new_energyLawENDF: LNW is not 0 after reading all energy laws");
  }
};

// ---- reactions ----------------------------------------------------------------------------------------------
// uncorrelatedReactionCE interface: sampleOut(mu, phi, E_out, E_in, rand[, lambda]), release(E), inCMframe()
struct ElasticScatter {                                               // elasticNeutronScatter_class.f90
  AngleLaw angle;
  void init(AceCard& ACE) {
    if (ACE.LOCB(N_N_ELASTIC) == 0) angle.kind = AngleLaw::ISOTROPIC;
    else { ACE.head = ACE.JXS[8]; angle.initTabular(ACE); }           // setToAngleEscatter
  }
  void sampleOut(double& mu, double& phi, double& E_out, double E_in, RNG& rand) const {   // :120-139
    E_out = E_in;
    mu = angle.sample(E_in, rand);
    phi = rand.get() * TWO_PI;
  }
};

struct NeutronScatter {                                               // neutronScatter_class.f90
  int MT = 0; bool cmFrame = true;
  int TY = 1; bool tabRelease = false; EndfTable releaseTab;          // constantRelease(TY) or tabularRelease
  AngleLaw mu; EnergyLaw e;
  void init(AceCard& ACE, int MT_) {                                  // buildFromACE :205-240
    MT = MT_;
    const auto& m = ACE.rec(MT);
    if (m.isCapture) throw CeError("neutronScatter: reaction does not produce 2nd-ary neutrons");
    int LOCB = ACE.LOCB(MT);
    cmFrame = m.CMframe;
    TY = m.TY;
    if (TY == 19) throw CeError("neutronScatter: reaction is fission");
    if (TY > 100) { tabRelease = true; ACE.head = ACE.JXS[10] + TY - 101; releaseTab.read(ACE); }   // setToReleaseMT + tabularRelease
    if (LOCB == -1) throw CeError("neutronScatter: correlated angle-energy laws (Kalbach-87, law 61, N-body) are not supported by the oracle");
    if (LOCB == 0) mu.kind = AngleLaw::ISOTROPIC;
    else { ACE.head = ACE.JXS[8] + LOCB - 1; mu.initTabular(ACE); }   // setToAngleMT: ANDp = JXS(9) + LOCB - 1
    e.build(ACE, MT, false);
  }
  double release(double E) const { return tabRelease ? releaseTab.at(E) : (double)TY; }
  void sampleOut(double& mu_, double& phi, double& E_out, double E_in, RNG& rand) const {   // :150-170
    mu_ = mu.sample(E_in, rand);
    E_out = e.sample(E_in, rand);
    E_out = std::max(E_out, MINIMUM_ENERGY);
    phi = rand.get() * TWO_PI;
  }
};

struct FissionCE {                                                    // fissionCE_class.f90
  Release nuTotal, nuDelayed; EnergyLaw eLawPrompt; double Q = 0.0;
  struct Precursor { double lambda = 0.0; EndfTable prob; EnergyLaw eLaw; };
  std::vector<Precursor> delayed;
  void init(AceCard& ACE, int MT) {                                   // buildFromACE :283-345
    bool onlyOneNu = ACE.totalNUp != 0 && ACE.promptNUp == 0, withDelayed = ACE.delayNUp != 0;
    if (withDelayed && onlyOneNu) throw CeError("Prompt/Total Nu is given with delayed data. Which one is which?");
    if (ACE.totalNUp == 0 && withDelayed) throw CeError("Has delayed neutron data but does not have total NuBar");
    nuTotal = ACE.readNu(ACE.totalNUp);
    eLawPrompt.build(ACE, MT, false);
    Q = ACE.rec(MT).Q;
    if (withDelayed) {
      nuDelayed = ACE.readNu(ACE.delayNUp);
      int NP = ACE.NXS[7];
      if (NP == 0) throw CeError("Has delayed neutrons but not precursors");
      delayed.assign(NP, Precursor());
      if (ACE.JXS[24] == 0) throw CeError("Missing fission data: cannot locate precursor pdf, JXS(25) == 0");
      ACE.head = ACE.JXS[24];                                         // setToPrecursors
      for (int i = 0; i < NP; ++i) {
        delayed[i].lambda = ACE.readReal() * shakesPerS;
        delayed[i].prob.read(ACE);
      }
      for (int i = 0; i < NP; ++i) delayed[i].eLaw.build(ACE, i + 1, true);
    }
  }
  double release(double E) const { return nuTotal.at(E); }
  double releaseDelayed(double E) const { if (!nuDelayed.present) return 0.0; if (!nuDelayed.hasEnergy(E)) return 0.0; return nuDelayed.at(E); }
  double releasePrompt(double E) const { return release(E) - releaseDelayed(E); }
  void sampleOut(double& mu, double& phi, double& E_out, double E_in, RNG& rand, double& lambda) const {   // :187-235
    mu = 2.0 * rand.get() - 1.0;
    phi = TWO_PI * rand.get();
    double p_del = delayed.empty() ? 0.0 : releaseDelayed(E_in) / release(E_in);
    double r1 = rand.get();
    if (r1 > p_del) { E_out = eLawPrompt.sample(E_in, rand); lambda = HUGE_LAMBDA; return; }
    double r2 = rand.get();
    for (auto& d : delayed) {
      r2 = r2 - d.prob.at(E_in);
      if (r2 < 0.0) { E_out = d.eLaw.sample(E_in, rand); lambda = d.lambda; return; }
    }
    E_out = delayed.back().eLaw.sample(E_in, rand);
    lambda = delayed.back().lambda;
  }
};

}  // namespace orc_ce
