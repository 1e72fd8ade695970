// ORACLE (test infrastructure, not product code).
// CPU restatement of SCONE's continuous-energy history loop: aceNeutronDatabase as the eigenvalue driver uses it,
// ceNeutronMaterial nuclide sampling, the neutronCEstd collision processor with the scattering kernels, the CE branch
// of fissionSource, and the eigenPhysicsPackage set-up for `dataType ce`.  Transport (DT / ST / HT), tallies, dungeon
// and the cycle driver are the ones of physics.hpp (they are data-type agnostic in the reference too).
//
//   NuclearData/ceNeutronData/aceDatabase/aceNeutronDatabase_class.f90:560-760   init (materials, kT, fissile, eBounds)
//   NuclearData/ceNeutronData/ceNeutronDatabase_inter.f90:120-230                 getTrackMatXS / getTotalMatXS / getMajorantXS
//   NuclearData/ceNeutronData/ceNeutronMaterial_class.f90:258-276,338-455         sampleNuclide, sampleFission
//   NuclearData/ceNeutronData/aceLibrary_mod.f90                                   library file: NAME; LINE; PATH;
//   NuclearData/materialMenu_mod.f90 init_materialItem                             temp, composition
//   CollisionOperator/CollisionProcessors/neutronCEstd_class.f90:157-588          sampleCollision, implicit, elastic, inelastic, cutoffs
//   CollisionOperator/scatteringKernels_func.f90:38-378                            asymptotic + free-gas kernels
//   ParticleObjects/Source/fissionSource_class.f90:211-235                         CE source particle
//
// Not restated (absent from the bundled data, refused at load time): S(a,b) (`moder`), URR tables (`ures 1`), TMS (`tms 1`),
// DBRC, correlated angle-energy laws.
#pragma once
#include <fstream>
#include <sstream>

#include "cedata.hpp"
#include "physics.hpp"

namespace orc {

constexpr double kBoltzmannMeV = 1.380649e-23 / 1.60218e-13;          // universalVariables.f90: kBoltzmann / joulesPerMeV

// binary form of an ACE card (data/ace/*.acebin, written by tests/golden/make_ace_fixtures.py): the GPU box has no
// /root/reference, so the card arrays travel as a fixture.  Layout: "SBACE1\0\0", ZAID[16], AW, TZ, NXS[16] i32, JXS[32] i32, n i64, XSS[n]
inline void readAceBin(orc_ce::AceCard& ace, const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw orc_ce::CeError("Cannot open ACE fixture: " + path);
  char magic[8], zaid[16]; double aw, tz; int nxs[16], jxs[32]; long long n;
  f.read(magic, 8); f.read(zaid, 16); f.read((char*)&aw, 8); f.read((char*)&tz, 8); f.read((char*)nxs, 64); f.read((char*)jxs, 128); f.read((char*)&n, 8);
  if (!f || std::string(magic, 6) != "SBACE1") throw orc_ce::CeError("Not an ACE fixture: " + path);
  std::vector<double> xss((size_t)n);
  f.read((char*)xss.data(), 8 * n);
  if (!f) throw orc_ce::CeError("ACE fixture is truncated: " + path);
  ace.fromArrays(std::string(zaid, strnlen(zaid, 16)), aw, tz, nxs, jxs, xss.data(), (long)n);
}

struct CeData : XsView {
  orc_ce::Database db;
  std::vector<double> matkT; std::vector<char> matFissile;
  std::map<std::string, int> nameMap;
  double collXS = 0.0; bool hasMajorant = true;
  std::vector<std::string> nuclideNames;                              // library names, database order

  struct LibEntry { std::string path; int line; };
  static std::map<std::string, LibEntry> loadLibrary(const std::string& path) {      // aceLibrary_mod.f90 aceLib_load
    std::ifstream f(path);
    if (!f) throw FatalError("aceLib_load", "Cannot open ACE library file: " + path);
    std::string dir = path.substr(0, path.find_last_of('/') == std::string::npos ? 0 : path.find_last_of('/'));
    std::map<std::string, LibEntry> lib;
    std::string line;
    while (std::getline(f, line)) {
      size_t c = line.find('!'); if (c != std::string::npos) line = line.substr(0, c);
      std::vector<std::string> fld; std::stringstream ss(line); std::string t;
      while (std::getline(ss, t, ';')) { size_t a = t.find_first_not_of(" \t\r"), b = t.find_last_not_of(" \t\r"); if (a != std::string::npos) fld.push_back(t.substr(a, b - a + 1)); }
      if (fld.empty()) continue;
      if (fld.size() < 3) throw FatalError("aceLib_load", "Ill-formatted line in the ACE library: " + line);
      LibEntry e; e.line = std::stoi(fld[1]); e.path = fld[2];
      if (e.path[0] != '/') e.path = dir.empty() ? e.path : dir + "/" + e.path;        // relative to the library file (fixtures)
      lib[fld[0]] = e;
    }
    return lib;
  }

  void init(const Dict& nuclearData, const std::string& handleName, const std::string& baseDir) {
    const Dict& h = nuclearData.getDict("handles").getDict(handleName);
    if (h.getWord("type") != "aceNeutronDatabase") throw FatalError("ndReg", "oracle supports aceNeutronDatabase for CE");
    if (h.isPresent("avgDist")) {
      double t = h.getReal("avgDist");
      if (t <= 0.0) throw FatalError("init (aceNeutronDatabase)", "Must have a finite, positive minimum average collision distance");
      collXS = 1.0 / t;
    }
    if (h.getBool("ures", false)) throw FatalError("init (aceNeutronDatabase)", "oracle: URR probability tables are not supported (no data in the reference checkout)");
    if (h.isPresent("DBRC")) throw FatalError("init (aceNeutronDatabase)", "oracle: DBRC is not supported");
    hasMajorant = h.getBool("majorant", true);
    std::string libPath = h.getWord("aceLibrary");
    if (!libPath.empty() && libPath[0] != '/') libPath = baseDir + "/" + libPath;
    auto lib = loadLibrary(libPath);
    std::vector<std::string> names;
    nameMap = MgDatabase::materialMenu(nuclearData, &names);
    const Dict& md = nuclearData.getDict("materials");
    std::map<std::string, int> nucIdx;
    for (auto& n : names) {
      const Dict& m = md.getDict(n);
      if (m.getBool("tms", false)) throw FatalError("init_materialItem", "oracle: TMS is not supported");
      if (m.isPresent("moder")) throw FatalError("init_materialItem", "oracle: S(alpha,beta) data are not supported (no data in the reference checkout)");
      double T = m.getReal("temp", 0.0);
      if (T < 0.0) throw FatalError("init_materialItem", "The temperature of material is negative");
      const Dict& comp = m.getDict("composition");
      orc_ce::Material mat; bool fiss = false;
      for (auto& key : comp.keys("all")) {
        auto it = nucIdx.find(key);
        if (it == nucIdx.end()) {
          auto le = lib.find(key + "c");
          if (le == lib.end()) le = lib.find(key);
          if (le == lib.end()) throw FatalError("new_neutronACE", "Nuclide " + key + " was not found in the ACE library");
          orc_ce::AceCard ace;
          const std::string& p = le->second.path;
          if (p.size() > 7 && p.substr(p.size() - 7) == ".acebin") readAceBin(ace, p); else ace.readFromFile(p, le->second.line);
          orc_ce::Nuclide nuc; nuc.init(ace, true);
          db.nuclides.push_back(std::move(nuc));
          nuclideNames.push_back(key);
          it = nucIdx.emplace(key, (int)db.nuclides.size()).first;
        }
        mat.nuclides.push_back(it->second);
        mat.dens.push_back(comp.getReal(key));
        if (mat.dens.back() < 0.0) throw FatalError("setComposition", "-ve nuclide densities are present");
        fiss = fiss || db.nuclides[it->second - 1].fissile;
      }
      if (mat.nuclides.empty()) throw FatalError("setComposition", "Empty composition is not allowed");
      db.materials.push_back(mat); matkT.push_back(kBoltzmannMeV * T); matFissile.push_back(fiss ? 1 : 0);
    }
    db.finalise();
  }
  void activate(const std::vector<int>& active) { db.activeMat = active; if (hasMajorant) db.initMajorant(); }

  // XsView
  int nMat() const override { return (int)db.materials.size(); }
  double totalMatXS(const Particle& p, int matIdx) const override {
    if (matIdx == VOID_MAT) return 0.0;
    if (matIdx < 1 || matIdx > nMat()) throw FatalError("getTotalMatXS", "Particle is in an undefined material");
    return db.totalMatXS(p.E, matIdx) + 0.0;
  }
  double trackMatXS(const Particle& p, int matIdx) const override { return totalMatXS(p, matIdx); }   // no TMS: trackXS = total
  double majorantXS(const Particle& p) const override {
    if (hasMajorant) return db.majorantXS(p.E) + 0.0;
    double maj = 0.0;
    for (int m : db.activeMat) maj = std::max(maj, db.totalMatXS(p.E, m));
    return maj + 0.0;
  }
  double collisionXS() const override { return collXS; }
  bool isFissileMat(int matIdx) const override { return matFissile.at(matIdx - 1) != 0; }
  void macroXSs(MacroXSs& x, const Particle& p, int matIdx) const override {
    double o[8]; db.macroXSs(o, p.E, matIdx);
    x.total = o[0]; x.elasticScatter = o[1]; x.inelasticScatter = o[2]; x.capture = o[3]; x.fission = o[4]; x.nuFission = o[5]; x.kappaXS = o[6]; x.promptNuFission = o[7];
  }

  // ceNeutronMaterial%sampleNuclide (:338-395), no TMS
  int sampleNuclide(double E, int matIdx, RNG& rand) const {
    const auto& m = db.materials.at(matIdx - 1);
    double trackMatXS = db.totalMatXS(E, matIdx) * rand.get();
    for (size_t i = 0; i < m.nuclides.size(); ++i) {
      const auto& n = db.nuclides[m.nuclides[i] - 1];
      int idx; double f; n.search(idx, f, E);
      double totNucXS = n.totalXS(idx, f);
      trackMatXS = trackMatXS - totNucXS * (m.dens[i] * 1.0);
      if (trackMatXS < 0.0) return m.nuclides[i];
    }
    throw FatalError("sampleNuclide", "Nuclide sampling loop failed to terminate");
  }
  // ceNeutronMaterial%sampleFission (:397-455), no TMS
  int sampleFission(double E, int matIdx, RNG& rand) const {
    if (!isFissileMat(matIdx)) return 0;
    const auto& m = db.materials.at(matIdx - 1);
    double o[8]; db.macroXSs(o, E, matIdx);
    double xs = o[5] * rand.get();
    for (size_t i = 0; i < m.nuclides.size(); ++i) {
      const auto& n = db.nuclides[m.nuclides[i] - 1];
      int idx; double f; n.search(idx, f, E);
      double mic[8]; n.microXSs(mic, idx, f);
      xs = xs - mic[5] * m.dens[i] * 1.0 * 1.0;
      if (xs < 0.0) return m.nuclides[i];
    }
    throw FatalError("sampleFission", "Nuclide sampling loop failed to terminate");
  }
};

// scatteringKernels_func.f90
namespace kernels {
inline void asymptoticScatter(double& E, double& mu, double A) {      // :38-57
  double E_in = E, inv_Ap1 = 1.0 / (A + 1.0);
  E = (1.0 + A * A + 2 * A * mu) * E_in * inv_Ap1 * inv_Ap1;
  mu = (A * mu + 1) * std::sqrt(E_in / E) * inv_Ap1;
  if (mu > 1.0) mu = 1.0;
}
inline void asymptoticInelasticScatter(double& E, double& mu, double E_out, double A) {   // :59-80
  double E_in = E, inv_Ap1 = 1.0 / (A + 1.0);
  E = E_out + (E_in + 2.0 * mu * (A + 1.0) * std::sqrt(E_in * E_out)) * inv_Ap1 * inv_Ap1;
  mu = mu * std::sqrt(E_out / E) + std::sqrt(E_in / E) * inv_Ap1;
  if (mu > 1.0) mu = 1.0;
}
inline double sample_x2expx2(RNG& rand) {                             // :290-310
  double r1 = rand.get(), r2 = rand.get(), r3 = rand.get();
  double cosine = mcos(0.5 * orc_ce::PI * r1);
  double beta = cosine * cosine;
  double gamma05 = -mlog(r2) * beta;
  double sample = -mlog(r3) + gamma05;
  return std::sqrt(sample);
}
inline double sample_x3expx2(RNG& rand) {                             // :316-330
  double r1 = rand.get(), r2 = rand.get();
  double sample = -mlog(r1) - mlog(r2);
  return std::sqrt(sample);
}
inline void sample_targetVelocity(double& X, bool& accept, double& rel_v, double& mu, RNG& rand, double Y, double alpha) {   // :340-378
  double r1 = rand.get(), r2 = rand.get(), r3 = rand.get();
  if (r1 > alpha) X = sample_x2expx2(rand); else X = sample_x3expx2(rand);
  mu = 2.0 * r2 - 1.0;
  rel_v = std::sqrt(Y * Y + X * X - 2.0 * X * Y * mu);
  double P_acc = rel_v / (Y + X);
  accept = P_acc > r3;
}
inline Vec3 targetVelocity_constXS(double E, const Vec3& dir, double A, double kT, RNG& rand) {   // :82-125
  double Y = std::sqrt(A * E / kT);
  double alpha = 2.0 / (Y * orc_ce::SQRT_PI + 2.0);
  double X, rel_v, mu; bool accept;
  for (;;) { sample_targetVelocity(X, accept, rel_v, mu, rand, Y, alpha); if (accept) break; }
  double r1 = rand.get();
  double phi = 2.0 * orc_ce::PI * r1;
  Vec3 V_t = rotateVector(dir, mu, phi);
  double s = X * std::sqrt(kT / A);
  for (int k = 0; k < 3; ++k) V_t[k] = V_t[k] * s;
  return V_t;
}
}  // namespace kernels

struct CeEigenPP : EigenPP {
  CeData ce;
  double minE = orc_ce::MINIMUM_ENERGY, maxE = orc_ce::MAXIMUM_ENERGY, threshE = 400.0, threshA = 1.0;

  void init(const Dict& dict, const std::string& baseDir) override {   // eigenPhysicsPackage_class.f90:417-645 with dataType ce
    initCycles(dict);
    std::string nucData = dict.getWord("XSdata");
    if (dict.getWord("dataType") != "ce") throw FatalError("init (eigenPhysicsPackage)", "oracle CE driver: dataType must be 'ce'");
    if (!dict.isPresent("seed")) throw FatalError("init (eigenPhysicsPackage)", "oracle requires an explicit seed");
    pRNG.init((int64_t)dict.getInt("seed"));
    keff_0 = dict.getReal("keff_0", 1.0);
    const Dict& nd = dict.getDict("nuclearData");
    auto mats = MgDatabase::materialMenu(nd);
    geom.init(dict.getDict("geometry"), mats);
    ce.init(nd, nucData, baseDir);
    ce.activate(geom.activeMats());
    xs = &ce;
    const Dict& co = dict.getDict("collisionOperator");
    if (!co.isPresent("neutronCE") || co.getDict("neutronCE").getWord("type") != "neutronCEstd")
      throw FatalError("collisionOperator init", "oracle supports neutronCEstd only for CE");
    {                                                                 // neutronCEstd init (:110-150)
      const Dict& c = co.getDict("neutronCE");
      minE = c.getReal("minEnergy", orc_ce::MINIMUM_ENERGY); maxE = c.getReal("maxEnergy", orc_ce::MAXIMUM_ENERGY);
      threshE = c.getReal("energyThreshold", 400.0); threshA = c.getReal("massThreshold", 1.0);
      if (c.getBool("makePrec", false)) throw FatalError("init (neutronCEstd)", "oracle: precursors are not supported in eigenvalue mode");
      if (c.getBool("neglectDelayed", false)) throw FatalError("init (neutronCEstd)", "oracle: neglectDelayed is not supported");
      if (minE < 0.0 || maxE < 0.0 || minE >= maxE || threshE < 0 || threshA < 0) throw FatalError("init (neutronCEstd)", "invalid settings");
    }
    const Dict& to = dict.getDict("transportOperator");
    std::string tt = to.getWord("type");
    if (tt == "transportOperatorDT") tracking = TRACK_DT;
    else if (tt == "transportOperatorST") { tracking = TRACK_ST; stCache = to.getBool("cache", true); }
    else if (tt == "transportOperatorHT") { tracking = TRACK_HT; htCutoff = to.getReal("cutoff", 0.9); stCache = to.getBool("cache", true); }
    else throw FatalError("new_transportOperator", "Unrecognised type of transportOperator: " + tt);
    if (fixedSource) {
      activeTally.init(dict.getDict("tally"), mats);
      inactiveTally.init(Dict::fromString(""), mats);
      sourceMats = mats;
      initSource(dict.getDict("source"), 0);
      return;
    }
    inactiveTally.init(dict.getDict("inactiveTally"), mats);
    activeTally.init(dict.getDict("activeTally"), mats);
    if (dict.isPresent("source")) throw FatalError("init (eigenPhysicsPackage)", "oracle supports the default fissionSource only");
    source.init(&geom, nullptr, xs);
    source.sampleCE = [this](ParticleState& p, int matIdx, RNG& rand) {      // fissionSource_class.f90:211-235
      int nucIdx = ce.sampleFission(source.E, matIdx, rand);
      double mu, phi, E_out, lambda;
      ce.db.nuclides.at(nucIdx - 1).fission.sampleOut(mu, phi, E_out, source.E, rand, lambda);
      p.E = E_out; p.isMG = false;
      Vec3 ex; ex[0] = 1.0;
      p.dir = rotateVector(ex, mu, phi);
      if (p.E > ce.db.eBounds[1]) p.E = ce.db.eBounds[1];
    };
    inactiveAtch.init(Dict::fromString("keff { type keffAnalogClerk; } display (keff); mpiSync 1;"), mats);
    activeAtch.init(Dict::fromString("keff { type keffImplicitClerk; } display (keff); mpiSync 1;"), mats);
    inactiveTally.atch = &inactiveAtch;
    activeTally.atch = &activeAtch;
  }

  // collisionProcessor_inter.f90:114-195 with the neutronCEstd hooks
  void collide(Particle& p, TallyAdmin& tally, double trackXS, Dungeon& next) const override {
    using namespace orc_ce;
    const int matIdx = p.matIdx();
    // ---- sampleCollision (:157-215)
    double denom = ce.trackMatXS(p, matIdx);
    double probAlpha = 0.0 / denom;
    int MT;
    int nucIdx = 0; double collE = p.E;
    if (p.pRNG->get() < probAlpha) MT = 0;                            // unreachable for alpha = 0
    else {
      nucIdx = ce.sampleNuclide(p.E, matIdx, *p.pRNG);
      collE = p.E;
      const Nuclide& nuc = ce.db.nuclides[nucIdx - 1];
      int idx; double f; nuc.search(idx, f, collE);
      double mic[8]; nuc.microXSs(mic, idx, f);
      double r = p.pRNG->get();
      int C = 1;                                                      // neutronMicroXSs%invert
      double xs = mic[0] * r - mic[1];
      if (xs > 0.0) C += 1;
      xs = xs - mic[2];
      if (xs > 0.0) C += 1;
      xs = xs - mic[3];
      if (xs > 0.0) C += 1;
      MT = (C == 1) ? N_N_ELASTIC : (C == 2) ? N_N_INELASTIC : (C == 3) ? 101 : N_FISSION;
    }
    tally.reportInColl(p, *xs, trackXS, false);
    p.preCollision = p.state();
    const Nuclide& nuc = ce.db.nuclides.at(nucIdx - 1);
    // ---- implicit (:217-300)
    if (nuc.fissile) {
      double wgt = p.w, w0 = p.preHistory.wgt, k_eff = p.k_eff;
      double rand1 = p.pRNG->get();
      int idx; double f; nuc.search(idx, f, collE);
      double mic[8]; nuc.microXSs(mic, idx, f);
      double sig_nufiss = mic[5], sig_tot = mic[0];
      int n = (int)(std::fabs((wgt * sig_nufiss) / (w0 * sig_tot * k_eff)) + rand1);
      if (n >= 1) {
        wgt = fsign(w0, wgt);
        Vec3 r = p.coords.lvl[0].r;
        for (int i = 0; i < n; ++i) {
          double mu, phi, E_out, lambda;
          nuc.fission.sampleOut(mu, phi, E_out, p.E, *p.pRNG, lambda);
          double wD = 1.0;
          Vec3 dir = rotateVector(p.coords.lvl[0].dir, mu, phi);
          if (E_out > maxE) E_out = maxE;
          ParticleState t = p.state();
          t.r = r; t.dir = dir; t.E = E_out; t.wgt = wgt * wD; t.collisionN = 0;
          next.detain(t);
        }
      }
    }
    // ---- channel
    double muL = 1.0;
    switch (MT) {
      case N_N_ELASTIC: {                                             // elastic (:330-375)
        double A = nuc.mass, kT = nuc.kT;                             // no TMS, no per-particle temperature: nuc%getkT()
        bool isFixed = (p.E > kT * threshE) && (A > threshA);
        if (isFixed) scatterFromFixed(p, nuc.elastic, N_N_ELASTIC, A, muL);
        else scatterFromMoving(p, nuc.elastic, A, kT, muL);
        break;
      }
      case N_N_INELASTIC: {                                           // inelastic (:377-405)
        int which = -1;
        int MTi = nuc.invertInelastic(collE, *p.pRNG, &which);
        const NeutronScatter& reac = nuc.mtData[which].kin;
        double E_before = p.E;
        if (reac.cmFrame) scatterFromFixedMT(p, reac, MTi, nuc.mass, muL);
        else scatterInLAB(p, reac, muL);
        (void)E_before;
        p.w = p.w * reac.release(p.E);                                // release at the post-collision energy, as the reference
        MT = MTi;
        break;
      }
      case 101: case N_FISSION: p.isDead = true; break;               // capture / fission
      default: throw FatalError("collide", "Unsupported MT number");
    }
    if (p.E < minE) p.isDead = true;                                  // cutoffs (:407-417)
    p.collisionN += 1;
    // tally%reportOutColl: keffImplicitClerk scores the (n,xn) multiplicities by MT (keffImplicitClerk_class.f90:245-270)
    tally.reportOutColl(p, MT);
    if (p.isDead) { p.fate = ABS_FATE; tally.reportHist(p); }
  }

  // scatterFromFixed for elastic scattering (:447-480)
  void scatterFromFixed(Particle& p, const orc_ce::ElasticScatter& reac, int, double A, double& muL) const {
    double mu, phi, E_outCM;
    reac.sampleOut(mu, phi, E_outCM, p.E, *p.pRNG);
    double E_out = p.E;
    kernels::asymptoticScatter(E_out, mu, A);
    p.coords.rotate(mu, phi);
    p.E = E_out;
    muL = mu;
  }
  void scatterFromFixedMT(Particle& p, const orc_ce::NeutronScatter& reac, int MT, double A, double& muL) const {
    double mu, phi, E_outCM;
    reac.sampleOut(mu, phi, E_outCM, p.E, *p.pRNG);
    double E_out = p.E;
    if (MT == orc_ce::N_N_ELASTIC) kernels::asymptoticScatter(E_out, mu, A);
    else kernels::asymptoticInelasticScatter(E_out, mu, E_outCM, A);
    p.coords.rotate(mu, phi);
    p.E = E_out;
    muL = mu;
  }
  void scatterInLAB(Particle& p, const orc_ce::NeutronScatter& reac, double& muL) const {   // (:419-445)
    double mu, phi, E_out;
    reac.sampleOut(mu, phi, E_out, p.E, *p.pRNG);
    p.E = E_out;
    p.coords.rotate(mu, phi);
    muL = mu;
  }
  // scatterFromMoving (:482-588), constant cross-section free gas (no DBRC)
  void scatterFromMoving(Particle& p, const orc_ce::ElasticScatter& reac, double A, double kT, double& muL) const {
    Vec3 dir_pre = p.coords.lvl[0].dir;
    double sqE = std::sqrt(p.E);
    Vec3 V_n; for (int k = 0; k < 3; ++k) V_n[k] = dir_pre[k] * sqE;
    Vec3 V_t = kernels::targetVelocity_constXS(p.E, dir_pre, A, kT, *p.pRNG);
    Vec3 V_cm; for (int k = 0; k < 3; ++k) V_cm[k] = (V_n[k] + V_t[k] * A) / (A + 1);
    for (int k = 0; k < 3; ++k) V_n[k] = V_n[k] - V_cm[k];
    double U_n = norm2(V_n);
    for (int k = 0; k < 3; ++k) V_n[k] = V_n[k] / U_n;
    double mu, phi, dummy;
    reac.sampleOut(mu, phi, dummy, p.E, *p.pRNG);
    V_n = rotateVector(V_n, mu, phi);
    for (int k = 0; k < 3; ++k) V_n[k] = V_n[k] * U_n;
    for (int k = 0; k < 3; ++k) V_n[k] = V_n[k] + V_cm[k];
    U_n = norm2(V_n);
    Vec3 dir_post; for (int k = 0; k < 3; ++k) dir_post[k] = V_n[k] / U_n;
    p.E = U_n * U_n;
    p.coords.point(dir_post);
    muL = dir_pre[0] * dir_post[0] + dir_pre[1] * dir_post[1] + dir_pre[2] * dir_post[2];
  }
  static double norm2(const Vec3& v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
};

}  // namespace orc
