// Host side of the drop-in boundary for continuous-energy decks: reads what aceNeutronDatabase%init reads (the ACE
// library file and the cards of the nuclides the materials name) and what materialMenu holds (composition, temperature),
// and hands them to the engine as the flat sb_ce_model of include/scone_b200.h.  In SCONE itself the Fortran shim would
// pass the arrays of the aceCard objects it has just read; this file does the reading where there is
// no Fortran.
//   NuclearData/ceNeutronData/aceLibrary_mod.f90                       library file: NAME; LINE; PATH;
//   NuclearData/DataDecks/ACE/aceCard_class.f90:1454-1534              readFromFile (header, NXS, JXS, XSS)
//   NuclearData/materialMenu_mod.f90 init_materialItem                 temp, composition (key order = nuclide order)
//   NuclearData/ceNeutronData/aceDatabase/aceNeutronDatabase_class.f90:873-1163   init: nuclide set, materials, options
//   CollisionOperator/CollisionProcessors/neutronCEstd_class.f90:110-150          minEnergy, maxEnergy, thresholds
// Product code (no oracle involvement).
#pragma once
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../../include/scone_b200.h"
#include "dict.hpp"
#include "model.hpp"

namespace sb {

struct AceCardData {
  std::string zaid; double aw = 0, tz = 0; int32_t nxs[16], jxs[32]; std::vector<double> xss;
  sb_ace_card view() const { sb_ace_card c{}; c.zaid = zaid.c_str(); c.aw = aw; c.tz = tz; c.nxs = nxs; c.jxs = jxs; c.xss = xss.data(); c.n_xss = (int64_t)xss.size(); return c; }
};

// aceCard%readFromFile: the card starts at line `lineNum` (1-based) of a type-1 (text) ACE file
inline AceCardData readAceText(const std::string& path, int lineNum) {
  std::ifstream f(path);
  if (!f) throw FatalError("readFromFile (aceCard)", "Cannot open ACE file: " + path);
  std::string line;
  for (int i = 1; i < lineNum; ++i) if (!std::getline(f, line)) throw FatalError("readFromFile (aceCard)", "ACE file is shorter than the requested line");
  if (!std::getline(f, line) || line.size() < 34) throw FatalError("readFromFile (aceCard)", "ACE header is missing");
  AceCardData c;
  c.zaid = line.substr(0, 10);
  { size_t a = c.zaid.find_first_not_of(' '), b = c.zaid.find_last_not_of(' '); c.zaid = (a == std::string::npos) ? "" : c.zaid.substr(a, b - a + 1); }
  c.aw = std::stod(line.substr(10, 12)); c.tz = std::stod(line.substr(22, 12));
  for (int i = 0; i < 5; ++i) std::getline(f, line);                     // comment line + 4 lines of IZ/AW pairs
  for (int i = 0; i < 16; ++i) f >> c.nxs[i];
  for (int i = 0; i < 32; ++i) f >> c.jxs[i];
  if (!f || c.nxs[0] < 1) throw FatalError("readFromFile (aceCard)", "NXS / JXS arrays could not be read");
  c.xss.resize((size_t)c.nxs[0]);
  std::string tok;
  for (int i = 0; i < c.nxs[0]; ++i) {
    if (!(f >> tok)) throw FatalError("readFromFile (aceCard)", "XSS array is truncated");
    c.xss[i] = std::strtod(tok.c_str(), nullptr);
  }
  return c;
}
// the same card as a binary fixture (data/ace/*.acebin): "SBACE1\0\0", ZAID[16], AW, TZ, NXS[16] i32, JXS[32] i32, n i64, XSS[n] f64
inline AceCardData readAceBinary(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw FatalError("readFromFile (aceCard)", "Cannot open ACE file: " + path);
  char magic[8], zaid[17]; long long n = 0; AceCardData c;
  memset(zaid, 0, sizeof(zaid));
  f.read(magic, 8); f.read(zaid, 16); f.read((char*)&c.aw, 8); f.read((char*)&c.tz, 8); f.read((char*)c.nxs, 64); f.read((char*)c.jxs, 128); f.read((char*)&n, 8);
  if (!f || std::string(magic, 6) != "SBACE1" || n < 1) throw FatalError("readFromFile (aceCard)", "Not a binary ACE card: " + path);
  c.zaid = zaid; c.xss.resize((size_t)n);
  f.read((char*)c.xss.data(), 8 * n);
  if (!f) throw FatalError("readFromFile (aceCard)", "Binary ACE card is truncated: " + path);
  return c;
}

struct AceLibEntry { std::string path; int line = 1; };
inline std::map<std::string, AceLibEntry> loadAceLibrary(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw FatalError("aceLib_load", "Cannot open the ACE library file: " + path);
  const size_t slash = path.find_last_of('/');
  const std::string dir = slash == std::string::npos ? "" : path.substr(0, slash);
  std::map<std::string, AceLibEntry> lib; std::string line;
  while (std::getline(f, line)) {
    size_t c = line.find('!'); if (c != std::string::npos) line = line.substr(0, c);
    std::vector<std::string> fld; std::stringstream ss(line); std::string t;
    while (std::getline(ss, t, ';')) { size_t a = t.find_first_not_of(" \t\r"), b = t.find_last_not_of(" \t\r"); if (a != std::string::npos) fld.push_back(t.substr(a, b - a + 1)); }
    if (fld.empty()) continue;
    if (fld.size() < 3) throw FatalError("aceLib_load", "Ill-formatted line in the ACE library: " + line);
    AceLibEntry e; e.line = Dict::toInt(fld[1]); e.path = fld[2];
    if (e.path[0] != '/' && !dir.empty()) e.path = dir + "/" + e.path;     // relative paths: relative to the library file
    lib[fld[0]] = e;
  }
  return lib;
}

struct FlatCeData {
  std::vector<AceCardData> cards; std::vector<std::string> nuclideNames;
  int nMat = 0; std::vector<int> matOff{0}, matNuc; std::vector<double> matDens, matTemp; std::vector<int> active;
  double collisionXS = 0.0, energyPerFission = 202.27;
  double minE = 1.0E-11, maxE = 20.0, threshE = 400.0, threshA = 1.0, sourceE = 1.0E-6;
  std::vector<sb_ace_card> views;
  sb_ce_model view() {
    views.clear(); for (auto& c : cards) views.push_back(c.view());
    sb_ce_model m{};
    m.n_nuc = (int)cards.size(); m.cards = views.data();
    m.n_mat = nMat; m.mat_off = matOff.data(); m.mat_nuc = matNuc.data(); m.mat_dens = matDens.data(); m.mat_temp = matTemp.data();
    m.n_active = (int)active.size(); m.active_mats = active.data();
    m.collision_xs = collisionXS; m.energy_per_fission = energyPerFission;
    m.min_energy = minE; m.max_energy = maxE; m.thresh_energy = threshE; m.thresh_mass = threshA; m.source_energy = sourceE;
    return m;
  }
};

inline FlatCeData buildCeData(const Dict& nuclearData, const std::string& handle, const std::string& baseDir, const std::vector<int>& activeMats,
                              const Dict& collisionOperator) {
  FlatCeData D;
  const Dict& h = nuclearData.getDict("handles").getDict(handle);
  if (h.getWord("type") != "aceNeutronDatabase") throw FatalError("ndReg", "continuous-energy data must be an aceNeutronDatabase");
  if (h.isPresent("avgDist")) {
    double t = h.getReal("avgDist");
    if (t <= 0.0) throw FatalError("init (aceNeutronDatabase)", "Must have a finite, positive minimum average collision distance");
    D.collisionXS = 1.0 / t;
  }
  D.energyPerFission = h.getReal("energyPerFission", 202.27);
  if (h.getBool("ures", false)) throw FatalError("init (aceNeutronDatabase)", "URR probability tables (ures 1) are not on the device");
  if (h.isPresent("DBRC")) throw FatalError("init (aceNeutronDatabase)", "DBRC is not on the device");
  if (!h.getBool("majorant", true)) throw FatalError("init (aceNeutronDatabase)", "the device path needs the unionised majorant (majorant 1)");
  std::string libPath = h.getWord("aceLibrary");
  if (!libPath.empty() && libPath[0] == '$') {
    const char* env = std::getenv(libPath.c_str() + 1);
    if (!env) throw FatalError("init (aceNeutronDatabase)", "EnVar " + libPath + " does not exist! Need to point to ACE Library");
    libPath = env;
  } else if (!libPath.empty() && libPath[0] != '/') libPath = baseDir + "/" + libPath;
  auto lib = loadAceLibrary(libPath);
  std::vector<std::string> names;
  materialMenu(nuclearData, &names);
  const Dict& md = nuclearData.getDict("materials");
  std::map<std::string, int> nucIdx;
  for (auto& n : names) {
    const Dict& m = md.getDict(n);
    if (m.getBool("tms", false)) throw FatalError("init_materialItem", "TMS is not on the device");
    if (m.isPresent("moder")) throw FatalError("init_materialItem", "S(alpha,beta) data are not on the device");
    double T = m.getReal("temp", 0.0);
    if (T < 0.0) throw FatalError("init_materialItem", "The temperature of material " + n + " is negative");
    const Dict& comp = m.getDict("composition");
    auto keys = comp.keys("all");
    if (keys.empty()) throw FatalError("setComposition", "Empty composition is not allowed");
    for (auto& key : keys) {
      auto it = nucIdx.find(key);
      if (it == nucIdx.end()) {
        auto le = lib.find(key + "c");
        if (le == lib.end()) le = lib.find(key);
        if (le == lib.end()) throw FatalError("new_neutronACE", "Nuclide " + key + " was not found in the ACE library " + libPath);
        const std::string& p = le->second.path;
        const bool bin = p.size() > 7 && p.compare(p.size() - 7, 7, ".acebin") == 0;
        D.cards.push_back(bin ? readAceBinary(p) : readAceText(p, le->second.line));
        D.nuclideNames.push_back(key);
        it = nucIdx.emplace(key, (int)D.cards.size()).first;
      }
      double dens = comp.getReal(key);
      if (dens < 0.0) throw FatalError("setComposition", "-ve nuclide densities are present");
      D.matNuc.push_back(it->second); D.matDens.push_back(dens);
    }
    D.matOff.push_back((int)D.matNuc.size()); D.matTemp.push_back(T);
  }
  D.nMat = (int)names.size();
  D.active = activeMats;
  const Dict& c = collisionOperator.getDict("neutronCE");
  if (c.getWord("type") != "neutronCEstd") throw FatalError("collisionOperator", "neutronCEstd is required for continuous-energy neutrons");
  D.minE = c.getReal("minEnergy", 1.0E-11); D.maxE = c.getReal("maxEnergy", 20.0);
  D.threshE = c.getReal("energyThreshold", 400.0); D.threshA = c.getReal("massThreshold", 1.0);
  if (c.getBool("makePrec", false) || c.getBool("neglectDelayed", false)) throw FatalError("init (neutronCEstd)", "makePrec / neglectDelayed are not on the device");
  if (D.minE < 0.0) throw FatalError("init (neutronCEstd)", "-ve minEnergy");
  if (D.maxE < 0.0) throw FatalError("init (neutronCEstd)", "-ve maxEnergy");
  if (D.minE >= D.maxE) throw FatalError("init (neutronCEstd)", "minEnergy >= maxEnergy");
  if (D.threshE < 0) throw FatalError("init (neutronCEstd)", " -ve energyThreshold");
  if (D.threshA < 0) throw FatalError("init (neutronCEstd)", " -ve massThreshold");
  return D;
}

}  // namespace sb
