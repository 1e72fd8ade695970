// ORACLE (test infrastructure, not product code).
// CPU restatement of SCONE's history loop for multigroup neutrons: particle, dungeon,
// fission source, tallies, DT/ST/HT transport, neutronMGstd collisions and the
// eigenPhysicsPackage cycle driver.  History-based, scalar, OpenMP over histories,
// per-thread tally columns -- as the reference.  Each function cites what it follows.
//
// Math mode: the reference calls the Fortran intrinsics (glibc libm on CPU). MATH_LIBM
// does the same.  MATH_SB routes log/sin/cos through scone_b200/csrc/sb_math.h, the
// deterministic implementation the CUDA engine uses, so that oracle and engine follow
// bit-identical histories (the functions themselves are pinned against libm to <= 1 ulp
// in tests/test_sb_math.py).
#pragma once
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "geom.hpp"
#include "mathmode.hpp"
#include "mgdata.hpp"
#include "rng.hpp"

namespace orc {

// SharedModules/genericProcedures.f90:1047-1084
inline Vec3 rotateVector(const Vec3& dir, double mu, double phi) {
  double sinPol, cosPol;
  msincos(phi, sinPol, cosPol);
  double u = dir[0], v = dir[1], w = dir[2];
  double A = std::sqrt(std::max(0.0, 1.0 - mu * mu));
  double B = std::sqrt(std::max(0.0, 1.0 - w * w));
  Vec3 n;
  if (B > 1E-8) {
    n[0] = mu * u + A * (u * w * cosPol - v * sinPol) / B;
    n[1] = mu * v + A * (v * w * cosPol + u * sinPol) / B;
    n[2] = mu * w - A * B * cosPol;
  } else {
    B = std::sqrt(std::max(0.0, 1.0 - v * v));
    n[0] = mu * u + A * (u * v * cosPol + w * sinPol) / B;
    n[1] = mu * v - A * B * cosPol;
    n[2] = mu * w + A * (v * w * cosPol - u * sinPol) / B;
  }
  double nrm = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  for (int k = 0; k < 3; ++k) n[k] = n[k] / nrm;
  return n;
}

// Tallies/tallyCodes.f90
constexpr int NO_FATE = 5000, ABS_FATE = 5001, LEAK_FATE = 5002;

// ParticleObjects/particle_class.f90:50-65  (fields used by the MG eigenvalue path)
struct ParticleState {
  double wgt = 0.0;
  Vec3 r, dir;
  double E = 0.0;
  int G = 0;
  bool isMG = false;
  double time = 0.0;
  int matIdx = -1, cellIdx = -1, uniqueID = -1, collisionN = 0, broodID = 0;
};

struct Particle {
  CoordList coords;
  double E = 0.0; int G = 0; double w = 0.0, w0 = 0.0, time = 0.0;
  bool isDead = false, isMG = true;
  int fate = NO_FATE, collisionN = 0, broodID = 0;
  RNG* pRNG = nullptr;
  double k_eff = 1.0;
  ParticleState preHistory, preTransition, prePath, preCollision;

  int matIdx() const { return coords.matIdx; }
  ParticleState state() const {                                   // particleState_fromParticle
    ParticleState s;
    s.wgt = w; s.r = coords.lvl[0].r; s.dir = coords.lvl[0].dir; s.E = E; s.G = G; s.isMG = isMG;
    s.time = time; s.matIdx = coords.matIdx; s.uniqueID = coords.uniqueID;
    s.cellIdx = coords.lvl[std::max(coords.nesting, 1) - 1].cellIdx;
    s.collisionN = collisionN; s.broodID = broodID;
    return s;
  }
  void fromState(const ParticleState& s) {                        // particle_class.f90:291-312
    w = s.wgt; w0 = s.wgt;
    coords.takeAboveGeom();
    coords.lvl[0].r = s.r; coords.lvl[0].dir = s.dir;
    E = s.E; G = s.G; isMG = s.isMG; time = s.time;
    fate = NO_FATE; collisionN = s.collisionN; broodID = s.broodID;
  }
};

// ===========================================================================
// particleDungeon  (ParticleObjects/particleDungeon_class.f90)
// ===========================================================================
struct HeapQueue {                                                  // DataStructures/heapQueue_class.f90:70-160
  std::vector<double> heap; int size = 0;
  void init(int maxSize) { heap.assign(maxSize, 0.0); size = 0; }
  void pushReplace(double val) { if (size < (int)heap.size()) push(val); else replace(val); }
  void push(double val) {
    size += 1; heap[size - 1] = val;
    if (size == 1) return;
    int child = size, parent = child / 2;
    while (heap[parent - 1] < heap[child - 1]) {
      std::swap(heap[parent - 1], heap[child - 1]);
      child = parent; parent = child / 2;
      if (parent == 0) return;
    }
  }
  void replace(double val) {
    heap[0] = val;
    int parent = 1, child = 2;
    while (child <= size) {
      if (child != size && heap[child - 1] < heap[child]) child += 1;
      if (heap[parent - 1] >= heap[child - 1]) return;
      std::swap(heap[parent - 1], heap[child - 1]);
      parent = child; child = parent * 2;
    }
  }
  double maxValue() const { if (size == 0) throw FatalError("maxValue (heapQueue)", "The queue is empty!"); return heap[0]; }
};

struct Dungeon {
  std::vector<ParticleState> prisoners;
  int pop = 0;
  double k_eff = 1.0;
  void init(int maxSize) { prisoners.assign(maxSize, ParticleState()); pop = 0; }
  void detain(const ParticleState& p) {                             // :157-179 (atomic capture of pop)
    int idx;
#pragma omp atomic capture
    { pop += 1; idx = pop; }
    if (idx > (int)prisoners.size()) throw FatalError("detain_particle", "Run out of space for particles.");
    prisoners[idx - 1] = p;
  }
  void setSize(int n) {
    if (n <= 0) throw FatalError("setSize", "Requested population is not +ve");
    pop = n;
    if ((int)prisoners.size() < n) prisoners.resize(n);
    for (auto& p : prisoners) p = ParticleState();
  }
  double popWeight() const { double s = 0.0; for (int i = 0; i < pop; ++i) s += prisoners[i].wgt; return s; }
  // printToFile (particleDungeon_class.f90:1077-1112): r, dir, E, real(G), real(broodID), wgt per prisoner; stream binary or one text row.
  // Fortran's list-directed text is compiler-formatted; here 17 significant digits, which reads back to the same doubles.
  void printToFile(const std::string& name, bool writeBinary) const {
    FILE* f = fopen((name + (writeBinary ? ".bin" : ".txt")).c_str(), writeBinary ? "wb" : "w");
    if (!f) throw FatalError("printToFile", "cannot open " + name);
    for (int i = 0; i < pop; ++i) {
      const ParticleState& p = prisoners[i];
      double row[10] = {p.r[0], p.r[1], p.r[2], p.dir[0], p.dir[1], p.dir[2], p.E, (double)p.G, (double)p.broodID, p.wgt};
      if (writeBinary) fwrite(row, sizeof(double), 10, f);
      else { for (int k = 0; k < 10; ++k) fprintf(f, "%s%.17g", k ? " " : "  ", row[k]); fprintf(f, "\n"); }
    }
    fclose(f);
  }

  // :923-981, transcribed literally including the in-place cycle permutation
  void sortByBroodID(int k) {
    bool allZero = true;
    for (auto& p : prisoners) if (p.broodID != 0) { allZero = false; break; }
    if (allZero) return;
    std::vector<int> count(k, 0);
    for (int i = 0; i < pop; ++i) {
      int id = prisoners[i].broodID;
      if (id < 1 || id > k) throw FatalError("sortBybroodID", "Brood ID out of range");
      count[id - 1] += 1;
    }
    int loc = 1;
    for (int i = 0; i < k; ++i) { int c = count[i]; count[i] = loc; loc += c; }
    std::vector<int> perm(pop);
    for (int i = 1; i <= pop; ++i) {
      int id = prisoners[i - 1].broodID;
      loc = count[id - 1];
      count[id - 1] += 1;
      perm[loc - 1] = i;
    }
    for (int i = 1; i <= pop; ++i) {
      int j = i;
      while (i != perm[i - 1]) {
        loc = perm[i - 1];
        if (loc != j) std::swap(prisoners[j - 1], prisoners[loc - 1]);
        std::swap(perm[i - 1], perm[loc - 1]);
        j = loc;
      }
    }
  }

  // :431-602, single rank (nRanks = 1; MPI collectives degenerate to identities)
  void normSize_Repr(int totPop, RNG& rand) {
    // an empty fission bank: the reference divides by the number of sites (:478-481); the restatement stops with a message instead
    if (pop <= 0) throw FatalError("normSize_Repr", "the fission bank is empty");
    int maxBroodID = 0;
    for (int i = 0; i < pop; ++i) maxBroodID = std::max(maxBroodID, prisoners[i].broodID);
    sortByBroodID(maxBroodID);
    double threshold = 1.0;
    uint64_t seed0 = 0;
    int totSites = pop;
    int excess = totSites - totPop;
    int heapSize;
    if (excess < 0) heapSize = ((-excess) % totSites + totSites) % totSites;
    else heapSize = excess;
    if (heapSize != 0) {
      RNG masterRand = rand;
      HeapQueue heap; heap.init(heapSize);
      heap.pushReplace(2.0);
      seed0 = masterRand.currentState();
      for (int j = 0; j < pop; ++j) {
        double rn = masterRand.get();
        if (rn < heap.maxValue()) heap.pushReplace(rn);
      }
      threshold = heap.maxValue();
    }
    RNG rankRand; rankRand.init((int64_t)seed0);
    if (excess > 0) {
      std::vector<int> keepers; keepers.reserve(pop);
      for (int i = 1; i <= pop; ++i) if (rankRand.get() > threshold) keepers.push_back(i);
      for (size_t i = 1; i <= keepers.size(); ++i) if ((int)i != keepers[i - 1]) prisoners[i - 1] = prisoners[keepers[i - 1] - 1];
      pop = (int)keepers.size();
    } else if (excess < 0) {
      totSites = excess + totPop;
      int n_copies = -excess / totSites;
      int n_duplicates = ((-excess) % totSites + totSites) % totSites;
      if ((long)pop * (n_copies + 1) + n_duplicates > (long)prisoners.size()) prisoners.resize((size_t)pop * (n_copies + 2));
      for (int i = 1; i <= n_copies; ++i) for (int j = 0; j < pop; ++j) prisoners[pop * i + j] = prisoners[j];
      int count = pop * (n_copies + 1);
      if (n_duplicates != 0)
        for (int i = 0; i < pop; ++i) if (rankRand.get() <= threshold) { prisoners[count] = prisoners[i]; count += 1; }
      pop = count;
      maxBroodID = 0;
      for (int i = 0; i < pop; ++i) maxBroodID = std::max(maxBroodID, prisoners[i].broodID);
      sortByBroodID(maxBroodID);
    }
    if (pop != totPop) throw FatalError("normSize", "Normalisation failed!");
  }
};

// ===========================================================================
// Tallies
// ===========================================================================
// Tallies/scoreMemory_class.f90
struct ScoreMemory {
  static constexpr int ARRAY_PAD = 64;
  long N = 0; int nThreads = 1, batchN = 0, cycles = 0, batchSize = 1;
  std::vector<double> bin, csum, csum2;
  std::vector<double> parallelBins;      // (N + pad) x nThreads
  void init(long n, int batch = 1) {
    N = n; bin.assign(n, 0.0); csum.assign(n, 0.0); csum2.assign(n, 0.0);
    nThreads = omp_get_max_threads();
    parallelBins.assign((size_t)(n + ARRAY_PAD) * nThreads, 0.0);
    batchN = 0; cycles = 0; batchSize = batch;
  }
  void score(double s, long idx) {                                  // 1-based idx; :215-233
    if (idx < 0 || idx > N) throw FatalError("score_defReal", "Index is outside bounds of memory");
    int t = omp_get_thread_num();
    parallelBins[(size_t)t * (N + ARRAY_PAD) + (idx - 1)] += s;
  }
  void accumulate(double s, long idx) { csum[idx - 1] = csum[idx - 1] + s; csum2[idx - 1] = csum2[idx - 1] + s * s; }
  bool lastCycle() const { return (cycles + 1) % batchSize == 0; }
  void reduceBins() {                                               // :404-431
    if (!lastCycle()) return;
    for (long i = 0; i < N; ++i) {
      double s = 0.0;
      for (int t = 0; t < nThreads; ++t) { s += parallelBins[(size_t)t * (N + ARRAY_PAD) + i]; parallelBins[(size_t)t * (N + ARRAY_PAD) + i] = 0.0; }
      bin[i] = s;
    }
  }
  void resetBin(long idx) { if (idx > 0 && idx <= N) bin[idx - 1] = 0.0; }
  double getScore(long idx) const { if (idx <= 0 || idx > N) return 0.0; return bin[idx - 1]; }
  void closeCycle(double normFactor) {                              // :309-342
    cycles += 1;
    if (cycles % batchSize == 0) {
      for (long i = 0; i < N; ++i) {
        double res = bin[i] * normFactor;
        bin[i] = 0.0;
        csum[i] = csum[i] + res;
        csum2[i] = csum2[i] + res * res;
      }
      batchN += 1;
    }
  }
  void closeBin(double normFactor, long idx) {                      // :344-362
    if (idx < 0 || idx > N) throw FatalError("closeBin (scoreMemory)", "Index is outside bounds of memory");
    double res = bin[idx - 1] * normFactor;
    csum[idx - 1] = csum[idx - 1] + res;
    csum2[idx - 1] = csum2[idx - 1] + res * res;
    bin[idx - 1] = 0.0;
  }
  void getResult(double& mean, double& STD, long idx, int samples = -1) const {       // :537-573
    if (idx < 0 || idx > N) { mean = 0.0; STD = 0.0; return; }
    int n = samples > 0 ? samples : batchN;
    mean = csum[idx - 1] / n;
    double inv_N = 1.0 / n, inv_Nm1 = (n != 1) ? 1.0 / (n - 1) : 1.0;
    STD = csum2[idx - 1] * inv_N * inv_Nm1 - mean * mean * inv_Nm1;
    STD = std::sqrt(STD);
  }
};

// SharedModules/grid_class.f90:35-176
struct Grid {
  enum { LIN = 1, LOGAR = 2, UNSTRUCT = 3 };
  std::vector<double> bins; double step = 0.0; int type = 0;
  void initEqual(double mini, double maxi, int N, const std::string& t) {
    if (N < 1) throw FatalError("init_equalSpaced", "Number of bins must be +ve");
    if (std::fabs((maxi - mini) / maxi) < FP_REL_TOL) throw FatalError("init_equalSpaced", "Minimum value must be smaller then maximum");
    bins.assign(N + 1, 0.0);
    if (t == "lin") {
      step = (maxi - mini) / N;
      bins[0] = mini;
      for (int i = 2; i <= N + 1; ++i) bins[i - 1] = mini + (i - 1) * step;
      type = LIN;
    } else if (t == "log") {
      if (mini <= 0) throw FatalError("init_equalSpaced", "For logarithmic grid minimum must be +ve");
      step = std::log(maxi / mini) / N;
      bins[0] = mini;
      for (int i = 2; i <= N + 1; ++i) bins[i - 1] = bins[i - 2] * std::exp(step);
      type = LOGAR;
    } else throw FatalError("init_equalSpaced", "Grid type must be lin or log");
  }
  void initUnstruct(const std::vector<double>& b) {
    if (b.size() < 2) throw FatalError("init_unstruct", "Empty array or array of size 1 was provided");
    for (size_t i = 1; i < b.size(); ++i) if (b[i] < b[i - 1]) throw FatalError("init_unstruct", "Provided grid is not sorted");
    bins = b; type = UNSTRUCT;
  }
  static int binarySearch(const std::vector<double>& a, double value) {   // genericProcedures.f90:132-166
    int bottom = 1, top = (int)a.size();
    if (value < a[bottom - 1] || value > a[top - 1]) return -1;
    int idx = 0;
    for (int i = 0; i < 70; ++i) {
      idx = (top + bottom) / 2;
      if (bottom == idx) return idx;
      if (a[idx - 1] <= value) bottom = idx; else top = idx;
    }
    return -2;
  }
  int search(double value) const {
    int idx = 0;
    if (type == LIN) idx = (int)std::floor((value - bins[0]) / step) + 1;
    else if (type == LOGAR) idx = (int)std::floor(mlog(value / bins[0]) / step) + 1;
    else if (type == UNSTRUCT) idx = binarySearch(bins, value);
    if (idx < 1 || idx >= (int)bins.size()) idx = -1;   // valueOutsideArray
    return idx;
  }
};

// Tallies/TallyMaps
struct TallyMap {
  virtual ~TallyMap() = default;
  virtual int bins() const = 0;
  virtual int map(const ParticleState& s) const = 0;
};
struct SpaceMap : TallyMap {                                        // Maps1D/spaceMap_class.f90
  Grid grid; int N = 0, dir = 0;
  void init(const Dict& d) {
    std::string ax = d.getWord("axis");
    if (ax == "x") dir = 0; else if (ax == "y") dir = 1; else if (ax == "z") dir = 2;
    else throw FatalError("init (spaceMap)", "Unrecognised axis");
    std::string g = d.getWord("grid");
    if (g == "lin") { N = d.getInt("N"); grid.initEqual(d.getReal("min"), d.getReal("max"), N, "lin"); }
    else if (g == "unstruct") { auto b = d.getRealArray("bins"); grid.initUnstruct(b); N = (int)b.size() - 1; }
    else throw FatalError("init (spaceMap)", "'grid' keyword must be: lin or unstruct");
  }
  int bins() const override { return N; }
  int map(const ParticleState& s) const override { int idx = grid.search(s.r[dir]); return idx == -1 ? 0 : idx; }
};
struct EnergyMap : TallyMap {                                       // Maps1D/energyMap_class.f90
  Grid grid; int N = 0;
  void init(const Dict& d) {
    std::string g = d.getWord("grid");
    if (g == "lin" || g == "log") { N = d.getInt("N"); grid.initEqual(d.getReal("min"), d.getReal("max"), N, g); }
    else if (g == "unstruct") { auto b = d.getRealArray("bins"); std::sort(b.begin(), b.end()); grid.initUnstruct(b); N = (int)b.size() - 1; }
    else if (g == "predef") {                                       // build_predef (energyMap_class.f90:137-177): named grid, thermal to fast
      auto b = sb::namedEnergyGrid(d.getWord("name"));
      if (b.empty()) throw FatalError("build_predef (energyMap)", "Grid " + d.getWord("name") + " is undefined!");
      grid.initUnstruct(b); N = (int)b.size() - 1;
    }
    else throw FatalError("init (energyMap)", "'grid' keyword must be: lin, log, unstruct or predef");
  }
  int bins() const override { return N; }
  int map(const ParticleState& s) const override {
    if (s.isMG) return 0;
    int idx = grid.search(s.E); return idx == -1 ? 0 : idx;
  }
};
struct MaterialMap : TallyMap {                                     // Maps1D/materialMap_class.f90
  std::map<int, int> binMap; int Nbins = 0, def = 0;
  void init(const Dict& d, const std::map<std::string, int>& mats) {
    auto names = d.getWordArray("materials");
    std::string undef = d.getWord("undefBin", "false");
    bool track;
    if (undef == "yes" || undef == "y" || undef == "true" || undef == "TRUE" || undef == "T") track = true;
    else if (undef == "no" || undef == "n" || undef == "false" || undef == "FALSE" || undef == "F") track = false;
    else throw FatalError("init (materialMap)", undef + " is an unrecognised entry!");
    int i = 0;
    for (auto& n : names) {
      auto it = mats.find(n);
      if (it == mats.end()) throw FatalError("build (materialMap)", "Material " + n + " does not exist in the input materials");
      binMap[it->second] = ++i;
    }
    if (track) { Nbins = i + 1; def = i + 1; } else { Nbins = i; def = 0; }
  }
  int bins() const override { return Nbins; }
  int map(const ParticleState& s) const override { auto it = binMap.find(s.matIdx); return it == binMap.end() ? def : it->second; }
};
std::unique_ptr<TallyMap> newTallyMap(const Dict& d, const std::map<std::string, int>& mats);
struct MultiMap : TallyMap {                                        // multiMap_class.f90:100-173
  std::vector<std::unique_ptr<TallyMap>> maps; std::vector<int> multi;
  void init(const Dict& d, const std::map<std::string, int>& mats) {
    for (auto& n : d.getWordArray("maps")) maps.push_back(newTallyMap(d.getDict(n), mats));
    int mul = 1;
    for (auto& m : maps) { multi.push_back(mul); mul *= m->bins(); }
  }
  int bins() const override { int n = 1; for (auto& m : maps) n *= m->bins(); return n; }
  int map(const ParticleState& s) const override {
    int idx = 1;
    for (size_t i = 0; i < maps.size(); ++i) {
      int b = maps[i]->map(s);
      if (b == 0) return 0;
      idx = idx + (b - 1) * multi[i];
    }
    return idx;
  }
};
inline std::unique_ptr<TallyMap> newTallyMap(const Dict& d, const std::map<std::string, int>& mats) {
  std::string t = d.getWord("type");
  if (t == "spaceMap") { auto m = std::make_unique<SpaceMap>(); m->init(d); return m; }
  if (t == "energyMap") { auto m = std::make_unique<EnergyMap>(); m->init(d); return m; }
  if (t == "materialMap") { auto m = std::make_unique<MaterialMap>(); m->init(d, mats); return m; }
  if (t == "multiMap") { auto m = std::make_unique<MultiMap>(); m->init(d, mats); return m; }
  throw FatalError("new_tallyMap", "Unrecognised / unsupported type of tallyMap in oracle: " + t);
}

// what the clerks ask of the nuclear database for the particle at hand (nuclearDatabase_inter.f90:38-53: getTotalMatXS,
// getMaterial -> getMacroXSs); implemented over the MG database here and over the CE database in cephysics.hpp
struct XsView {
  virtual ~XsView() = default;
  virtual int nMat() const = 0;
  virtual double totalMatXS(const Particle& p, int matIdx) const = 0;
  virtual void macroXSs(MacroXSs& x, const Particle& p, int matIdx) const = 0;
  // tracking interface (getTrackMatXS / getMajorantXS; collisionXS = 1/avgDist of the database)
  virtual double trackMatXS(const Particle& p, int matIdx) const = 0;
  virtual double majorantXS(const Particle& p) const = 0;
  virtual double collisionXS() const = 0;
  virtual bool isFissileMat(int matIdx) const = 0;
};
struct MgXsView : XsView {
  const MgDatabase* db = nullptr;
  int nMat() const override { return (int)db->mats.size(); }
  double totalMatXS(const Particle& p, int matIdx) const override { return db->getTotalMatXS(p.G, matIdx); }
  void macroXSs(MacroXSs& x, const Particle& p, int matIdx) const override { db->mats.at(matIdx - 1).getMacroXSs(x, p.G); }
  double trackMatXS(const Particle& p, int matIdx) const override { return db->getTrackMatXS(p.G, matIdx); }
  double majorantXS(const Particle& p) const override { return db->getMajorantXS(p.G); }
  double collisionXS() const override { return db->collisionXS; }
  bool isFissileMat(int matIdx) const override { return db->mats.at(matIdx - 1).fissile; }
};

// Tallies/TallyResponses: fluxResponse (=1) and macroResponse (macroResponse_class.f90:70-172)
struct Response {
  bool isFlux = true; int MT = 0;
  void init(const Dict& d) {
    std::string t = d.getWord("type");
    if (t == "fluxResponse") { isFlux = true; return; }
    if (t != "macroResponse") throw FatalError("new_tallyResponse", "Unsupported response in oracle: " + t);
    isFlux = false;
    int mt = d.getInt("MT");
    if (mt > 0) {
      switch (mt) {
        case 1: MT = macroTotal; break;
        case 2: MT = macroEscatter; break;
        case 3: MT = macroNonElastic; break;
        case 101: MT = macroDisappearance; break;
        case 18: MT = macroFission; break;
        case 27: MT = macroAbsorbtion; break;
        case 301: MT = macroKappaFission; break;    // N_KAPPA
        default: throw FatalError("build (macroResponse)", "MT numbers outside main data are not supported for MG");
      }
    } else MT = mt;
  }
  double get(const XsView& db, const Particle& p, int matOverride = -1) const {      // matOverride: trackClerk scores in the pre-path material
    if (isFlux) return 1.0;
    int matIdx = matOverride >= 0 ? matOverride : p.matIdx();
    if (matIdx == VOID_MAT) return 0.0;
    if (matIdx < 1 || matIdx > db.nMat()) return 0.0;
    MacroXSs x; db.macroXSs(x, p, matIdx);
    return x.get(MT);
  }
};

struct Clerk {
  enum Kind { COLLISION, KEFF_ANALOG, KEFF_IMPLICIT, TRACK, SHANNON } kind = COLLISION;
  int maxCycles = 0, currentCycle = 0;                                // shannonEntropyClerk
  std::string name;
  long addr = 1;
  // collisionClerk
  std::unique_ptr<TallyMap> map; std::vector<Response> response; bool handleVirtual = true;
  long size() const {
    if (kind == KEFF_ANALOG) return 3;
    if (kind == KEFF_IMPLICIT) return 5;
    if (kind == SHANNON) return map->bins() + 1 + maxCycles;          // shannonEntropyClerk_class.f90:106-111
    long S = (long)response.size();
    if (map) S *= map->bins();
    return S;
  }
};

// Tallies/tallyAdmin_class.f90 (one admin + at most one attachment, as eigenPP builds them)
struct TallyAdmin {
  std::vector<Clerk> clerks;
  ScoreMemory mem;
  long normBinAddr = -1; double normValue = 1.0;
  TallyAdmin* atch = nullptr;

  void init(const Dict& d, const std::map<std::string, int>& mats) {
    for (auto& n : d.keys("dict")) {
      const Dict& cd = d.getDict(n);
      std::string t = cd.getWord("type");
      Clerk c; c.name = n;
      if (t == "collisionClerk") {
        c.kind = Clerk::COLLISION;
        if (cd.isPresent("filter")) throw FatalError("collisionClerk init", "filters are not supported in oracle");
        if (cd.isPresent("map")) c.map = newTallyMap(cd.getDict("map"), mats);
        for (auto& rn : cd.getWordArray("response")) { Response r; r.init(cd.getDict(rn)); c.response.push_back(r); }
        c.handleVirtual = cd.getBool("handleVirtual", true);
      } else if (t == "trackClerk") {                                 // trackClerk_class.f90:100-150 (same dictionary as collisionClerk, no handleVirtual)
        c.kind = Clerk::TRACK;
        if (cd.isPresent("filter")) throw FatalError("trackClerk init", "filters are not supported in oracle");
        if (cd.isPresent("map")) c.map = newTallyMap(cd.getDict("map"), mats);
        for (auto& rn : cd.getWordArray("response")) { Response r; r.init(cd.getDict(rn)); c.response.push_back(r); }
      } else if (t == "keffAnalogClerk") c.kind = Clerk::KEFF_ANALOG;
      else if (t == "keffImplicitClerk") { c.kind = Clerk::KEFF_IMPLICIT; c.handleVirtual = cd.getBool("handleVirtual", true); }
      else if (t == "shannonEntropyClerk") {                           // shannonEntropyClerk_class.f90:75-92
        c.kind = Clerk::SHANNON;
        c.map = newTallyMap(cd.getDict("map"), mats);
        c.maxCycles = cd.getInt("cycles");
      }
      else throw FatalError("new_tallyClerk", "Unsupported clerk in oracle: " + t);
      clerks.push_back(std::move(c));
    }
    int batch = d.getInt("batchSize", 1);
    long memSize = 0;
    for (auto& c : clerks) memSize += c.size();
    mem.init(memSize, batch);
    long loc = 1;
    for (auto& c : clerks) { c.addr = loc; loc += c.size(); }
    if (d.isPresent("norm")) {
      std::string nn = d.getWord("norm");
      normValue = d.getReal("normVal");
      bool found = false;
      for (auto& c : clerks) if (c.name == nn) { normBinAddr = c.addr; found = true; }
      if (!found) throw FatalError("tallyAdmin init", "norm clerk not found: " + nn);
    }
  }

  // trackingXS: the value last stored in the tracking cache by the transport operator
  // (baseMgNeutronDatabase_class.f90:119-133)
  void reportInColl(const Particle& p, const XsView& db, double trackingXS, bool virt) {
    if (atch) atch->reportInColl(p, db, trackingXS, virt);
    for (auto& c : clerks) {
      if (c.kind == Clerk::COLLISION) {                             // collisionClerk_class.f90:192-244
        if (!c.handleVirtual && virt) continue;
        ParticleState s = p.state();
        int binIdx = c.map ? c.map->map(s) : 1;
        if (binIdx == 0) continue;
        double flux = c.handleVirtual ? p.w / trackingXS : p.w / db.totalMatXS(p, p.matIdx());
        long a = c.addr + (long)c.response.size() * (binIdx - 1) - 1;
        for (size_t i = 1; i <= c.response.size(); ++i) mem.score(c.response[i - 1].get(db, p) * flux, a + (long)i);
      } else if (c.kind == Clerk::KEFF_IMPLICIT) {                  // keffImplicitClerk_class.f90:180-236
        if (!c.handleVirtual && virt) continue;
        if (p.matIdx() == VOID_MAT) continue;
        double flux = c.handleVirtual ? p.w / trackingXS : p.w / db.totalMatXS(p, p.matIdx());
        MacroXSs x; db.macroXSs(x, p, p.matIdx());
        double s1 = x.nuFission * flux, s2 = (x.capture + x.fission) * flux;
        mem.score(s1, c.addr + 0);   // IMP_PROD
        mem.score(s2, c.addr + 1);   // IMP_ABS
      }
    }
  }
  void reportPath(const Particle& p, const XsView& db, double L) {   // tallyAdmin_class.f90:545-568 ; trackClerk_class.f90:185-232
    if (atch) atch->reportPath(p, db, L);
    for (auto& c : clerks) if (c.kind == Clerk::TRACK) {
      const ParticleState& s = p.prePath;
      int binIdx = c.map ? c.map->map(s) : 1;
      if (binIdx == 0) continue;
      long a = c.addr + (long)c.response.size() * (binIdx - 1) - 1;
      for (size_t i = 1; i <= c.response.size(); ++i) mem.score(c.response[i - 1].get(db, p, s.matIdx) * p.w * L, a + (long)i);
    }
  }
  void reportOutColl(const Particle& p, int MT) {                   // keffImplicitClerk_class.f90:238-270
    if (atch) atch->reportOutColl(p, MT);
    for (auto& c : clerks) if (c.kind == Clerk::KEFF_IMPLICIT) {
      double score = 0.0;
      if (MT == 16 || MT == 11 || MT == 24 || MT == 30 || MT == 41 || (MT >= 875 && MT <= 891)) score = 1.0 * p.preCollision.wgt;   // N_2N, N_2Nd, N_2Na, N_2N2a, N_2Np, N_2Nl(1):N_2Ncont
      else if (MT == 17 || MT == 25 || MT == 42) score = 2.0 * p.preCollision.wgt;                                                   // N_3N, N_3Na, N_3Np
      else if (MT == 37) score = 3.0 * p.preCollision.wgt;                                                                           // N_4N
      else if (MT == macroAllScatter || MT == macroIEscatter) score = std::max(p.w - p.preCollision.wgt, 0.0);
      if (score > 0.0) mem.score(score, c.addr + 2);   // SCATTER_PROD
    }
  }
  void reportHist(const Particle& p) {                              // keffImplicitClerk_class.f90:272-290
    if (atch) atch->reportHist(p);
    for (auto& c : clerks) if (c.kind == Clerk::KEFF_IMPLICIT) if (p.fate == LEAK_FATE) mem.score(p.w, c.addr + 3);   // ANA_LEAK
  }
  void reportCycleStart(const Dungeon& start) {                     // keffAnalogClerk_class.f90:132-140
    if (atch) atch->reportCycleStart(start);
    for (auto& c : clerks) if (c.kind == Clerk::KEFF_ANALOG) mem.score(start.popWeight(), c.addr + 0);
  }
  void reportCycleEnd(const Dungeon& end) {                         // tallyAdmin_class.f90:735-794
    if (atch) atch->reportCycleEnd(end);
    for (auto& c : clerks) if (c.kind == Clerk::KEFF_ANALOG) mem.score(end.popWeight(), c.addr + 1);
    for (auto& c : clerks) if (c.kind == Clerk::SHANNON) {            // shannonEntropyClerk_class.f90:117-144 reportCycleEnd
      c.currentCycle += 1;
      if (c.currentCycle > c.maxCycles) continue;
      mem.score(end.popWeight(), c.addr);
      for (int i = 0; i < end.pop; ++i) {
        int idx = c.map->map(end.prisoners[i]);
        if (idx == 0) continue;
        mem.score(end.prisoners[i].wgt, c.addr + idx);
      }
    }
    mem.reduceBins();
    for (auto& c : clerks) if (c.kind == Clerk::SHANNON && c.currentCycle <= c.maxCycles) {   // closeCycle :149-190
      const int N = c.map->bins();
      double totWgt = mem.getScore(c.addr), val = 0.0;
      const double one_log2 = 1.0 / mlog(2.0);
      for (int i = 1; i <= N; ++i) {
        double prob = mem.getScore(c.addr + i) / totWgt;
        if (prob > 0.0 && prob < 1.0) val = val - prob * mlog(prob) * one_log2;
      }
      mem.accumulate(val, c.addr + N + c.currentCycle);
      for (int i = 0; i <= N; ++i) mem.resetBin(c.addr + i);
    }
    for (auto& c : clerks) {
      if (c.kind == Clerk::KEFF_ANALOG && mem.lastCycle()) {        // keffAnalogClerk_class.f90:156-176
        double k_norm = end.k_eff;
        double startPopWgt = mem.getScore(c.addr + 0), endPopWgt = mem.getScore(c.addr + 1);
        double k = endPopWgt / startPopWgt * k_norm;
        mem.accumulate(k, c.addr + 2);
      } else if (c.kind == Clerk::KEFF_IMPLICIT && mem.lastCycle()) {   // keffImplicitClerk_class.f90:292-312
        double nuFiss = mem.getScore(c.addr + 0), absorb = mem.getScore(c.addr + 1);
        double leakage = mem.getScore(c.addr + 3), scatterMul = mem.getScore(c.addr + 2);
        double k_est = nuFiss / (absorb + leakage - scatterMul);
        mem.accumulate(k_est, c.addr + 4);
      }
    }
    double normFactor = 1.0;
    if (normBinAddr != -1) {
      double normScore = mem.getScore(normBinAddr);
      if (normScore == 0.0) throw FatalError("reportCycleEnd", "Normalisation score is 0");
      normFactor = normValue / normScore;
    }
    mem.closeCycle(normFactor);
  }
  bool getKeff(double& k, double& std_) const {
    for (auto& c : clerks) {
      if (c.kind == Clerk::KEFF_ANALOG) { mem.getResult(k, std_, c.addr + 2); return true; }
      if (c.kind == Clerk::KEFF_IMPLICIT) { mem.getResult(k, std_, c.addr + 4); return true; }
    }
    return false;
  }
};

// ===========================================================================
// Source  (ParticleObjects/Source/fissionSource_class.f90:149-271, source_inter.f90:98-118)
// ===========================================================================
struct FissionSource {
  const GeometryStd* geom = nullptr; const MgDatabase* db = nullptr; const XsView* xs = nullptr;
  double bottom[3], top[3]; int G = 1, attempts = 10000; double E = 1.0E-6;
  // continuous-energy branch of sampleParticle (fissionSource_class.f90:211-235): set by the CE driver
  std::function<void(ParticleState&, int, RNG&)> sampleCE;
  void init(const GeometryStd* g, const MgDatabase* d, const XsView* v) {
    geom = g; db = d; xs = v;
    double b[6]; g->bounds(b);
    for (int i = 0; i < 3; ++i) { bottom[i] = b[i]; top[i] = b[i + 3]; }
  }
  ParticleState sampleParticle(RNG& rand) const {
    ParticleState p;
    int i = 0;
    for (;;) {
      i += 1;
      if (i > attempts) throw FatalError("sampleParticle (fissionSource)", "Failed to find a fissile material");
      double r3[3];
      r3[0] = rand.get(); r3[1] = rand.get(); r3[2] = rand.get();
      Vec3 r;
      for (int k = 0; k < 3; ++k) r[k] = (top[k] - bottom[k]) * r3[k] + bottom[k];
      int matIdx, uniqueID;
      geom->whatIsAt(matIdx, uniqueID, r);
      if (matIdx == VOID_MAT || matIdx == OUTSIDE_MAT) continue;
      if (matIdx == UNDEF_MAT) throw FatalError("sampleParticle (fissionSource)", "Particle position was sampled in an undefined material");
      if (matIdx == OVERLAP_MAT) throw FatalError("sampleParticle (fissionSource)", "Particle position was sampled in an overlapping cell region");
      if (!xs->isFissileMat(matIdx)) continue;
      p.matIdx = matIdx; p.uniqueID = uniqueID; p.wgt = 1.0; p.time = 0.0; p.r = r;
      if (sampleCE) { sampleCE(p, matIdx, rand); return p; }
      const MgMaterial& mat = db->mats.at(matIdx - 1);
      double mu, phi; int G_out;
      mat.fissionSampleOut(mu, phi, G_out, rand);
      p.G = G_out; p.isMG = true;
      Vec3 ex; ex[0] = 1.0;
      p.dir = rotateVector(ex, mu, phi);
      return p;
    }
  }
  void generate(Dungeon& dungeon, int n, const RNG& rand) const {
    dungeon.setSize(n);
    std::string srcErr;                                              // an exception may not leave the parallel region
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= n; ++i) {
      if (!srcErr.empty()) continue;
      RNG pRand = rand;
      pRand.stride(i);
      try { dungeon.prisoners[i - 1] = sampleParticle(pRand); }
      catch (const std::exception& ex) {
#pragma omp critical
        srcErr = ex.what();
      }
    }
    if (!srcErr.empty()) throw FatalError("generate (source)", srcErr);
  }
};

// ===========================================================================
// Transport + collision operators and the eigenvalue driver
// ===========================================================================
enum Tracking { TRACK_DT = 0, TRACK_ST = 1, TRACK_HT = 2 };

struct EigenPP {
  // configuration (PhysicsPackages/eigenPhysicsPackage_class.f90:417-645)
  int pop = 0, N_inactive = 0, N_active = 0;
  double keff_0 = 1.0;
  int tracking = TRACK_DT; double htCutoff = 0.9; bool stCache = true;
  RNG pRNG;
  GeometryStd geom;
  MgDatabase db; MgXsView view; const XsView* xs = &view;
  virtual ~EigenPP() = default;
  TallyAdmin inactiveTally, activeTally, inactiveAtch, activeAtch;
  FissionSource source;
  Dungeon dungeonA, dungeonB; Dungeon* thisCycle = &dungeonA; Dungeon* nextCycle = &dungeonB;
  // fixedSourcePhysicsPackage (PhysicsPackages/fixedSourcePhysicsPackage_class.f90): cycles, private secondary buffer, pointSource
  bool fixedSource = false; int N_cycles = 0, bufferSize = 50;
  struct PointSource { Vec3 r, dir; bool isotropic = true, isMG = true; double E = 0.0; int G = 1; std::vector<double> probG; } psrc;
  // fileSource (ParticleObjects/Source/fileSource_class.f90): rows of a printToFile dump
  struct FileSource { bool on = false, isMG = false; long N = 0; std::vector<double> rows; } fsrc;
  // materialSource (ParticleObjects/Source/materialSource_class.f90)
  struct MaterialSource { bool on = false, isMG = false; int matIdx = 0, G = 1; double E = 1.0E-6; Vec3 bottom, top; } msrc;
  // printSource / outputFile (eigenPhysicsPackage_class.f90:278-281,463,501-504)
  int printSource = 0; std::string outputFile = "./output"; int cycleInPhase[2] = {0, 0};
  // statistics the reference does not keep (for the segments/s metric)
  long nSegments = 0, nCollisions = 0, nHistories = 0;
  std::vector<double> cycleK;            // k_new after each cycle (both phases)

  // the keys the two packages read differently (eigenPhysicsPackage_class.f90:417-470, fixedSourcePhysicsPackage_class.f90:296-330)
  void initCycles(const Dict& dict) {
    std::string t = dict.getWord("type", "eigenPhysicsPackage");
    fixedSource = (t == "fixedSourcePhysicsPackage");
    pop = dict.getInt("pop");
    if (fixedSource) { N_cycles = dict.getInt("cycles"); bufferSize = dict.getInt("buffer", 50); N_inactive = 0; N_active = N_cycles; }
    else { N_inactive = dict.getInt("inactive"); N_active = dict.getInt("active"); }
    outputFile = dict.getWord("outputFile", "./output");
    printSource = dict.getInt("printSource", 0);
    if (printSource < 0 || printSource > 2) throw FatalError("init (eigenPhysicsPackage)", "printSource must be 0 (No printing), 1 (ASCII) or 2 (BINARY)");
  }
  // fileSource%init (fileSource_class.f90:46-144): every row of the file is kept; broodID (column 9) is ignored
  void initFileSource(const Dict& d, bool dataIsMG) {
    std::string energy = d.getWord("data", "ce");
    if (energy != "ce" && energy != "mg") throw FatalError("init (fileSource)", "Invalid source data type specified: must be ce or mg");
    fsrc.isMG = (energy == "mg");
    if (!d.isPresent("path")) throw FatalError("init (fileSource)", "path must be specified in the dictionary for fileSource");
    std::string path = d.getWord("path");
    bool binary = d.getBool("binary", false);
    FILE* f = fopen(path.c_str(), binary ? "rb" : "r");
    if (!f) throw FatalError("init (fileSource)", "cannot open " + path);
    double row[10];
    for (;;) {
      bool ok = true;
      if (binary) ok = fread(row, sizeof(double), 10, f) == 10;
      else for (int k = 0; k < 10 && ok; ++k) ok = fscanf(f, "%lf", &row[k]) == 1;
      if (!ok) break;
      fsrc.rows.insert(fsrc.rows.end(), row, row + 10);
    }
    fclose(f);
    fsrc.N = (long)(fsrc.rows.size() / 10);
    fsrc.on = true;
    if (fsrc.isMG != dataIsMG) throw FatalError("init (fileSource)", "source data type inconsistent with nuclear database");
  }
  // fileSource%sampleParticle (:151-196)
  ParticleState sampleFile(RNG& rand) const {
    long idx = (long)(rand.get() * (double)fsrc.N) + 1;
    if (idx > fsrc.N) throw FatalError("sampleParticle (fileSource)", "Requested neutron is not in the file source file");
    const double* row = &fsrc.rows[10 * (size_t)(idx - 1)];
    ParticleState p;
    for (int k = 0; k < 3; ++k) { p.r[k] = row[k]; p.dir[k] = row[3 + k]; }
    int m, uid; geom.whatIsAt(m, uid, p.r);
    if (m == OUTSIDE_MAT || m == UNDEF_MAT) throw FatalError("sampleParticle (fileSource)", "Neutron sampled from file source is outside of geometry or in undefined region.");
    p.time = 0.0; p.wgt = row[9];
    if (fsrc.isMG) { p.G = (int)row[7]; p.isMG = true; } else { p.E = row[6]; p.isMG = false; }
    return p;
  }
  // materialSource%init (:75-134) and %sampleParticle (:136-210)
  void initMaterialSource(const Dict& d, bool dataIsMG, const std::map<std::string, int>& mats) {
    std::string energy = d.getWord("data", "ce");
    if (energy != "ce" && energy != "mg") throw FatalError("init (materialSource)", "Invalid source data type specified: must be ce or mg");
    msrc.isMG = (energy == "mg");
    if (msrc.isMG != dataIsMG) throw FatalError("init (materialSource)", "source data type does not match the nuclear database");
    msrc.E = d.getReal("E", 1.0E-6); msrc.G = d.getInt("G", 1);
    auto it = mats.find(d.getWord("mat"));
    if (it == mats.end()) throw FatalError("init (materialSource)", "Source material " + d.getWord("mat") + " was not found in the material definitions");
    msrc.matIdx = it->second;
    if (d.isPresent("boundingBox")) {
      auto b = d.getRealArray("boundingBox");
      if (b.size() != 6) throw FatalError("init (materialSource)", "Bounding box must have 6 entries");
      for (int k = 0; k < 3; ++k) { msrc.bottom[k] = b[k]; msrc.top[k] = b[3 + k]; }
    } else {
      double b[6]; geom.bounds(b);
      for (int k = 0; k < 3; ++k) { msrc.bottom[k] = b[k]; msrc.top[k] = b[3 + k]; }
    }
    msrc.on = true;
  }
  ParticleState sampleMaterial(RNG& rand) const {
    for (int i = 1; i <= 200; ++i) {
      double r3[3] = {0, 0, 0};
      r3[0] = rand.get(); r3[1] = rand.get(); r3[2] = rand.get();
      Vec3 r;
      for (int k = 0; k < 3; ++k) r[k] = (msrc.top[k] - msrc.bottom[k]) * r3[k] + msrc.bottom[k];
      (void)rand.get();                                              // time
      int m, uid; geom.whatIsAt(m, uid, r);
      if (m == OUTSIDE_MAT) continue;
      if (m == VOID_MAT || m == UNDEF_MAT || m == OVERLAP_MAT) throw FatalError("sampleParticle (materialSource)", "Nuclear data did not return neutron material.");
      if (m != msrc.matIdx) continue;
      ParticleState p;
      p.r = r; p.wgt = 1.0; p.time = 0.0;
      double mu = 2.0 * rand.get() - 1.0;
      double phi = TWO_PI * rand.get();
      Vec3 ex; ex[0] = 1.0;
      p.dir = rotateVector(ex, mu, phi);
      if (msrc.isMG) { p.G = msrc.G; p.isMG = true; } else { p.E = msrc.E; p.isMG = false; }
      return p;
    }
    throw FatalError("sampleParticle (materialSource)", "Infinite loop in sampling source. Please check that defined volume contains source material.");
  }
  ParticleState sampleSource(RNG& rand) const { return fsrc.on ? sampleFile(rand) : msrc.on ? sampleMaterial(rand) : samplePoint(rand); }
  std::map<std::string, int> sourceMats;                             // material names, set by init() before initSource
  void initSource(const Dict& d, int nG) {
    if (d.getWord("type") == "fileSource") initFileSource(d, nG > 0);
    else if (d.getWord("type") == "materialSource") initMaterialSource(d, nG > 0, sourceMats);
    else initPointSource(d, nG);
  }
  // pointSource%init (ParticleObjects/Source/pointSource_class.f90:60-140)
  void initPointSource(const Dict& d, int nG) {
    if (d.getWord("type") != "pointSource") throw FatalError("new_source", "oracle supports pointSource and fileSource for fixed-source calculations");
    if (d.getWord("particle", "neutron") != "neutron") throw FatalError("init (pointSource)", "oracle supports neutrons only");
    auto rr = d.getRealArray("r");
    if (rr.size() != 3) throw FatalError("init (pointSource)", "Source position must have three components");
    for (int k = 0; k < 3; ++k) psrc.r[k] = rr[k];
    int m, uid; geom.whatIsAt(m, uid, psrc.r);
    if (m == OUTSIDE_MAT) throw FatalError("init (pointSource)", "Source has been placed outside geometry");
    psrc.isotropic = !d.isPresent("dir");
    if (!psrc.isotropic) {
      auto dd = d.getRealArray("dir");
      if (dd.size() != 3) throw FatalError("init (pointSource)", "Source direction must have three components");
      double n = std::sqrt(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2]);
      for (int k = 0; k < 3; ++k) psrc.dir[k] = dd[k] / n;
    }
    bool isCE = d.isPresent("E"), isMGs = d.isPresent("G") || d.isPresent("probG");
    if (isCE && isMGs) throw FatalError("init (pointSource)", "Source may be either continuous energy or MG, not both");
    if (isCE) { psrc.E = d.getReal("E"); psrc.isMG = false; }
    else if (isMGs) {
      if (d.isPresent("probG") && d.isPresent("G")) throw FatalError("init (pointSource)", "Source may be either monoenergetic or a distribution, not both");
      if (d.isPresent("probG")) {
        psrc.probG = d.getRealArray("probG");
        if ((int)psrc.probG.size() != nG) throw FatalError("init (pointSource)", "Source energy group distribution has the wrong number of groups");
        double S = 0.0; for (double v : psrc.probG) S += v;
        for (double& v : psrc.probG) v = v / S;
      } else psrc.G = d.getInt("G");
      psrc.isMG = true;
    } else throw FatalError("init (pointSource)", "Must specify source energy, either Energy(E) or Group distribution (probG)");
  }
  // configSource%sampleParticle (configSource_inter.f90:75-90) with the pointSource procedures
  ParticleState samplePoint(RNG& rand) const {
    ParticleState p;
    p.r = psrc.r;
    if (psrc.isotropic) {
      double mu = 2.0 * rand.get() - 1.0;
      double phi = TWO_PI * rand.get();
      Vec3 ex; ex[0] = 1.0;
      p.dir = rotateVector(ex, mu, phi);
    } else p.dir = psrc.dir;
    if (psrc.isMG) {
      if (!psrc.probG.empty()) {
        double r = rand.get(); int g;
        for (g = 1; g <= (int)psrc.probG.size(); ++g) { r = r - psrc.probG[g - 1]; if (r < 0.0) break; }
        p.G = g;
      } else p.G = psrc.G;
      p.isMG = true;
    } else { p.E = psrc.E; p.isMG = false; }
    p.time = 0.0; p.wgt = 1.0;
    return p;
  }

  virtual void init(const Dict& dict, const std::string& baseDir) {
    initCycles(dict);
    std::string nucData = dict.getWord("XSdata");
    std::string energy = dict.getWord("dataType");
    if (energy != "mg") throw FatalError("init (eigenPhysicsPackage)", "oracle MG driver: dataType must be 'mg'");
    if (!dict.isPresent("seed")) throw FatalError("init (eigenPhysicsPackage)", "oracle requires an explicit seed");
    pRNG.init((int64_t)dict.getInt("seed"));
    keff_0 = dict.getReal("keff_0", 1.0);
    const Dict& nd = dict.getDict("nuclearData");
    auto mats = MgDatabase::materialMenu(nd);
    geom.init(dict.getDict("geometry"), mats);
    db.init(nd, nucData, baseDir);
    db.activate(geom.activeMats());
    view.db = &db;
    const Dict& co = dict.getDict("collisionOperator");
    if (!co.isPresent("neutronMG") || co.getDict("neutronMG").getWord("type") != "neutronMGstd")
      throw FatalError("collisionOperator init", "oracle supports neutronMGstd only");
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
      initSource(dict.getDict("source"), db.nG);
      return;
    }
    inactiveTally.init(dict.getDict("inactiveTally"), mats);
    activeTally.init(dict.getDict("activeTally"), mats);
    if (dict.isPresent("source")) throw FatalError("init (eigenPhysicsPackage)", "oracle supports the default fissionSource only");
    source.init(&geom, &db, xs);
    inactiveAtch.init(Dict::fromString("keff { type keffAnalogClerk; } display (keff); mpiSync 1;"), mats);
    activeAtch.init(Dict::fromString("keff { type keffImplicitClerk; } display (keff); mpiSync 1;"), mats);
    inactiveTally.atch = &inactiveAtch;
    activeTally.atch = &activeAtch;
  }

  // one source batch: fixedSourcePhysicsPackage_class.f90:149-289 (private buffer only, no common buffer)
  void fixedCycle() {
    TallyAdmin& tally = activeTally;
    thisCycle = &dungeonA;
    if ((int)dungeonA.prisoners.size() < pop) dungeonA.init(pop);
    // source%generate (source_inter.f90:98-118)
    dungeonA.setSize(pop);
    std::string srcErr;                                              // an exception may not leave the parallel region
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= pop; ++i) {
      RNG pRand = pRNG; pRand.stride(i);
      try { dungeonA.prisoners[i - 1] = sampleSource(pRand); }
      catch (const std::exception& ex) {
#pragma omp critical
        srcErr = ex.what();
      }
    }
    if (!srcErr.empty()) throw FatalError("generate (source)", srcErr);
    pRNG.stride(pop);
    cycleInPhase[1] += 1;                                            // fixedSourcePhysicsPackage_class.f90:191-194
    if (printSource != 0) dungeonA.printToFile(outputFile + "_source" + std::to_string(cycleInPhase[1]), printSource == 2);
    tally.reportCycleStart(*thisCycle);
    long seg = 0, coll = 0;
#pragma omp parallel reduction(+ : seg, coll)
    {
      Dungeon buffer; buffer.init(bufferSize);
#pragma omp for schedule(dynamic)
      for (int n = 1; n <= pop; ++n) {
        RNG rng = pRNG; rng.stride(n);
        Particle p; p.pRNG = &rng; p.k_eff = 1.0;
        p.fromState(thisCycle->prisoners[n - 1]);
        p.isDead = false;
        for (;;) {                                                   // bufferLoop
          geom.placeCoord(p.coords);
          p.preHistory = p.state();
          p.preCollision = p.state();
          double trackXS = 0.0;
          for (;;) {
            transport(p, tally, trackXS, seg);
            if (p.isDead) break;
            collide(p, tally, trackXS, buffer);
            ++coll;
            if (p.isDead) break;
          }
          if (buffer.pop == 0) break;
          ParticleState st = buffer.prisoners[buffer.pop - 1];       // release: the last one detained
          buffer.pop -= 1;
          p.fromState(st);
          p.isDead = false;
        }
      }
    }
    nSegments += seg; nCollisions += coll; nHistories += pop;
    pRNG.stride(pop);
    tally.reportCycleEnd(*thisCycle);
  }
  void runFixed() { for (int i = 0; i < N_cycles; ++i) fixedCycle(); }

  // -------------------------------------------------------------------------
  // transport operators; return with p.isDead or at a real collision site.
  // trackXS (out) = content of the tracking cache for the tallies.
  // -------------------------------------------------------------------------
  void deltaTracking(Particle& p, TallyAdmin& tally, double& trackXS, long& seg) const {   // transportOperatorDT_class.f90:47-130
    trackXS = std::max(xs->majorantXS(p), xs->collisionXS());
    double majorant_inv = 1.0 / trackXS;
    for (;;) {
      double distance = -mlog(p.pRNG->get()) * majorant_inv;
      geom.teleport(p.coords, distance);
      p.time = p.time + distance / 1.0;
      ++seg;
      int m = p.matIdx();
      if (m == OUTSIDE_MAT) { p.fate = LEAK_FATE; p.isDead = true; break; }
      if (m == VOID_MAT) { tally.reportInColl(p, *xs, trackXS, true); continue; }
      if (m == UNDEF_MAT) throw FatalError("deltaTracking", "Particle is in undefined material");
      if (m == OVERLAP_MAT) throw FatalError("deltaTracking", "Particle is in overlapping cells");
      double sigmaT = xs->trackMatXS(p, m);
      if (p.pRNG->get() < sigmaT * majorant_inv) break;
      tally.reportInColl(p, *xs, trackXS, true);
    }
  }
  void surfaceTracking(Particle& p, TallyAdmin& tally, double& trackXS, long& seg) const {  // transportOperatorHT_class.f90:156-258 (== ST_class.f90:48-166)
    const double tol = 1.0E-12;
    DistCache cache;
    for (;;) {
      int m = p.matIdx();
      double sigmaTrack = (m == VOID_MAT) ? xs->collisionXS() : std::max(xs->trackMatXS(p, m), xs->collisionXS());
      trackXS = sigmaTrack;
      double dist, invSigmaTrack, sigmaT;
      if (sigmaTrack < tol) { dist = INF; invSigmaTrack = INF; sigmaT = 0.0; }
      else {
        invSigmaTrack = 1.0 / sigmaTrack;
        dist = -mlog(p.pRNG->get()) * invSigmaTrack;
        sigmaT = xs->trackMatXS(p, m);
        if (dist != dist) throw FatalError("surfaceTracking", "Distance is NaN");
      }
      p.prePath = p.state();
      int event;
      geom.move(p.coords, dist, event, stCache ? &cache : nullptr);
      p.time = p.time + dist / 1.0;
      ++seg;
      if (event == COLL_EV) p.fate = NO_FATE;
      tally.reportPath(p, *xs, dist);                                // transportOperatorST_class.f90:125
      m = p.matIdx();
      if (m == OUTSIDE_MAT) { p.isDead = true; p.fate = LEAK_FATE; }
      else if (m == UNDEF_MAT) throw FatalError("surfaceTracking", "Particle is in undefined material");
      else if (m == OVERLAP_MAT) throw FatalError("surfaceTracking", "Particle is in overlapping cells");
      if (p.isDead) break;
      if (event == COLL_EV) {
        if (p.pRNG->get() < sigmaT * invSigmaTrack) break;
        tally.reportInColl(p, *xs, trackXS, true);
      }
    }
  }
  void transport(Particle& p, TallyAdmin& tally, double& trackXS, long& seg) const {       // transportOperator_inter.f90:105-130
    p.preTransition = p.state();
    if (tracking == TRACK_DT) deltaTracking(p, tally, trackXS, seg);
    else if (tracking == TRACK_ST) surfaceTracking(p, tally, trackXS, seg);
    else {                                                          // transportOperatorHT_class.f90:49-81
      double majorant_inv = 1.0 / std::max(xs->majorantXS(p), xs->collisionXS());
      double sigmaT = (p.matIdx() == VOID_MAT) ? 0.0 : xs->trackMatXS(p, p.matIdx());
      double ratio = sigmaT * majorant_inv;
      if (ratio > (1.0 - htCutoff)) deltaTracking(p, tally, trackXS, seg);
      else surfaceTracking(p, tally, trackXS, seg);
    }
    if (p.isDead) tally.reportHist(p);
  }

  // collisionProcessor_inter.f90:114-195 with the neutronMGstd hooks (neutronMGstd_class.f90:85-297)
  virtual void collide(Particle& p, TallyAdmin& tally, double trackXS, Dungeon& next) const {
    int matIdx = p.matIdx();
    // sampleCollision: alpha-absorption test always draws (alpha = 0 => never taken)
    double denom = db.getTrackMatXS(p.G, matIdx);
    double probAlpha = 0.0 / denom;
    int MT;
    const MgMaterial& mat = db.mats.at(matIdx - 1);
    if (p.pRNG->get() < probAlpha) MT = 0;   // unreachable for alpha = 0
    else {
      MacroXSs x; mat.getMacroXSs(x, p.G);
      double r = p.pRNG->get();
      MT = x.invert(r);
    }
    tally.reportInColl(p, view, trackXS, false);
    p.preCollision = p.state();
    // implicit
    if (mat.fissile) {
      double wgt = p.w, w0 = p.preHistory.wgt, k_eff = p.k_eff;
      double rand1 = p.pRNG->get();
      MacroXSs x; mat.getMacroXSs(x, p.G);
      double sig_tot = x.total, sig_nuFiss = x.nuFission;
      int n = (int)(std::fabs((wgt * sig_nuFiss) / (w0 * sig_tot * k_eff)) + rand1);
      if (n >= 1) {
        wgt = fsign(w0, wgt);
        Vec3 r = p.coords.lvl[0].r;
        for (int i = 0; i < n; ++i) {
          double mu, phi; int G_out;
          mat.fissionSampleOut(mu, phi, G_out, *p.pRNG);
          Vec3 dir = rotateVector(p.coords.lvl[0].dir, mu, phi);
          ParticleState t = p.state();
          t.r = r; t.dir = dir; t.G = G_out; t.wgt = wgt; t.collisionN = 0;
          next.detain(t);
        }
      }
    }
    switch (MT) {
      case macroIEscatter: {                                        // inelastic :221-252
        double mu, phi; int G_out;
        mat.scatterSampleOut(mu, phi, G_out, p.G, *p.pRNG);
        double w_mul = mat.production(p.G, G_out);
        p.G = G_out;
        p.w = p.w * w_mul;
        p.coords.rotate(mu, phi);
        break;
      }
      case macroDisappearance: case macroFission: p.isDead = true; break;
      case macroEscatter: case macroAllScatter: break;              // elastic: "Do nothing. Should not be called"
      default: throw FatalError("collide", "Unsupported MT number");
    }
    p.collisionN += 1;
    tally.reportOutColl(p, MT);
    if (p.isDead) { p.fate = ABS_FATE; tally.reportHist(p); }
  }

  // eigenPhysicsPackage_class.f90:348-366
  void generateInitialState() {
    dungeonA.init(2 * pop); dungeonB.init(2 * pop);
    thisCycle = &dungeonA; nextCycle = &dungeonB;
    source.generate(*thisCycle, pop, pRNG);
    pRNG.stride(pop);
  }

  // one history: eigenPhysicsPackage_class.f90:213-252
  void history(int n, double k_new, TallyAdmin& tally, long& seg, long& coll) {
    RNG rng = pRNG;
    rng.stride(n);
    Particle neutron;
    neutron.pRNG = &rng;
    neutron.fromState(thisCycle->prisoners[n - 1]);
    neutron.isDead = false;
    neutron.broodID = n;
    geom.placeCoord(neutron.coords);
    neutron.k_eff = k_new;
    neutron.preHistory = neutron.state();
    neutron.preCollision = neutron.state();
    double trackXS = 0.0;
    for (;;) {
      transport(neutron, tally, trackXS, seg);
      if (neutron.isDead) break;
      collide(neutron, tally, trackXS, *nextCycle);
      ++coll;
      if (neutron.isDead) break;
    }
  }

  // one cycle: eigenPhysicsPackage_class.f90:203-307 ; returns k_new
  double cycle(bool active, double k_new) {
    TallyAdmin& tally = active ? activeTally : inactiveTally;
    TallyAdmin& atchT = active ? activeAtch : inactiveAtch;
    tally.reportCycleStart(*thisCycle);
    int nParticles = thisCycle->pop;
    long seg = 0, coll = 0;
    std::string histErr;                                             // fatalError inside a history: stop after the loop (an exception may not leave the parallel region)
#pragma omp parallel for schedule(dynamic) reduction(+ : seg, coll)
    for (int n = 1; n <= nParticles; ++n) {
      if (!histErr.empty()) continue;
      try { history(n, k_new, tally, seg, coll); }
      catch (const std::exception& ex) {
#pragma omp critical
        histErr = std::string(ex.what()) + " [history " + std::to_string(n) + "]";
      }
    }
    if (!histErr.empty()) throw FatalError("cycles", histErr);
    nSegments += seg; nCollisions += coll; nHistories += nParticles;
    thisCycle->pop = 0;                                             // cleanPop
    pRNG.stride(pop + 1);
    tally.reportCycleEnd(*nextCycle);
    nextCycle->normSize_Repr(pop, pRNG);
    pRNG.stride(1);
    cycleInPhase[active ? 1 : 0] += 1;                               // the `i` of the cycles loop restarts with each phase
    if (printSource != 0)
      nextCycle->printToFile(outputFile + "_source" + std::to_string(cycleInPhase[active ? 1 : 0]) + "_rank0", printSource == 2);
    std::swap(thisCycle, nextCycle);
    double k, s;
    atchT.getKeff(k, s);
    nextCycle->k_eff = k;
    keff_0 = k;
    cycleK.push_back(k);
    return k;
  }

  // eigenPhysicsPackage_class.f90:135-159,164-343
  void runCycles(bool active, int N) {
    double k_new = keff_0;
    for (int i = 0; i < N; ++i) k_new = cycle(active, k_new);
  }
  void run() {
    if (fixedSource) { runFixed(); return; }
    generateInitialState();
    runCycles(false, N_inactive);
    runCycles(true, N_active);
  }
};

}  // namespace orc
