// Host-side mirror of SCONE's eigenPhysicsPackage, driving the engine through the C ABI only.
//   PhysicsPackages/eigenPhysicsPackage_class.f90:135-159  run
//   PhysicsPackages/eigenPhysicsPackage_class.f90:164-343  cycles
//   PhysicsPackages/eigenPhysicsPackage_class.f90:348-366  generateInitialState
//   PhysicsPackages/eigenPhysicsPackage_class.f90:417-645  init
// The Fortran package would keep this control flow and call sb_* through iso_c_binding
// (INTEGRATION.md); there is no Fortran compiler in the build image, so the same sequence of
// calls is written here in C++ and exported with a small C API (sbh_*) for bench.py / tests.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/scone_b200.h"
#include "../sb_rng.h"
#include "model.hpp"
#include <sstream>
#include "ce_model.hpp"
#include "../sb_cekin.cuh"

namespace {

struct eigenPhysicsPackage {
  // settings
  int pop = 0, totalPop = 0, N_inactive = 0, N_active = 0;
  sb_options opt{};
  double* hBinsP = nullptr; size_t hBinsCap = 0;          // page-locked copy of the cycle's BIN column (host-buffer cycles)
  double keff_0 = 1.0;
  uint64_t pRNG = 0, masterRNG = 0;      // masterRNG: the pRNG of rank 0 (normSize_Repr draws from the master's stream)
  int rank = 0, nRanks = 1;
  int rankOffset = 0;                      // getOffset(totalPop) of this rank (mpi_func.f90:133-159)
  sb::FlatGeometry geom; sb::FlatMgData data; sb::FlatCeData ceData; bool isCE = false; sb::TallyDefs tallies[2];
  // fixedSourcePhysicsPackage (PhysicsPackages/fixedSourcePhysicsPackage_class.f90): cycles, buffer, pointSource
  bool isFixed = false; int N_cycles = 0, bufferSize = 50; sb_point_source psrc{}; std::vector<double> probG;
  // fileSource (ParticleObjects/Source/fileSource_class.f90): the rows of a printToFile dump
  bool isFileSrc = false, fileSrcMG = false; std::vector<double> fileRows;
  bool isMatSrc = false, matSrcBox = false; sb_material_source msrc{};
  bool fixedStarted = false;
  // printSource / outputFile (eigenPhysicsPackage_class.f90:278-281,463,501-504)
  int printSource = 0; std::string outputFile = "./output"; int cycleInPhase[2] = {0, 0}, lastActive = 0; std::vector<int32_t> hBrood;
  sb_engine* eng = nullptr;
  std::string err;
  // results
  std::vector<double> cycleK; sb_cycle_result last{};
  long long nSegActive = 0, nSegInactive = 0, nHist = 0;
  double timeTransport = 0.0;
  // host copy of the bank for the end-to-end (host buffer) mode
  double *hr = nullptr, *hdir = nullptr, *hw = nullptr, *hE = nullptr; int32_t* hG = nullptr; int hN = 0, hCap = 0;   // page-locked
  std::vector<double> hBins;

  static std::string dirName(const std::string& p) { size_t k = p.rfind('/'); return k == std::string::npos ? "." : p.substr(0, k); }

  void stride(int64_t n) { pRNG = sbd::rng_skip(pRNG, sbd::RNG_STRIDE * n); masterRNG = sbd::rng_skip(masterRNG, sbd::RNG_STRIDE * n); }

  int fail(const std::string& m) { err = m; return -1; }
  int engFail() { err = sb_last_error(eng); return -1; }

  int init(const std::string& deckPath, const char* overrides, int device, int rank_, int nRanks_) {
    rank = rank_; nRanks = nRanks_;
    try {
      sb::Dict dict = sb::Dict::fromFile(deckPath);
      if (overrides && *overrides) {
        sb::Dict ov = sb::Dict::fromString(overrides);
        auto dk = ov.keys("dict");
        for (auto& k : ov.keys("all")) {
          if (std::find(dk.begin(), dk.end(), k) != dk.end()) dict.setDict(k, ov.getDict(k));
          else dict.setScalar(k, ov.getWord(k));
        }
      }
      const std::string ppType = dict.getWord("type");
      if (ppType != "eigenPhysicsPackage" && ppType != "fixedSourcePhysicsPackage") return fail("eigenPhysicsPackage and fixedSourcePhysicsPackage decks are driven by this host");
      isFixed = (ppType == "fixedSourcePhysicsPackage");
      totalPop = dict.getInt("pop");
      // getWorkshare / getOffset (mpi_func.f90:133-159): contiguous shares, the remainder goes to the high ranks
      pop = (totalPop + rank) / nRanks;
      rankOffset = totalPop / nRanks * rank + std::max(0, totalPop % nRanks + rank - nRanks);
      if (isFixed) { N_cycles = dict.getInt("cycles"); bufferSize = dict.getInt("buffer", 50); N_inactive = 0; N_active = N_cycles; }
      else { N_inactive = dict.getInt("inactive"); N_active = dict.getInt("active"); }
      // options of the reference packages that change the physics and are not on the device: refuse them rather than ignore them
      if (!dict.getBool("reproducible", true)) return fail("reproducible 0 (normSize without the reproducible resampling) is not supported: the device bank is always resampled by normSize_Repr");
      for (const char* key : {"temperature", "density", "uniformFissionSites"})
        if (dict.isPresent(key)) return fail(std::string("superimposed field `") + key + "` is not supported by the device engine");
      outputFile = dict.getWord("outputFile", "./output");
      printSource = dict.getInt("printSource", 0);
      if (printSource < 0 || printSource > 2) return fail("printSource must be 0 (No printing), 1 (ASCII) or 2 (BINARY)");
      if (isFixed) for (const char* key : {"commonBufferSize", "varianceReduction"})
        if (dict.isPresent(key)) return fail(std::string("`") + key + "` of fixedSourcePhysicsPackage is not supported by the device engine");
      std::string nucData = dict.getWord("XSdata"), energy = dict.getWord("dataType");
      if (energy != "mg" && energy != "ce") return fail("dataType must be 'mg' or 'ce'");
      isCE = (energy == "ce");
      if (!dict.isPresent("seed")) return fail("an explicit `seed` is required for a reproducible run");
      pRNG = (uint64_t)(int64_t)dict.getInt("seed"); masterRNG = pRNG;
      keff_0 = dict.getReal("keff_0", 1.0);
      const sb::Dict& nd = dict.getDict("nuclearData");
      sb::MatMap mats = sb::materialMenu(nd);
      geom = sb::buildGeometry(dict.getDict("geometry"), mats);
      const sb::Dict& co = dict.getDict("collisionOperator");
      if (isCE) {
        if (!co.isPresent("neutronCE")) return fail("collisionOperator: neutronCE { type neutronCEstd; } is required for dataType ce");
        ceData = sb::buildCeData(nd, nucData, dirName(deckPath), geom.activeMats(), co);
        data.nMat = ceData.nMat; data.nG = 0;
      } else {
        data = sb::buildMgData(nd, nucData, dirName(deckPath), geom.activeMats());
        if (!co.isPresent("neutronMG") || co.getDict("neutronMG").getWord("type") != "neutronMGstd") return fail("collisionOperator: neutronMGstd is required");
      }
      opt = sb_options{}; opt.max_pop = pop; opt.ht_cutoff = 0.9; opt.st_cache = 1;
      const sb::Dict& to = dict.getDict("transportOperator");
      std::string tt = to.getWord("type");
      if (tt == "transportOperatorDT") opt.tracking = SB_TRACK_DT;
      else if (tt == "transportOperatorST") { opt.tracking = SB_TRACK_ST; opt.st_cache = to.getBool("cache", true) ? 1 : 0; }
      else if (tt == "transportOperatorHT") { opt.tracking = SB_TRACK_HT; opt.ht_cutoff = to.getReal("cutoff", 0.9); opt.st_cache = to.getBool("cache", true) ? 1 : 0; }
      else return fail("Unrecognised type of transportOperator: " + tt);
      if (isFixed) {
        tallies[0] = sb::buildTallies(sb::Dict::fromString(""), mats, data.nMat);
        tallies[1] = sb::buildTallies(dict.getDict("tally"), mats, data.nMat);
        for (auto& c : tallies[1].clerks) if (c.kind == SB_CLERK_SHANNON) return fail("shannonEntropyClerk in a fixed-source tally is not supported by the device tallies");
        const sb::Dict& sd = dict.getDict("source");
        const std::string st = sd.getWord("type");
        if (int rc = (st == "fileSource") ? initFileSource(sd) : (st == "materialSource") ? initMaterialSource(sd, mats) : initPointSource(sd)) return rc;
      } else {
        tallies[0] = sb::buildTallies(dict.getDict("inactiveTally"), mats, data.nMat);
        tallies[1] = sb::buildTallies(dict.getDict("activeTally"), mats, data.nMat);
        if (dict.isPresent("source")) return fail("only the default fissionSource is supported in eigenvalue calculations");
      }

      if (device < 0) return 0;     // host model only (CPU-side tests of the flattening); no engine, no transport
      if (sb_create(&eng, device)) return fail(sb_last_error(nullptr));
      sb_geom_flat gv = geom.view(); if (sb_load_geometry(eng, &gv)) return engFail();
      if (isCE) { sb_ce_model cv = ceData.view(); if (sb_load_ce_model(eng, &cv)) return engFail(); }
      else { sb_mg_flat dv = data.view(); if (sb_load_mg_data(eng, &dv)) return engFail(); }
      for (int ph = 0; ph < 2; ++ph)
        if (sb_define_tallies(eng, ph, tallies[ph].clerks.data(), (int)tallies[ph].clerks.size(), tallies[ph].normClerk, tallies[ph].normVal)) return engFail();
      if (sb_set_options(eng, &opt)) return engFail();
      if (isFixed && sb_set_fixed_source(eng, 1, bufferSize)) return engFail();
      if (isMatSrc && !matSrcBox) {                                   // bounds = self % geom % bounds()
        double b[6]; if (sb_geometry_bounds(eng, b)) return engFail();
        for (int k = 0; k < 3; ++k) { msrc.bottom[k] = b[k]; msrc.top[k] = b[3 + k]; }
      }
      if (isFileSrc && sb_set_file_source(eng, (int64_t)(fileRows.size() / 10), fileRows.data(), fileSrcMG ? 1 : 0)) return engFail();
    } catch (const std::exception& e) { return fail(e.what()); }
    return 0;
  }

  // materialSource%init (materialSource_class.f90:75-134); boundingTime only moves the particle's time, which the engine does not carry
  int initMaterialSource(const sb::Dict& d, const sb::MatMap& mats) {
    std::string energy = d.getWord("data", "ce");
    if (energy != "ce" && energy != "mg") return fail("init (materialSource): Invalid source data type specified: must be ce or mg");
    msrc.is_mg = (energy == "mg") ? 1 : 0;
    if ((msrc.is_mg != 0) == isCE) return fail("init (materialSource): the source data type does not match dataType");
    msrc.E = d.getReal("E", 1.0E-6); msrc.G = d.getInt("G", 1);
    const std::string name = d.getWord("mat");
    auto it = mats.find(name);
    if (it == mats.end()) return fail("init (materialSource): Source material " + name + " was not found in the material definitions");
    msrc.mat_idx = it->second;
    if (d.isPresent("boundingBox")) {
      auto b = d.getRealArray("boundingBox");
      if (b.size() != 6) return fail("init (materialSource): Bounding box must have 6 entries");
      for (int k = 0; k < 3; ++k) { msrc.bottom[k] = b[k]; msrc.top[k] = b[3 + k]; }
      matSrcBox = true;
    }
    if (d.isPresent("boundingTime")) {
      auto t = d.getRealArray("boundingTime");
      if (t.size() != 2) return fail("init (materialSource): Bounding time must have 2 entries");
      if (t[1] < t[0]) return fail("init (materialSource): tHigh is less than tLow");
    }
    isMatSrc = true;
    return 0;
  }
  // fileSource%init (fileSource_class.f90:46-144): all rows are kept, broodID (column 9) is ignored by the sampling
  int initFileSource(const sb::Dict& d) {
    std::string energy = d.getWord("data", "ce");
    if (energy != "ce" && energy != "mg") return fail("init (fileSource): Invalid source data type specified: must be ce or mg");
    fileSrcMG = (energy == "mg");
    if (fileSrcMG == isCE) return fail("init (fileSource): source data type inconsistent with nuclear database");
    if (!d.isPresent("path")) return fail("init (fileSource): path must be specified in the dictionary for fileSource");
    const std::string path = d.getWord("path");
    const bool binary = d.getBool("binary", false);
    FILE* f = fopen(path.c_str(), binary ? "rb" : "r");
    if (!f) return fail("init (fileSource): cannot open the source file " + path);
    double row[10];
    for (;;) {
      bool ok = true;
      if (binary) ok = fread(row, sizeof(double), 10, f) == 10;
      else for (int k = 0; k < 10 && ok; ++k) ok = fscanf(f, "%lf", &row[k]) == 1;
      if (!ok) break;
      fileRows.insert(fileRows.end(), row, row + 10);
    }
    fclose(f);
    if (fileRows.empty()) return fail("init (fileSource): the source file holds no particles: " + path);
    isFileSrc = true;
    return 0;
  }
  // particleDungeon%printToFile (particleDungeon_class.f90:1077-1112) for the bank the cycle has just normalised: r, dir, E, real(G),
  // real(broodID), wgt per site; stream binary (.bin) or one text row per site (.txt, 17 significant digits: reads back exactly)
  // deferred = the caller still has to balance the banks of the ranks (loadBalancing is part of normSize_Repr): it prints afterwards
  int printBank(int active, bool deferred = false, bool fixedBatch = false) {
    if (!deferred || nRanks == 1) cycleInPhase[active ? 1 : 0] += 1;      // the `i` of the cycles loop restarts with each phase
    if (printSource == 0 || (deferred && nRanks > 1)) return 0;
    if (downloadBank()) return -1;
    hBrood.resize((size_t)std::max(1, hN));
    if (sb_bank_brood(eng, (int)hBrood.size(), hBrood.data())) return engFail();
    const bool bin = (printSource == 2);
    const int i = cycleInPhase[active ? 1 : 0];
    // eigenvalue: <outputFile>_source<i>_rank<r> (eigenPhysicsPackage_class.f90:279); fixed source: <outputFile>_source<i>, the batch the
    // source has just generated (fixedSourcePhysicsPackage_class.f90:191-194)
    const std::string name = outputFile + "_source" + std::to_string(i) + (fixedBatch ? std::string() : "_rank" + std::to_string(rank)) + (bin ? ".bin" : ".txt");
    FILE* f = fopen(name.c_str(), bin ? "wb" : "w");
    if (!f) return fail("printToFile: cannot open " + name);
    std::vector<double> rows(10 * (size_t)hN);
    for (int s = 0; s < hN; ++s) {
      double* row = &rows[10 * (size_t)s];
      for (int k = 0; k < 3; ++k) { row[k] = hr[3 * (size_t)s + k]; row[3 + k] = hdir[3 * (size_t)s + k]; }
      row[6] = isCE ? hE[s] : 0.0; row[7] = isCE ? 0.0 : (double)hG[s]; row[8] = (double)hBrood[s]; row[9] = hw[s];
    }
    if (bin) fwrite(rows.data(), sizeof(double), rows.size(), f);
    else for (int s = 0; s < hN; ++s) { for (int k = 0; k < 10; ++k) fprintf(f, "%s%.17g", k ? " " : "  ", rows[10 * (size_t)s + k]); fprintf(f, "\n"); }
    fclose(f);
    return 0;
  }
  // pointSource%init (ParticleObjects/Source/pointSource_class.f90:60-140); the OUTSIDE check of the position is the engine's
  int initPointSource(const sb::Dict& d) {
    if (d.getWord("type") != "pointSource") return fail("fixed-source calculations: only pointSource is supported");
    if (d.getWord("particle", "neutron") != "neutron") return fail("init (pointSource): only neutrons are supported");
    auto rr = d.getRealArray("r");
    if (rr.size() != 3) return fail("init (pointSource): Source position must have three components");
    for (int k = 0; k < 3; ++k) psrc.r[k] = rr[k];
    psrc.isotropic = d.isPresent("dir") ? 0 : 1;
    psrc.dir[0] = 1.0; psrc.dir[1] = 0.0; psrc.dir[2] = 0.0;
    if (!psrc.isotropic) {
      auto dd = d.getRealArray("dir");
      if (dd.size() != 3) return fail("init (pointSource): Source direction must have three components");
      double n = std::sqrt(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2]);
      for (int k = 0; k < 3; ++k) psrc.dir[k] = dd[k] / n;
    }
    const bool srcCE = d.isPresent("E"), srcMG = d.isPresent("G") || d.isPresent("probG");
    if (srcCE && srcMG) return fail("init (pointSource): Source may be either continuous energy or MG, not both");
    if (!srcCE && !srcMG) return fail("init (pointSource): Must specify source energy, either Energy(E) or Group distribution (probG)");
    psrc.is_mg = srcMG ? 1 : 0; psrc.E = 0.0; psrc.G = 1; psrc.n_prob = 0; psrc.prob_g = nullptr;
    if (srcCE) psrc.E = d.getReal("E");
    else if (d.isPresent("probG")) {
      if (d.isPresent("G")) return fail("init (pointSource): Source may be either monoenergetic or a distribution, not both");
      probG = d.getRealArray("probG");
      double S = 0.0; for (double v : probG) S += v;
      for (double& v : probG) v = v / S;
      psrc.n_prob = (int)probG.size(); psrc.prob_g = probG.data();
    } else psrc.G = d.getInt("G");
    if ((psrc.is_mg != 0) == isCE) return fail("init (pointSource): the source energy type does not match dataType");
    return 0;
  }

  // one source batch (fixedSourcePhysicsPackage_class.f90:168-268): generate, stride, histories with their secondaries, stride, reportCycleEnd
  int fixedCycle() {
    if (!eng) return fail("no engine: this handle was created without a device");
    if (!isFixed) return fail("not a fixedSourcePhysicsPackage deck");
    // several ranks (fixedSourcePhysicsPackage_class.f90:131,347): every rank owns getWorkshare(totalPop) source particles of each batch and
    // starts from the package RNG strided by getOffset(totalPop); no exchange during the run, tally%collectDistributed at the end
    if (!fixedStarted) { pRNG = sbd::rng_skip(pRNG, sbd::RNG_STRIDE * (int64_t)rankOffset); fixedStarted = true; }
    if (isFileSrc ? sb_source_file(eng, pop, pRNG, 0) : isMatSrc ? sb_source_material(eng, pop, pRNG, 0, &msrc) : sb_source_point(eng, pop, pRNG, 0, &psrc)) return engFail();
    stride(totalPop);
    if (printBank(1, false, true)) return -1;
    if (sb_run_cycle(eng, pRNG, 0, 1.0, 1, &last)) return engFail();
    stride(totalPop);
    nSegActive += last.n_segments; nHist += last.n_start;
    return 0;
  }

  // run(): pRNG%stride(getOffset(totalPop)) then generateInitialState
  int generateInitialState() {
    if (!eng) return fail("no engine: this handle was created without a device");
    pRNG = sbd::rng_skip(pRNG, sbd::RNG_STRIDE * (int64_t)rankOffset);      // self%pRNG%stride(getOffset(totalPop)), :142
    if (sb_source_generate(eng, pop, pRNG, 0)) return engFail();
    stride(totalPop);
    return 0;
  }

  // one pass of the cycle body (eigenPhysicsPackage_class.f90:203-307) with the bank resident on the device
  int cycle(int active, double& k_new) {
    if (!eng) return fail("no engine: this handle was created without a device");
    auto t0 = std::chrono::steady_clock::now();
    if (nRanks > 1) return fail("this package owns a share of the bank: use the cycleBegin / cycleEnd / resampleRanked steps");
    const uint64_t rng0 = pRNG;
    stride(totalPop + 1);
    if (sb_run_cycle_resample(eng, rng0, 0, k_new, active, pop, pRNG, &last)) return engFail();      // one synchronisation per cycle
    stride(1);
    if (printBank(active)) return -1;
    k_new = last.k_cum;
    keff_0 = k_new;
    cycleK.push_back(k_new);
    (active ? nSegActive : nSegInactive) += last.n_segments;
    nHist += last.n_start;
    timeTransport += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
  }

  // ---- the cycle in three steps for several ranks; the caller moves the data between ranks ----
  // 1. transport; this rank's k-eff score sums land in dev_sums (device, 6 doubles) for the all-reduce
  int cycleBegin(int active, double k_new, double* devSums, int32_t* nSites) {
    if (!eng) return fail("no engine: this handle was created without a device");
    if (sb_cycle_begin(eng, pRNG, 0, k_new, active, devSums, nSites)) return engFail();
    return 0;
  }
  // 2. cycle close from the reduced sums; pRNG%stride(totalPop + 1)
  int cycleEnd(int active, const double* devSums) {
    if (sb_cycle_end(eng, devSums, &last)) return engFail();
    stride(totalPop + 1);
    lastActive = active;
    (active ? nSegActive : nSegInactive) += last.n_segments;
    nHist += last.n_start;
    return 0;
  }
  // 3. normSize_Repr over all ranks' banks (sizes gathered by the caller); pRNG%stride(1); k_new
  int resampleRanked(const int32_t* popSizes, int32_t* newLocal, double& k_new) {
    if (sb_resample_ranked(eng, totalPop, masterRNG, nRanks, rank, popSizes, newLocal)) return engFail();
    stride(1);
    if (printBank(lastActive, true)) return -1;
    k_new = last.k_cum; keff_0 = k_new; cycleK.push_back(k_new);
    return 0;
  }

  // steps 2 + 3 with one synchronisation (host-side reduced sums, every rank's new size computed locally)
  int cycleEndResampleRanked(int active, const double* hostSums, const int32_t* popSizes, int32_t* newSizes, double& k_new) {
    stride(totalPop + 1);
    if (sb_cycle_end_resample_ranked(eng, hostSums, totalPop, masterRNG, nRanks, rank, popSizes, newSizes, &last)) return engFail();
    stride(1);
    if (printBank(active, true)) return -1;
    (active ? nSegActive : nSegInactive) += last.n_segments;
    nHist += last.n_start;
    k_new = last.k_cum; keff_0 = k_new; cycleK.push_back(k_new);
    return 0;
  }

  // the whole cycle of one rank with the other ranks reached through peer memory (sb_peer_*): no collective, one synchronisation
  int cyclePeer(int active, int32_t* finalSizes, double& k_new) {
    if (!eng) return fail("no engine: this handle was created without a device");
    const uint64_t rng0 = pRNG;
    stride(totalPop + 1);
    if (sb_run_cycle_ranked_peer(eng, rng0, 0, k_new, active, totalPop, masterRNG, finalSizes, &last)) return engFail();
    stride(1);
    (active ? nSegActive : nSegInactive) += last.n_segments;
    nHist += last.n_start;
    k_new = last.k_cum; keff_0 = k_new; cycleK.push_back(k_new);
    return printBank(active);
  }

  // the same cycle with the dungeons kept in HOST memory, as a shim that leaves thisCycle/nextCycle in
  // Fortran arrays would do: upload bank, run, resample, download bank, read the cycle's bins
  int cycleHostBuffers(int active, double& k_new) {
    if (!eng) return fail("no engine: this handle was created without a device");
    if (hN == 0) { if (downloadBank()) return -1; }
    const uint64_t rng0 = pRNG;
    stride(totalPop + 1);
    int64_t nb = sb_tally_size(eng, active);
    if ((int64_t)hBinsCap < std::max<int64_t>(1, nb)) { sb_pinned_free(hBinsP); hBinsCap = (size_t)std::max<int64_t>(1, nb); hBinsP = (double*)sb_pinned_alloc(sizeof(double) * hBinsCap); if (!hBinsP) return fail("pinned host allocation failed"); }
    // upload, cycle, normSize_Repr, read-back of the bank and of the BIN column: one call, one synchronisation
    if (sb_run_cycle_resample_host(eng, hN, hr, hdir, hw, isCE ? nullptr : hG, isCE ? hE : nullptr, rng0, 0, k_new, active, pop, pRNG,
                                   &hN, hr, hdir, hw, isCE ? nullptr : hG, isCE ? hE : nullptr, nb > 0 ? hBinsP : nullptr, &last)) return engFail();
    stride(1);
    if (printBank(active)) return -1;
    hBins.assign(hBinsP, hBinsP + (size_t)std::max<int64_t>(1, nb));
    k_new = last.k_cum; keff_0 = k_new; cycleK.push_back(k_new);
    (active ? nSegActive : nSegInactive) += last.n_segments;
    nHist += last.n_start;
    return 0;
  }
  int downloadBank() {
    int cap = 2 * pop + 1024;
    if (cap > hCap) {
      sb_pinned_free(hr); sb_pinned_free(hdir); sb_pinned_free(hw); sb_pinned_free(hG); sb_pinned_free(hE);
      hr = (double*)sb_pinned_alloc(sizeof(double) * 3 * (size_t)cap); hdir = (double*)sb_pinned_alloc(sizeof(double) * 3 * (size_t)cap);
      hw = (double*)sb_pinned_alloc(sizeof(double) * (size_t)cap); hG = (int32_t*)sb_pinned_alloc(sizeof(int32_t) * (size_t)cap);
      hE = (double*)sb_pinned_alloc(sizeof(double) * (size_t)cap);
      if (!hr || !hdir || !hw || !hG || !hE) return fail("pinned host allocation failed");
      hCap = cap;
    }
    if (isCE ? sb_bank_download_ce(eng, cap, &hN, hr, hdir, hw, hE) : sb_bank_download(eng, cap, &hN, hr, hdir, hw, hG)) return engFail();
    return 0;
  }

  int cycles(int active, int N) {
    double k_new = keff_0;
    for (int i = 0; i < N; ++i) if (cycle(active, k_new)) return -1;
    return 0;
  }
  int run() {
    if (generateInitialState()) return -1;
    if (cycles(0, N_inactive)) return -1;
    return cycles(1, N_active);
  }
};

}  // namespace

extern "C" {

static std::string g_sbhErr;
const char* sbh_last_error(void* p) { return p ? ((eigenPhysicsPackage*)p)->err.c_str() : g_sbhErr.c_str(); }

void* sbh_eigen_create(const char* deckPath, const char* overrides, int device, int rank, int nRanks) {
  auto* p = new eigenPhysicsPackage();
  if (p->init(deckPath, overrides, device, rank, nRanks < 1 ? 1 : nRanks)) { g_sbhErr = p->err; if (p->eng) sb_destroy(p->eng); delete p; return nullptr; }
  return p;
}
void sbh_eigen_destroy(void* pv) {
  auto* p = (eigenPhysicsPackage*)pv; if (!p) return;
  sb_pinned_free(p->hBinsP); sb_pinned_free(p->hr); sb_pinned_free(p->hdir); sb_pinned_free(p->hw); sb_pinned_free(p->hG); sb_pinned_free(p->hE);
  if (p->eng) sb_destroy(p->eng);
  delete p;
}
// The flat model of a multigroup eigenvalue deck - exactly the arrays that cross the C ABI (sb_geom_flat, sb_mg_flat, the sb_clerk
// lists of both phases, sb_options) plus the package scalars a driver needs (pop, pRNG, keff_0) - written to a little-endian file:
// "SBFLAT1\0", then a sequence of blocks { int64 count, count items }. tests/c_driver/run_cycle.c reads it and drives the engine
// through sb_* alone; tests/golden/make_flat_model.py writes the committed fixture with it (no device needed).
int sbh_model_dump(void* pv, const char* path) {
  auto* p = (eigenPhysicsPackage*)pv;
  if (p->isCE) { p->err = "sbh_model_dump: multigroup decks only"; return -1; }
  FILE* f = fopen(path, "wb");
  if (!f) { p->err = std::string("sbh_model_dump: cannot open ") + path; return -1; }
  auto blk = [&](const void* d, int64_t n, size_t sz) { fwrite(&n, 8, 1, f); if (n > 0) fwrite(d, sz, (size_t)n, f); };
  auto i32 = [&](int32_t v) { blk(&v, 1, 4); };
  auto f64 = [&](double v) { blk(&v, 1, 8); };
  fwrite("SBFLAT1\0", 1, 8, f);
  const sb_geom_flat g = p->geom.view();
  i32(g.n_surf); blk(g.surf_type, g.n_surf, 4); blk(g.surf_par, (int64_t)g.n_surf * SB_SURF_NPAR, 8);
  i32(g.n_cell); blk(g.cell_off, g.n_cell + 1, 4); blk(g.cell_surf, g.n_cell > 0 ? g.cell_off[g.n_cell] : 0, 4);
  i32(g.n_uni); blk(g.uni_type, g.n_uni, 4); blk(g.uni_ipar, (int64_t)g.n_uni * SB_UNI_NIPAR, 4); blk(g.uni_dpar, (int64_t)g.n_uni * SB_UNI_NDPAR, 8);
  i32(g.n_aux_d); blk(g.aux_d, g.n_aux_d, 8); i32(g.n_aux_i); blk(g.aux_i, g.n_aux_i, 4);
  i32(g.n_graph); blk(g.graph_idx, g.n_graph, 4); blk(g.graph_id, g.n_graph, 4);
  i32(g.root_idx); i32(g.border_idx); blk(g.bc, 6, 4);
  const sb_mg_flat d = p->data.view();
  const int64_t nm = d.n_mat, ng = d.n_g;
  i32(d.n_mat); i32(d.n_g); blk(d.data, nm * ng * 6, 8); blk(d.P0, nm * ng * ng, 8); blk(d.prod, nm * ng * ng, 8);
  blk(d.P1, d.P1 ? nm * ng * ng : 0, 8); blk(d.chi, nm * ng, 8); blk(d.fissile, nm, 4); blk(d.majorant, ng, 8); f64(d.collision_xs);
  for (int ph = 0; ph < 2; ++ph) {
    const auto& T = p->tallies[ph];
    i32((int32_t)T.clerks.size()); i32(T.normClerk); f64(T.normVal);
    for (const sb_clerk& c : T.clerks) {
      i32(c.n_maps); i32(c.n_resp); blk(c.resp_mt, SB_MAX_RESP, 4); i32(c.handle_virtual); i32(c.kind); i32(c.cycles);
      for (int m = 0; m < c.n_maps; ++m) {
        const sb_map1d& q = c.maps[m];
        i32(q.type); i32(q.axis); i32(q.grid); i32(q.n_bins); f64(q.first); f64(q.step); i32(q.default_bin);
        blk(q.bounds, q.bounds ? q.n_bins + 1 : 0, 8); blk(q.mat_bin, q.mat_bin ? nm : 0, 4);
      }
    }
  }
  i32(p->opt.tracking); f64(p->opt.ht_cutoff); i32(p->opt.st_cache); i32(p->opt.max_pop);
  i32(p->pop); blk(&p->pRNG, 1, 8); f64(p->keff_0);
  fclose(f);
  return 0;
}
sb_engine* sbh_engine(void* pv) { return ((eigenPhysicsPackage*)pv)->eng; }
int sbh_eigen_info(void* pv, int* pop, int* nInactive, int* nActive, int* nG, int* nMat, int* nGraph, int* uniqueCells) {
  auto* p = (eigenPhysicsPackage*)pv;
  *pop = p->pop; *nInactive = p->N_inactive; *nActive = p->N_active; *nG = p->data.nG; *nMat = p->data.nMat;
  *nGraph = (int)p->geom.graphIdx.size(); *uniqueCells = p->geom.uniqueCells;
  return 0;
}
int sbh_eigen_total_pop(void* pv) { return ((eigenPhysicsPackage*)pv)->totalPop; }
uint64_t sbh_eigen_rng_state(void* pv) { return ((eigenPhysicsPackage*)pv)->pRNG; }
void sbh_eigen_set_rng_state(void* pv, uint64_t s) { ((eigenPhysicsPackage*)pv)->pRNG = s; }
double sbh_eigen_keff0(void* pv) { return ((eigenPhysicsPackage*)pv)->keff_0; }
int sbh_eigen_generate_initial_state(void* pv) { return ((eigenPhysicsPackage*)pv)->generateInitialState(); }
int sbh_eigen_cycle(void* pv, int active, double* k, sb_cycle_result* res) {
  auto* p = (eigenPhysicsPackage*)pv; int rc = p->cycle(active, *k); if (res) *res = p->last; return rc;
}
int sbh_eigen_cycle_host_buffers(void* pv, int active, double* k, sb_cycle_result* res) {
  auto* p = (eigenPhysicsPackage*)pv; int rc = p->cycleHostBuffers(active, *k); if (res) *res = p->last; return rc;
}
int sbh_eigen_cycle_begin(void* pv, int active, double k, double* devSums, int32_t* nSites) { return ((eigenPhysicsPackage*)pv)->cycleBegin(active, k, devSums, nSites); }
int sbh_eigen_cycle_end(void* pv, int active, const double* devSums, sb_cycle_result* res) {
  auto* p = (eigenPhysicsPackage*)pv; int rc = p->cycleEnd(active, devSums); if (res) *res = p->last; return rc;
}
int sbh_eigen_resample_ranked(void* pv, const int32_t* popSizes, int32_t* newLocal, double* k) {
  return ((eigenPhysicsPackage*)pv)->resampleRanked(popSizes, newLocal, *k);
}
int sbh_eigen_cycle_end_resample_ranked(void* pv, int active, const double* hostSums, const int32_t* popSizes, int32_t* newSizes, double* k, sb_cycle_result* res) {
  auto* p = (eigenPhysicsPackage*)pv; int rc = p->cycleEndResampleRanked(active, hostSums, popSizes, newSizes, *k); if (res) *res = p->last; return rc;
}
// mpi_func.f90:133-159 getWorkshare / getOffset
int sbh_workshare(int totPop, int nRanks, int rank, int* share, int* offset) {
  *share = (totPop + rank) / nRanks;
  *offset = totPop / nRanks * rank + std::max(0, totPop % nRanks + rank - nRanks);
  return 0;
}
// loadBalancing (particleDungeon_class.f90:607-698): numbers of sites this rank sends to / receives from its neighbours.
// out = { send to rank+1 (from the end), receive from rank+1 (to the end), send to rank-1 (from the beginning),
//         receive from rank-1 (to the beginning) }
int sbh_balance_plan(int totPop, int nRanks, int rank, const int32_t* popSizes, int32_t* out) {
  long long off1 = 0, off2 = 0;
  for (int i = 0; i < rank; ++i) off1 += popSizes[i];
  off2 = off1 + popSizes[rank];
  int share, t1, t2;
  sbh_workshare(totPop, nRanks, rank, &share, &t1);
  if (rank + 1 == nRanks) t2 = totPop; else sbh_workshare(totPop, nRanks, rank + 1, &share, &t2);
  long long excessEnd = off2 - t2, excessBeg = off1 - t1;
  out[0] = excessEnd > 0 ? (int)excessEnd : 0; out[1] = excessEnd < 0 ? (int)(-excessEnd) : 0;
  out[2] = excessBeg < 0 ? (int)(-excessBeg) : 0; out[3] = excessBeg > 0 ? (int)excessBeg : 0;
  if (out[0] + out[2] > popSizes[rank]) return -1;          // nearest-neighbour balancing cannot fix this distribution
  return 0;
}
// bank <-> page-locked host arrays (the end-to-end mode of a caller that keeps the dungeons in host memory)
int sbh_eigen_download_bank(void* pv) { return ((eigenPhysicsPackage*)pv)->downloadBank(); }
int sbh_eigen_upload_bank(void* pv) {
  auto* p = (eigenPhysicsPackage*)pv;
  if (p->hN == 0 && p->downloadBank()) return -1;
  if (p->isCE ? sb_bank_upload_ce(p->eng, p->hN, p->hr, p->hdir, p->hw, p->hE) : sb_bank_upload(p->eng, p->hN, p->hr, p->hdir, p->hw, p->hG)) return p->engFail();
  return 0;
}
// the page-locked host copy of the bank (after sbh_eigen_download_bank): n, then pointers to r(3,n), dir(3,n), w(n), G(n), E(n)
int sbh_eigen_host_bank(void* pv, int* n, double** r, double** dir, double** w, int32_t** G, double** E) {
  auto* p = (eigenPhysicsPackage*)pv; *n = p->hN; *r = p->hr; *dir = p->hdir; *w = p->hw; *G = p->hG; *E = p->hE; return 0;
}
// continuous-energy decks: what the engine builds from card `nuc` (1-based) of the deck, computed on the host (no device needed):
// sizes first (grid == NULL), then eGrid(N), mainData(rows, N) and the MT numbers in invertInelastic order
int sbh_ce_card_process(void* pv, int nuc, int* gridSize, int* rows, int* nMT, double* grid, double* mainData, int* mtList, double* awr_kT) {
  auto* p = (eigenPhysicsPackage*)pv;
  if (!p->isCE || nuc < 1 || nuc > (int)p->ceData.cards.size()) { p->err = "sbh_ce_card_process: not a continuous-energy deck or invalid nuclide index"; return -1; }
  try {
    sbk::CardOut out; sb_ace_card c = p->ceData.cards[nuc - 1].view();
    sbk::ceProcessCard(c, p->ceData.energyPerFission, out);
    *gridSize = (int)out.grid.size(); *rows = out.rec.rows; *nMT = out.rec.nMT;
    if (grid) {
      std::copy(out.grid.begin(), out.grid.end(), grid); std::copy(out.main.begin(), out.main.end(), mainData);
      for (size_t i = 0; i < out.mt.size(); ++i) mtList[i] = out.mt[i].MT;
      awr_kT[0] = out.rec.awr; awr_kT[1] = out.rec.kT;
    }
  } catch (const std::exception& e) { p->err = e.what(); return -1; }
  return 0;
}
int sbh_ce_info(void* pv, int* nNuc, int* nMat) { auto* p = (eigenPhysicsPackage*)pv; *nNuc = (int)p->ceData.cards.size(); *nMat = p->ceData.nMat; return 0; }
int sbh_eigen_cycle_peer(void* pv, int active, double* k, int32_t* finalSizes, sb_cycle_result* res) {
  auto* p = (eigenPhysicsPackage*)pv; int rc = p->cyclePeer(active, finalSizes, *k); if (res) *res = p->last; return rc;
}
// input dictionaries (DataStructures/dictParser_func.f90, dictionary_class.f90): value at `path` ("key" or "sub/key") of a dictionary given
// as text or as a file, as text: kind 'i' int, 'r' real (%.17g), 'w' word, 'I' / 'R' / 'W' arrays (space separated), 'k' keys of the
// (sub)dictionary at path ("" = top).  Returns the length, -1 on error (message in sbh_last_error(NULL)).
int sbh_dict_get(const char* text, int isPath, const char* path, char kind, char* out, int cap) {
  try {
    sb::Dict top = isPath ? sb::Dict::fromFile(text) : sb::Dict::fromString(text);
    const sb::Dict* d = &top;
    std::string p = path ? path : "", key;
    for (;;) {
      size_t k = p.find('/');
      if (k == std::string::npos) { key = p; break; }
      d = &d->getDict(p.substr(0, k)); p = p.substr(k + 1);
    }
    std::ostringstream os; os.precision(17);
    if (kind == 'i') os << d->getInt(key);
    else if (kind == 'r') os << d->getReal(key);
    else if (kind == 'w') os << d->getWord(key);
    else if (kind == 'I') { bool f = true; for (int v : d->getIntArray(key)) { os << (f ? "" : " ") << v; f = false; } }
    else if (kind == 'R') { bool f = true; for (double v : d->getRealArray(key)) { os << (f ? "" : " ") << v; f = false; } }
    else if (kind == 'W') { bool f = true; for (auto& v : d->getWordArray(key)) { os << (f ? "" : " ") << v; f = false; } }
    else if (kind == 'k') { const sb::Dict& q = key.empty() ? *d : d->getDict(key); bool f = true; for (auto& v : q.keys("all")) { os << (f ? "" : " ") << v; f = false; } }
    else throw sb::FatalError("sbh_dict_get", "unknown kind");
    std::string r = os.str();
    if ((int)r.size() + 1 > cap) throw sb::FatalError("sbh_dict_get", "buffer too small");
    memcpy(out, r.c_str(), r.size() + 1);
    return (int)r.size();
  } catch (const std::exception& e) { g_sbhErr = e.what(); return -1; }
}
// several ranks: the source dump of this rank after the caller has balanced the banks
int sbh_eigen_print_source(void* pv, int active) { return ((eigenPhysicsPackage*)pv)->printBank(active); }
int sbh_eigen_print_source_mode(void* pv) { return ((eigenPhysicsPackage*)pv)->printSource; }
int sbh_fixed_cycle(void* pv, sb_cycle_result* res) { auto* p = (eigenPhysicsPackage*)pv; int rc = p->fixedCycle(); if (res) *res = p->last; return rc; }
int sbh_eigen_is_fixed(void* pv) { return ((eigenPhysicsPackage*)pv)->isFixed ? 1 : 0; }
int sbh_eigen_is_ce(void* pv) { return ((eigenPhysicsPackage*)pv)->isCE ? 1 : 0; }
int sbh_eigen_cycles(void* pv, int active, int N) { return ((eigenPhysicsPackage*)pv)->cycles(active, N); }
int sbh_eigen_run(void* pv) { return ((eigenPhysicsPackage*)pv)->run(); }
int sbh_eigen_stats(void* pv, long long* segInactive, long long* segActive, long long* hist, double* tTransport) {
  auto* p = (eigenPhysicsPackage*)pv; *segInactive = p->nSegInactive; *segActive = p->nSegActive; *hist = p->nHist; *tTransport = p->timeTransport; return 0;
}
// bytes moved per host-buffer cycle: bank up (pop sites) + bank down (pop sites) + bins
int sbh_eigen_host_bytes(void* pv, int active, long long* h2d, long long* d2h) {
  auto* p = (eigenPhysicsPackage*)pv;
  long long site = 3 * 8 + 3 * 8 + 8 + (p->isCE ? 8 : 4);
  *h2d = site * p->pop; *d2h = site * p->pop + 8 * (long long)sb_tally_size(p->eng, active) + (long long)sizeof(sb_cycle_result);
  return 0;
}

// ---- flat model access for the tests (what the engine was given) ---------------------------------
int sbh_model_graph(void* pv, int* idx, int* id) {
  auto* p = (eigenPhysicsPackage*)pv;
  for (size_t i = 0; i < p->geom.graphIdx.size(); ++i) { idx[i] = p->geom.graphIdx[i]; id[i] = p->geom.graphId[i]; }
  return 0;
}
int sbh_model_xs(void* pv, double* data /*[nMat][nG][6]*/, double* majorant) {
  auto* p = (eigenPhysicsPackage*)pv;
  std::memcpy(data, p->data.data.data(), sizeof(double) * p->data.data.size());
  std::memcpy(majorant, p->data.majorant.data(), sizeof(double) * p->data.majorant.size());
  return 0;
}

// geometry-only handle for geometry parity tests: dictionary text (deck or geometry-level) -> engine
void* sbh_geom_create(const char* text, int isPath, int device) {
  try {
    sb::Dict d = isPath ? sb::Dict::fromFile(text) : sb::Dict::fromString(text);
    sb::MatMap mats = sb::materialMenu(d.getDict("nuclearData"));
    const sb::Dict& gd = d.isPresent("geometry") ? d.getDict("geometry") : d;
    auto* p = new eigenPhysicsPackage();
    p->geom = sb::buildGeometry(gd, mats);
    if (device >= 0) {
      if (sb_create(&p->eng, device)) { g_sbhErr = sb_last_error(nullptr); delete p; return nullptr; }
      sb_geom_flat gv = p->geom.view();
      if (sb_load_geometry(p->eng, &gv)) { g_sbhErr = sb_last_error(p->eng); sb_destroy(p->eng); delete p; return nullptr; }
    }
    return p;
  } catch (const std::exception& e) { g_sbhErr = e.what(); return nullptr; }
}
int sbh_geom_info(void* pv, int* nSurf, int* nCell, int* nUni, int* nGraph, int* uniqueCells, int* rootIdx, int* borderIdx, int* nesting) {
  auto* p = (eigenPhysicsPackage*)pv;
  *nSurf = (int)p->geom.surfType.size(); *nCell = (int)p->geom.cellOff.size() - 1; *nUni = (int)p->geom.uniType.size();
  *nGraph = (int)p->geom.graphIdx.size(); *uniqueCells = p->geom.uniqueCells; *rootIdx = p->geom.rootIdx; *borderIdx = p->geom.borderIdx; *nesting = p->geom.nesting;
  return 0;
}
int sbh_geom_uni_fill(void* pv, int uniIdx, int* out, int cap) {
  auto* p = (eigenPhysicsPackage*)pv;
  auto& f = p->geom.fills.at(uniIdx - 1);
  for (size_t i = 0; i < f.size() && (int)i < cap; ++i) out[i] = f[i];
  return (int)f.size();
}
int sbh_geom_active_mats(void* pv, int* out, int cap) {
  auto a = ((eigenPhysicsPackage*)pv)->geom.activeMats();
  for (size_t i = 0; i < a.size() && (int)i < cap; ++i) out[i] = a[i];
  return (int)a.size();
}

}  // extern "C"
