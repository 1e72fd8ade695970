// ORACLE (test infrastructure, not product code).
// C entry points over the CPU restatement, loaded with ctypes by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs only.
// Also builds as the `scone_oracle` executable (see main at the bottom).
//
// Parity status: component level pinned against the reference's own unit/integration
// test vectors (tests/test_oracle_*.py cite them).  End-to-end k-eff is "parity unpinned"
// by the reference itself: no test in /root/reference runs a physics package.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <sstream>
#include <string>

#include "physics.hpp"
#include "cephysics.hpp"

using namespace orc;

static thread_local std::string g_err;

#define ORC_TRY try {
#define ORC_CATCH(ret)                                   \
  }                                                      \
  catch (const std::exception& e) { g_err = e.what(); return ret; }

struct GeomHandle {
  std::map<std::string, int> mats;
  GeometryStd geom;
};

struct CoordsHandle {
  GeomHandle* g;
  CoordList c;
  DistCache cache;
};

static std::string dirName(const std::string& path) {
  size_t p = path.rfind('/');
  return p == std::string::npos ? std::string(".") : path.substr(0, p);
}

// apply overrides: every top-level entry of `ov` replaces the entry of the same key
static void applyOverrides(Dict& d, const char* overrides) {
  if (!overrides || !*overrides) return;
  Dict ov = Dict::fromString(overrides);
  for (auto& k : ov.keys("all")) {
    bool isDict = false;
    for (auto& dk : ov.keys("dict")) if (dk == k) isDict = true;
    if (isDict) d.setDict(k, ov.getDict(k));
    else d.setScalar(k, ov.getWord(k));
  }
}

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }
void orc_set_math_mode(int m) { mathMode() = m; }
int orc_get_math_mode() { return mathMode(); }

// ---- RNG -------------------------------------------------------------------
uint64_t orc_rng_next(uint64_t state) { RNG r; r.seed = state; return r.getInt(); }
double orc_rng_real(uint64_t state) { return (double)(int64_t)state * (1.0 / 9223372036854775808.0); }
uint64_t orc_rng_skip(uint64_t state, int64_t k) { RNG r; r.seed = state; r.skip(k); return r.seed; }
uint64_t orc_rng_stride(uint64_t state, int32_t n) { RNG r; r.seed = state; r.stride(n); return r.seed; }

double orc_math_log(double x) { return mlog(x); }
void orc_math_sincos(double x, double* s, double* c) { msincos(x, *s, *c); }
void orc_rotate_vector(const double* dir, double mu, double phi, double* out) {
  Vec3 d; for (int k = 0; k < 3; ++k) d[k] = dir[k];
  Vec3 n = rotateVector(d, mu, phi);
  for (int k = 0; k < 3; ++k) out[k] = n[k];
}
int orc_grid_search_lin(double mini, double maxi, int N, double v) { Grid g; g.initEqual(mini, maxi, N, "lin"); return g.search(v); }
// ---- tally maps (for the reference's map unit tests) -------------------------------------------
void* orc_map_new(const char* text, const char* matList) {
  ORC_TRY
  std::map<std::string, int> mats;
  { std::istringstream is(matList ? matList : ""); std::string n; int i; while (is >> n >> i) mats[n] = i; }
  return newTallyMap(Dict::fromString(text), mats).release();
  ORC_CATCH(nullptr)
}
void orc_map_free(void* m) { delete (TallyMap*)m; }
int orc_map_bins(void* m) { return ((TallyMap*)m)->bins(); }
int orc_map_map(void* m, const double* r, double E, int isMG, int G, int matIdx) {
  ParticleState s; for (int k = 0; k < 3; ++k) s.r[k] = r[k];
  s.E = E; s.isMG = isMG != 0; s.G = G; s.matIdx = matIdx;
  return ((TallyMap*)m)->map(s);
}
// grid_class (SharedModules/grid_class.f90): kind 0 lin, 1 log, 2 unstruct (bins given); returns the number of bin boundaries,
// fills idx[] = search(keys[]) and bounds[] = the boundaries
int orc_grid(int kind, double mini, double maxi, int N, const double* binsIn, int nBinsIn, int nKeys, const double* keys, int* idx, double* bounds, int cap) {
  ORC_TRY
  Grid g;
  if (kind == 2) g.initUnstruct(std::vector<double>(binsIn, binsIn + nBinsIn));
  else g.initEqual(mini, maxi, N, kind == 0 ? "lin" : "log");
  for (int i = 0; i < nKeys; ++i) idx[i] = g.search(keys[i]);
  for (size_t i = 0; i < g.bins.size() && (int)i < cap; ++i) bounds[i] = g.bins[i];
  return (int)g.bins.size();
  ORC_CATCH(-1)
}
int orc_binary_search(const double* a, int n, double v) { std::vector<double> x(a, a + n); return Grid::binarySearch(x, v); }

// ---- geometry ----------------------------------------------------------------
// `text` is either a whole deck (with `geometry` and `nuclearData` sub-dictionaries) or a
// geometry-level dictionary that carries its own nuclearData{materials{}} (the layout of
// IntegrationTestFiles/Geometry/test_lat).
void* orc_geom_load(const char* text_or_path, int isPath) {
  ORC_TRY
  Dict d = isPath ? Dict::fromFile(text_or_path) : Dict::fromString(text_or_path);
  auto* h = new GeomHandle();
  h->mats = MgDatabase::materialMenu(d.getDict("nuclearData"));
  const Dict& gd = d.isPresent("geometry") ? d.getDict("geometry") : d;
  h->geom.init(gd, h->mats);
  return h;
  ORC_CATCH(nullptr)
}
void orc_geom_free(void* h) { delete (GeomHandle*)h; }

int orc_geom_info(void* hv, int* nSurf, int* nCell, int* nUni, int* nGraph, int* uniqueCells, int* rootIdx, int* borderIdx, int* nesting) {
  auto* h = (GeomHandle*)hv;
  *nSurf = (int)h->geom.geom.surfs.surfs.size(); *nCell = (int)h->geom.geom.cells.cells.size();
  *nUni = (int)h->geom.geom.unis.size(); *nGraph = (int)h->geom.geom.graph.size();
  *uniqueCells = h->geom.geom.uniqueCells; *rootIdx = h->geom.geom.rootIdx; *borderIdx = h->geom.geom.borderIdx;
  *nesting = h->geom.geom.nesting;
  return 0;
}
int orc_geom_graph(void* hv, int* idx, int* id) {
  auto* h = (GeomHandle*)hv;
  for (size_t i = 0; i < h->geom.geom.graph.size(); ++i) { idx[i] = h->geom.geom.graph[i].idx; id[i] = h->geom.geom.graph[i].id; }
  return 0;
}
int orc_geom_active_mats(void* hv, int* out, int cap) {
  auto* h = (GeomHandle*)hv;
  auto a = h->geom.activeMats();
  for (size_t i = 0; i < a.size() && (int)i < cap; ++i) out[i] = a[i];
  return (int)a.size();
}
int orc_geom_uni_fill(void* hv, int uniIdx, int* out, int cap) {
  auto* h = (GeomHandle*)hv;
  auto& f = h->geom.geom.fills.at(uniIdx - 1);
  for (size_t i = 0; i < f.size() && (int)i < cap; ++i) out[i] = f[i];
  return (int)f.size();
}
int orc_geom_bounds(void* hv, double* b) { ((GeomHandle*)hv)->geom.bounds(b); return 0; }

int orc_geom_what_is_at(void* hv, const double* r, const double* u, int* mat, int* uid) {
  ORC_TRY
  auto* h = (GeomHandle*)hv;
  Vec3 rr, uu; for (int k = 0; k < 3; ++k) { rr[k] = r[k]; if (u) uu[k] = u[k]; }
  h->geom.whatIsAt(*mat, *uid, rr, u ? &uu : nullptr);
  return 0;
  ORC_CATCH(-1)
}
// vectorised whatIsAt for parity sweeps
int orc_geom_what_is_at_n(void* hv, long n, const double* r, const double* u, int* mat, int* uid) {
  ORC_TRY
  auto* h = (GeomHandle*)hv;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    Vec3 rr, uu; for (int k = 0; k < 3; ++k) { rr[k] = r[3 * i + k]; uu[k] = u[3 * i + k]; }
    h->geom.whatIsAt(mat[i], uid[i], rr, &uu);
  }
  return 0;
  ORC_CATCH(-1)
}
// vectorised teleport: place at r, move by dist along u with BC transformation
int orc_geom_teleport_n(void* hv, long n, double* r, double* u, const double* dist, int* mat, int* uid) {
  ORC_TRY
  auto* h = (GeomHandle*)hv;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    Vec3 rr, uu; for (int k = 0; k < 3; ++k) { rr[k] = r[3 * i + k]; uu[k] = u[3 * i + k]; }
    CoordList c; c.init(rr, uu);
    h->geom.placeCoord(c);
    h->geom.teleport(c, dist[i]);
    for (int k = 0; k < 3; ++k) { r[3 * i + k] = c.lvl[0].r[k]; u[3 * i + k] = c.lvl[0].dir[k]; }
    mat[i] = c.matIdx; uid[i] = c.uniqueID;
  }
  return 0;
  ORC_CATCH(-1)
}

void* orc_coords_new(void* hv) { auto* c = new CoordsHandle(); c->g = (GeomHandle*)hv; return c; }
void orc_coords_free(void* cv) { delete (CoordsHandle*)cv; }
int orc_coords_init(void* cv, const double* r, const double* u) {
  auto* c = (CoordsHandle*)cv;
  Vec3 rr, uu; for (int k = 0; k < 3; ++k) { rr[k] = r[k]; uu[k] = u[k]; }
  c->c.init(rr, uu); c->cache = DistCache();
  return 0;
}
int orc_coords_place(void* cv) { ORC_TRY auto* c = (CoordsHandle*)cv; c->g->geom.placeCoord(c->c); return 0; ORC_CATCH(-1) }
int orc_coords_move(void* cv, double* maxDist, int* event, int useCache) {
  ORC_TRY auto* c = (CoordsHandle*)cv; c->g->geom.move(c->c, *maxDist, *event, useCache ? &c->cache : nullptr); return 0; ORC_CATCH(-1)
}
int orc_coords_move_global(void* cv, double* maxDist, int* event) {
  ORC_TRY auto* c = (CoordsHandle*)cv; c->g->geom.moveGlobal(c->c, *maxDist, *event); return 0; ORC_CATCH(-1)
}
int orc_coords_teleport(void* cv, double dist) { ORC_TRY auto* c = (CoordsHandle*)cv; c->g->geom.teleport(c->c, dist); return 0; ORC_CATCH(-1) }
int orc_coords_rotate(void* cv, double mu, double phi) { auto* c = (CoordsHandle*)cv; c->c.rotate(mu, phi); return 0; }
int orc_coords_closest(void* cv, double* dist, int* surfIdx, int* lvl) {
  ORC_TRY auto* c = (CoordsHandle*)cv; c->g->geom.closestDist(*dist, *surfIdx, *lvl, c->c); return 0; ORC_CATCH(-1)
}
int orc_coords_level_distance(void* cv, int lvl, double* d, int* surfIdx) {
  ORC_TRY auto* c = (CoordsHandle*)cv;
  c->g->geom.geom.unis[c->c.lvl[lvl - 1].uniIdx - 1]->distance(*d, *surfIdx, c->c.lvl[lvl - 1]);
  return 0; ORC_CATCH(-1)
}
int orc_coords_get(void* cv, int* nesting, int* mat, int* uid, double* r /*12x3*/, double* dir /*12x3*/,
                   int* uniIdx, int* uniRootID, int* localID, int* cellIdx) {
  auto* c = (CoordsHandle*)cv;
  *nesting = c->c.nesting; *mat = c->c.matIdx; *uid = c->c.uniqueID;
  for (int l = 0; l < MAX_NEST; ++l) {
    for (int k = 0; k < 3; ++k) { r[3 * l + k] = c->c.lvl[l].r[k]; dir[3 * l + k] = c->c.lvl[l].dir[k]; }
    uniIdx[l] = c->c.lvl[l].uniIdx; uniRootID[l] = c->c.lvl[l].uniRootID; localID[l] = c->c.lvl[l].localID; cellIdx[l] = c->c.lvl[l].cellIdx;
  }
  return 0;
}

// ---- single universes (for the reference's universe unit tests) -------------------
struct UniHandle {
  SurfaceShelf surfs; CellShelf cells; std::map<std::string, int> mats;
  std::unique_ptr<Universe> uni; std::vector<int> fill;
};
// uniText: universe dictionary; envText: optional "surfaces { } cells { }"; matList: "name idx name idx ..."
void* orc_uni_new(const char* uniText, const char* envText, const char* matList, int uniIdx) {
  ORC_TRY
  auto* h = new UniHandle();
  { std::istringstream is(matList ? matList : ""); std::string n; int i; while (is >> n >> i) h->mats[n] = i; }
  if (envText && *envText) {
    Dict env = Dict::fromString(envText);
    if (env.isPresent("surfaces")) h->surfs.init(env.getDict("surfaces"));
    if (env.isPresent("cells")) h->cells.init(env.getDict("cells"), h->surfs, h->mats);
  }
  h->uni = newUniverse(h->fill, Dict::fromString(uniText), h->cells, h->surfs, h->mats);
  h->uni->uniIdx = uniIdx;
  return h;
  ORC_CATCH(nullptr)
}
void orc_uni_free(void* h) { delete (UniHandle*)h; }
int orc_uni_fill(void* hv, int* out, int cap) {
  auto* h = (UniHandle*)hv;
  for (size_t i = 0; i < h->fill.size() && (int)i < cap; ++i) out[i] = h->fill[i];
  return (int)h->fill.size();
}
int orc_uni_enter(void* hv, const double* r, const double* u, double* rOut, double* uOut, int* uniIdx, int* localID, int* cellIdx) {
  ORC_TRY
  auto* h = (UniHandle*)hv;
  Vec3 rr, uu; for (int k = 0; k < 3; ++k) { rr[k] = r[k]; uu[k] = u[k]; }
  Coord c; h->uni->enter(c, rr, uu);
  for (int k = 0; k < 3; ++k) { rOut[k] = c.r[k]; uOut[k] = c.dir[k]; }
  *uniIdx = c.uniIdx; *localID = c.localID; *cellIdx = c.cellIdx;
  return 0;
  ORC_CATCH(-1)
}
int orc_uni_distance(void* hv, int localID, const double* r, const double* u, double* d, int* surfIdx) {
  ORC_TRY
  auto* h = (UniHandle*)hv;
  Coord c; for (int k = 0; k < 3; ++k) { c.r[k] = r[k]; c.dir[k] = u[k]; }
  c.localID = localID; c.uniIdx = h->uni->uniIdx;
  h->uni->distance(*d, *surfIdx, c);
  return 0;
  ORC_CATCH(-1)
}
int orc_uni_cross(void* hv, int localID, const double* r, const double* u, int surfIdx, int* newLocalID) {
  ORC_TRY
  auto* h = (UniHandle*)hv;
  Coord c; for (int k = 0; k < 3; ++k) { c.r[k] = r[k]; c.dir[k] = u[k]; }
  c.localID = localID; c.uniIdx = h->uni->uniIdx;
  h->uni->cross(c, surfIdx);
  *newLocalID = c.localID;
  return 0;
  ORC_CATCH(-1)
}
int orc_uni_offset(void* hv, int localID, double* off) {
  auto* h = (UniHandle*)hv;
  Coord c; c.localID = localID;
  Vec3 o = h->uni->cellOffset(c);
  for (int k = 0; k < 3; ++k) off[k] = o[k];
  return 0;
}
// single surface from a dictionary: evaluate / distance / going / halfspace / BCs
void* orc_surf_new(const char* text) { ORC_TRY return newSurface(Dict::fromString(text)).release(); ORC_CATCH(nullptr) }
void orc_surf_free(void* s) { delete (Surface*)s; }
int orc_surf_set_bc(void* sv, const int* bc, int n) { ORC_TRY ((Surface*)sv)->setBC(std::vector<int>(bc, bc + n)); return 0; ORC_CATCH(-1) }
int orc_surf_query(void* sv, const double* r, const double* u, double* evaluate, double* distance, int* going, int* halfspace) {
  auto* s = (Surface*)sv;
  Vec3 rr, uu; for (int k = 0; k < 3; ++k) { rr[k] = r[k]; uu[k] = u[k]; }
  *evaluate = s->evaluate(rr); *distance = s->distance(rr, uu); *going = s->going(rr, uu) ? 1 : 0; *halfspace = s->halfspace(rr, uu) ? 1 : 0;
  return 0;
}
int orc_surf_bc(void* sv, int transform, double* r, double* u) {
  auto* s = (Surface*)sv;
  Vec3 rr, uu; for (int k = 0; k < 3; ++k) { rr[k] = r[k]; uu[k] = u[k]; }
  if (transform) s->transformBC(rr, uu); else s->explicitBC(rr, uu);
  for (int k = 0; k < 3; ++k) { r[k] = rr[k]; u[k] = uu[k]; }
  return 0;
}

// ---- MG data -----------------------------------------------------------------
// builds the database of a deck (materials + PN) without geometry; all materials active
void* orc_mg_load(const char* deckPath, const char* handleName) {
  ORC_TRY
  Dict d = Dict::fromFile(deckPath);
  const Dict& nd = d.isPresent("nuclearData") ? d.getDict("nuclearData") : d;
  auto* db = new MgDatabase();
  db->init(nd, handleName, dirName(deckPath));
  std::vector<int> act; for (size_t i = 1; i <= db->mats.size(); ++i) act.push_back((int)i);
  db->activate(act);
  return db;
  ORC_CATCH(nullptr)
}
void orc_mg_free(void* db) { delete (MgDatabase*)db; }
int orc_mg_info(void* dbv, int* nMat, int* nG) { auto* db = (MgDatabase*)dbv; *nMat = (int)db->mats.size(); *nG = db->nG; return 0; }
int orc_mg_mat_idx(void* dbv, const char* name) { auto* db = (MgDatabase*)dbv; auto it = db->nameMap.find(name); return it == db->nameMap.end() ? -1 : it->second; }
// out[8]: total, elastic, inelastic, capture, fission, nuFission, kappa, fissile
int orc_mg_macro(void* dbv, int matIdx, int G, double* out) {
  ORC_TRY
  auto* db = (MgDatabase*)dbv;
  MacroXSs x; db->mats.at(matIdx - 1).getMacroXSs(x, G);
  out[0] = x.total; out[1] = x.elasticScatter; out[2] = x.inelasticScatter; out[3] = x.capture; out[4] = x.fission;
  out[5] = x.nuFission; out[6] = x.kappaXS; out[7] = db->mats.at(matIdx - 1).fissile ? 1.0 : 0.0;
  return 0;
  ORC_CATCH(-1)
}
double orc_mg_majorant(void* dbv, int G) { return ((MgDatabase*)dbv)->getMajorantXS(G); }
double orc_mg_total(void* dbv, int matIdx, int G) { ORC_TRY return ((MgDatabase*)dbv)->getTotalMatXS(G, matIdx); ORC_CATCH(-1.0) }
int orc_mg_matrices(void* dbv, int matIdx, double* P0, double* prod, double* P1, double* chi, double* nu) {
  auto* db = (MgDatabase*)dbv; auto& m = db->mats.at(matIdx - 1);
  int n = m.nG * m.nG;
  for (int i = 0; i < n; ++i) { P0[i] = m.P0[i]; prod[i] = m.prod[i]; P1[i] = m.isP1 ? m.P1[i] : 0.0; }
  for (int g = 0; g < m.nG; ++g) { chi[g] = m.fissile ? m.chi[g] : 0.0; nu[g] = m.fissile ? m.nu[g] : 0.0; }
  return m.isP1 ? 1 : 0;
}
// reaction sampling with an explicit RNG state (returns the advanced state)
uint64_t orc_mg_sample_scatter(void* dbv, int matIdx, int G_in, uint64_t state, double* mu, double* phi, int* G_out) {
  auto* db = (MgDatabase*)dbv; RNG r; r.seed = state;
  db->mats.at(matIdx - 1).scatterSampleOut(*mu, *phi, *G_out, G_in, r);
  return r.seed;
}
uint64_t orc_mg_sample_fission(void* dbv, int matIdx, uint64_t state, double* mu, double* phi, int* G_out) {
  auto* db = (MgDatabase*)dbv; RNG r; r.seed = state;
  db->mats.at(matIdx - 1).fissionSampleOut(*mu, *phi, *G_out, r);
  return r.seed;
}

// heapQueue (DataStructures/heapQueue_class.f90): push the sequence (conditional = only values below the current maximum once the
// queue is non-empty, as normSize_Repr uses it); returns the maximum, *size = entries held
double orc_heap_queue(int maxSize, int n, const double* seq, int conditional, int* size) {
  ORC_TRY
  HeapQueue hq; hq.init(maxSize);
  for (int i = 0; i < n; ++i) {
    if (conditional && hq.size > 0 && !(seq[i] < hq.maxValue())) continue;
    hq.pushReplace(seq[i]);
  }
  *size = hq.size;
  return hq.maxValue();
  ORC_CATCH(std::nan(""))
}

// ---- dungeon -------------------------------------------------------------------
// normSize_Repr on a bank described by its broodIDs; `tag` travels with each site so the
// caller can see which sites survive and in which order. Returns new size (or -1).
int orc_dungeon_norm_size(int n, const int* brood, int* tag, int cap, int totPop, uint64_t rngState) {
  ORC_TRY
  Dungeon d; d.init(std::max(cap, 2 * totPop));
  for (int i = 0; i < n; ++i) { ParticleState p; p.broodID = brood[i]; p.collisionN = tag[i]; p.wgt = 1.0; d.prisoners[i] = p; }
  d.pop = n;
  RNG r; r.seed = rngState;
  d.normSize_Repr(totPop, r);
  for (int i = 0; i < d.pop && i < cap; ++i) tag[i] = d.prisoners[i].collisionN;
  return d.pop;
  ORC_CATCH(-1)
}
int orc_dungeon_sort(int n, const int* brood, int* tag) {
  ORC_TRY
  Dungeon d; d.init(n);
  int mx = 0;
  for (int i = 0; i < n; ++i) { ParticleState p; p.broodID = brood[i]; p.collisionN = tag[i]; d.prisoners[i] = p; mx = std::max(mx, brood[i]); }
  d.pop = n;
  d.sortByBroodID(mx);
  for (int i = 0; i < n; ++i) tag[i] = d.prisoners[i].collisionN;
  return 0;
  ORC_CATCH(-1)
}

// ---- eigenvalue driver -----------------------------------------------------------
void* orc_eigen_load(const char* deckPath, const char* overrides) {
  ORC_TRY
  Dict d = Dict::fromFile(deckPath);
  applyOverrides(d, overrides);
  EigenPP* e = (d.getWord("dataType") == "ce") ? (EigenPP*)new CeEigenPP() : new EigenPP();
  try { e->init(d, dirName(deckPath)); } catch (...) { delete e; throw; }
  return e;
  ORC_CATCH(nullptr)
}
void orc_eigen_free(void* e) { delete (EigenPP*)e; }
int orc_eigen_info(void* ev, int* pop, int* nInactive, int* nActive, int* nG, int* nMat, int* tracking) {
  auto* e = (EigenPP*)ev;
  *pop = e->pop; *nInactive = e->N_inactive; *nActive = e->N_active; *nG = e->db.nG; *nMat = (int)e->db.mats.size(); *tracking = e->tracking;
  return 0;
}
uint64_t orc_eigen_rng_state(void* ev) { return ((EigenPP*)ev)->pRNG.seed; }
void orc_eigen_set_rng_state(void* ev, uint64_t s) { ((EigenPP*)ev)->pRNG.seed = s; }
double orc_eigen_keff0(void* ev) { return ((EigenPP*)ev)->keff_0; }
// cumulative k-eff of the active (which = 1) or inactive (0) attachment clerk: mean and standard deviation of the mean
int orc_eigen_keff(void* ev, int which, double* k, double* std_) {
  auto* e = (EigenPP*)ev; return (which ? e->activeAtch : e->inactiveAtch).getKeff(*k, *std_) ? 0 : -1;
}
int orc_eigen_init_source(void* ev) { ORC_TRY ((EigenPP*)ev)->generateInitialState(); return 0; ORC_CATCH(-1) }
// one cycle; k_in is the k used for site generation; returns k_new (NaN on error)
double orc_eigen_cycle(void* ev, int active, double k_in) {
  ORC_TRY return ((EigenPP*)ev)->cycle(active != 0, k_in); ORC_CATCH(std::nan(""))
}
// ---- scoreMemory on its own (pins of Tallies/Tests/scoreMemory_test.f90) ----
void* orc_mem_new(long n, int batch) { auto* m = new ScoreMemory(); m->init(n, batch); return m; }
void orc_mem_free(void* m) { delete (ScoreMemory*)m; }
int orc_mem_score(void* m, double v, long idx) { ORC_TRY ((ScoreMemory*)m)->score(v, idx); return 0; ORC_CATCH(-1) }
int orc_mem_accumulate(void* m, double v, long idx) { ORC_TRY ((ScoreMemory*)m)->accumulate(v, idx); return 0; ORC_CATCH(-1) }
int orc_mem_reduce(void* m) { ((ScoreMemory*)m)->reduceBins(); return 0; }
int orc_mem_close_bin(void* m, double norm, long idx) { ORC_TRY ((ScoreMemory*)m)->closeBin(norm, idx); return 0; ORC_CATCH(-1) }
int orc_mem_close_cycle(void* m, double norm) { ((ScoreMemory*)m)->closeCycle(norm); return 0; }
int orc_mem_last_cycle(void* m) { return ((ScoreMemory*)m)->lastCycle() ? 1 : 0; }
double orc_mem_get_score(void* m, long idx) { return ((ScoreMemory*)m)->getScore(idx); }
int orc_mem_result(void* m, long idx, int samples, double* mean, double* std_) { ((ScoreMemory*)m)->getResult(*mean, *std_, idx, samples); return 0; }
// ---- k-eff clerks on their own (pins of Tallies/TallyClerks/Tests/keff{Implicit,Analog}Clerk_test.f90) ----
// testNeutronDatabase (NuclearData/testNeutronData/testNeutronDatabase_class.f90): the same cross sections everywhere
struct ConstXsView : XsView {
  MacroXSs x;
  int nMat() const override { return 1 << 30; }
  double totalMatXS(const Particle&, int) const override { return x.total; }
  void macroXSs(MacroXSs& o, const Particle&, int) const override { o = x; }
  double trackMatXS(const Particle&, int) const override { return x.total; }
  double majorantXS(const Particle&) const override { return x.total; }
  double collisionXS() const override { return 0.0; }
  bool isFissileMat(int) const override { return x.fission > 0.0; }
};
// keffImplicitClerk_test.f90 test1CycleBatch: two cycles of { collision at weight w_coll, (n,2n) with pre-collision weight w_pre, leak at w_leak }
int orc_keff_implicit_sequence(double total, double capture, double fission, double nuFission, int n, const double* w_coll, const double* w_pre,
                               const double* w_leak, double* k, double* std_) {
  ORC_TRY
  ConstXsView xs; xs.x.total = total; xs.x.capture = capture; xs.x.fission = fission; xs.x.nuFission = nuFission;
  TallyAdmin t; std::map<std::string, int> mats;
  t.init(Dict::fromString("k { type keffImplicitClerk; }"), mats);
  Dungeon pit; pit.init(4);
  for (int c = 0; c < n; ++c) {
    Particle p; p.coords.matIdx = 1;
    p.w = w_coll[c];
    t.reportInColl(p, xs, total, false);
    p.preCollision.wgt = w_pre[c];
    t.reportOutColl(p, 16);                                           // N_2N
    p.w = w_leak[c]; p.fate = LEAK_FATE;
    t.reportHist(p);
    t.reportCycleEnd(pit);
  }
  return t.getKeff(*k, *std_) ? 0 : -1;
  ORC_CATCH(-1)
}
// collisionClerk_test.f90 testScoring / testScoringVirtual and trackClerk_test.f90 testScoring: a clerk of the given dictionary is fed
// n events { material, weight, virtual flag (collision) or path length (track) } over the constant-XS database, then one cycle is closed
// with norm 1; out[] = the means of the clerk's bins.  kind 0: reportInColl, kind 1: reportPath.
int orc_clerk_sequence(const char* clerkText, const char* matList, double xsAll, double trackingXS, int kind, int n, const int* matIdx,
                       const double* w, const double* aux, double* out, int cap) {
  ORC_TRY
  ConstXsView xs; xs.x.total = xsAll; xs.x.capture = xsAll; xs.x.fission = xsAll; xs.x.nuFission = xsAll; xs.x.elasticScatter = xsAll; xs.x.inelasticScatter = xsAll;
  std::map<std::string, int> mats;
  { std::istringstream is(matList ? matList : ""); std::string nm; int i; while (is >> nm >> i) mats[nm] = i; }
  TallyAdmin t;
  t.init(Dict::fromString(std::string("myClerk { ") + clerkText + " }"), mats);
  for (int i = 0; i < n; ++i) {
    Particle p; p.coords.matIdx = matIdx[i]; p.w = w[i]; p.E = 10.0; p.isMG = false;
    p.preCollision = p.state(); p.prePath = p.state();
    if (kind == 0) t.reportInColl(p, xs, trackingXS, aux[i] != 0.0);
    else t.reportPath(p, xs, aux[i]);
  }
  Dungeon pit; pit.init(1);
  t.reportCycleEnd(pit);
  if (t.mem.N > cap) return -2;
  for (long i = 0; i < t.mem.N; ++i) { double m, sd; t.mem.getResult(m, sd, i + 1, 1); out[i] = m; }
  return (int)t.mem.N;
  ORC_CATCH(-1)
}
// tally responses over the constant-XS database (TallyResponses/Tests/macroResponse_test.f90, fluxResponse_test.f90):
// xs = { total, elastic, inelastic, capture, fission, nuFission, kappaXS }
double orc_response_value(const char* respText, const double* xs) {
  ORC_TRY
  ConstXsView v; v.x.total = xs[0]; v.x.elasticScatter = xs[1]; v.x.inelasticScatter = xs[2]; v.x.capture = xs[3]; v.x.fission = xs[4];
  v.x.nuFission = xs[5]; v.x.kappaXS = xs[6];
  Response r; r.init(Dict::fromString(respText));
  Particle p; p.coords.matIdx = 1; p.w = 1.0;
  return r.get(v, p);
  ORC_CATCH(std::nan(""))
}
// shannonEntropyClerk_test.f90 testSimpleUseCase: cycles of end-of-cycle populations { matIdx, x, wgt }; out[c] = entropy of cycle c
int orc_shannon_sequence(const char* clerkText, const char* matList, int nCycles, const int* counts, const int* matIdx, const double* x,
                         const double* w, double* out) {
  ORC_TRY
  std::map<std::string, int> mats;
  { std::istringstream is(matList ? matList : ""); std::string nm; int i; while (is >> nm >> i) mats[nm] = i; }
  TallyAdmin t;
  t.init(Dict::fromString(std::string("testClerk { ") + clerkText + " }"), mats);
  int k = 0;
  for (int c = 0; c < nCycles; ++c) {
    Dungeon pop; pop.init(counts[c] + 1);
    for (int i = 0; i < counts[c]; ++i, ++k) { ParticleState s; s.wgt = w[k]; s.matIdx = matIdx[k]; s.r[0] = x[k]; pop.detain(s); }
    t.reportCycleEnd(pop);
  }
  const Clerk& cl = t.clerks.at(0);
  for (int c = 0; c < std::min(nCycles, cl.maxCycles); ++c) { double m, sd; t.mem.getResult(m, sd, cl.addr + cl.map->bins() + 1 + c, 1); out[c] = m; }
  return (int)t.mem.N;
  ORC_CATCH(-1)
}
// keffAnalogClerk_test.f90 test1CycleBatch: cycles of { start population weight, end population weight, k_eff of the end dungeon }; closeCycle norm 0.8
int orc_keff_analog_sequence(int n, const double* w_start, const double* w_end, const double* k_norm, double norm, double* k, double* std_) {
  ORC_TRY
  TallyAdmin t; std::map<std::string, int> mats;
  t.init(Dict::fromString("k { type keffAnalogClerk; }"), mats);
  (void)norm;
  for (int c = 0; c < n; ++c) {
    Dungeon a, b; a.init(1); b.init(1);
    ParticleState s; s.wgt = w_start[c]; a.detain(s);
    t.reportCycleStart(a);
    s.wgt = w_end[c]; b.detain(s); b.k_eff = k_norm[c];
    t.reportCycleEnd(b);
  }
  return t.getKeff(*k, *std_) ? 0 : -1;
  ORC_CATCH(-1)
}
// fixedSourcePhysicsPackage: one source batch (returns 0 / -1); pop, cycles via orc_eigen_info
int orc_fixed_cycle(void* ev) { ORC_TRY ((EigenPP*)ev)->fixedCycle(); return 0; ORC_CATCH(-1) }
int orc_eigen_is_fixed(void* ev) { return ((EigenPP*)ev)->fixedSource ? 1 : 0; }
int orc_eigen_run(void* ev) { ORC_TRY ((EigenPP*)ev)->run(); return 0; ORC_CATCH(-1) }
int orc_eigen_bank_size(void* ev) { return ((EigenPP*)ev)->thisCycle->pop; }
// current source bank (thisCycle) as SoA
int orc_eigen_bank(void* ev, double* r, double* dir, double* w, int* G, int* brood) {
  auto* e = (EigenPP*)ev;
  for (int i = 0; i < e->thisCycle->pop; ++i) {
    auto& p = e->thisCycle->prisoners[i];
    for (int k = 0; k < 3; ++k) { r[3 * i + k] = p.r[k]; dir[3 * i + k] = p.dir[k]; }
    w[i] = p.wgt; G[i] = p.G; brood[i] = p.broodID;
  }
  return e->thisCycle->pop;
}
// energies of the bank (continuous-energy runs)
int orc_eigen_bank_E(void* ev, double* E) {
  auto* e = (EigenPP*)ev;
  for (int i = 0; i < e->thisCycle->pop; ++i) E[i] = e->thisCycle->prisoners[i].E;
  return e->thisCycle->pop;
}
int orc_eigen_set_bank(void* ev, int n, const double* r, const double* dir, const double* w, const int* G) {
  ORC_TRY
  auto* e = (EigenPP*)ev;
  if (e->dungeonA.prisoners.empty()) { e->dungeonA.init(2 * e->pop); e->dungeonB.init(2 * e->pop); e->thisCycle = &e->dungeonA; e->nextCycle = &e->dungeonB; }
  if (n > (int)e->thisCycle->prisoners.size()) throw FatalError("orc_eigen_set_bank", "bank too large");
  for (int i = 0; i < n; ++i) {
    ParticleState p;
    for (int k = 0; k < 3; ++k) { p.r[k] = r[3 * i + k]; p.dir[k] = dir[3 * i + k]; }
    p.wgt = w[i]; p.G = G[i]; p.isMG = true;
    e->thisCycle->prisoners[i] = p;
  }
  e->thisCycle->pop = n;
  return 0;
  ORC_CATCH(-1)
}
long orc_eigen_tally_size(void* ev, int which /*0 inactive,1 active,2 inactiveAtch,3 activeAtch*/) {
  auto* e = (EigenPP*)ev;
  TallyAdmin* t = which == 0 ? &e->inactiveTally : which == 1 ? &e->activeTally : which == 2 ? &e->inactiveAtch : &e->activeAtch;
  return t->mem.N;
}
int orc_eigen_tally(void* ev, int which, double* csum, double* csum2, int* batchN) {
  auto* e = (EigenPP*)ev;
  TallyAdmin* t = which == 0 ? &e->inactiveTally : which == 1 ? &e->activeTally : which == 2 ? &e->inactiveAtch : &e->activeAtch;
  for (long i = 0; i < t->mem.N; ++i) { csum[i] = t->mem.csum[i]; csum2[i] = t->mem.csum2[i]; }
  *batchN = t->mem.batchN;
  return 0;
}
int orc_eigen_stats(void* ev, long* seg, long* coll, long* hist) {
  auto* e = (EigenPP*)ev; *seg = e->nSegments; *coll = e->nCollisions; *hist = e->nHistories; return 0;
}

}  // extern "C"


// ---------------------------------------------------------------------------------------------------------
// continuous-energy data (oracle/cedata.hpp)
// ---------------------------------------------------------------------------------------------------------
#include "cedata.hpp"
#define CE_TRY try {
#define CE_CATCH(ret) } catch (const std::exception& ex) { g_err = ex.what(); return ret; }
extern "C" {
void* orc_ce_nuclide_from_ace(const char* path, int lineNum) {
  CE_TRY
  orc_ce::AceCard ace; ace.readFromFile(path, lineNum);
  auto* n = new orc_ce::Nuclide(); n->init(ace, true);
  return n;
  CE_CATCH(nullptr)
}
void* orc_ce_nuclide_from_arrays(int n, int rows, const double* grid, const double* data) {
  auto* nuc = new orc_ce::Nuclide(); nuc->fromArrays(n, rows, grid, data); return nuc;
}
void orc_ce_nuclide_free(void* h) { delete (orc_ce::Nuclide*)h; }
int orc_ce_nuclide_info(void* h, int* n, int* rows, double* mass, double* kT) {
  auto* x = (orc_ce::Nuclide*)h; *n = x->N(); *rows = x->rows; *mass = x->mass; *kT = x->kT; return 0;
}
int orc_ce_nuclide_data(void* h, double* grid, double* data) {
  auto* x = (orc_ce::Nuclide*)h;
  std::copy(x->eGrid.begin(), x->eGrid.end(), grid); std::copy(x->main.begin(), x->main.end(), data); return 0;
}
int orc_ce_nuclide_search(void* h, double E, int* idx, double* f) { CE_TRY ((orc_ce::Nuclide*)h)->search(*idx, *f, E); return 0; CE_CATCH(-1) }
int orc_ce_nuclide_micro(void* h, double E, double* out8) {
  CE_TRY auto* x = (orc_ce::Nuclide*)h; int idx; double f; x->search(idx, f, E); x->microXSs(out8, idx, f); return 0; CE_CATCH(-1)
}
double orc_ce_nuclide_total(void* h, double E) {
  CE_TRY auto* x = (orc_ce::Nuclide*)h; int idx; double f; x->search(idx, f, E); return x->totalXS(idx, f); CE_CATCH(std::nan(""))
}
int orc_ce_nuclide_nubar(void* h, double E, double* total, double* prompt, double* delayed) {
  CE_TRY auto* x = (orc_ce::Nuclide*)h; *total = x->release(E); *prompt = x->releasePrompt(E); *delayed = x->releaseDelayed(E); return 0; CE_CATCH(-1)
}
// ---- pins of the CE reaction restatement (tests/test_oracle_cereact.py) -------------------------------------------
double orc_tabpdf_sample(int n, const double* x, const double* pdf, const double* cdf, int flag, double r) {
  CE_TRY orc_ce::TabularPdf t; t.initCdf(std::vector<double>(x, x + n), std::vector<double>(pdf, pdf + n), std::vector<double>(cdf, cdf + n), flag); return t.sample(r); CE_CATCH(std::nan(""))
}
double orc_endftable_at(int n, const double* x, const double* y, int nr, const int* bounds, const int* inter, double v) {
  CE_TRY orc_ce::EndfTable t; t.x.assign(x, x + n); t.y.assign(y, y + n);
  if (nr > 0) { t.bounds.assign(bounds, bounds + nr); t.inter.assign(inter, inter + nr); }
  return t.at(v); CE_CATCH(std::nan(""))
}
void* orc_ce_nuclide_from_acebin(const char* path) {
  CE_TRY orc_ce::AceCard ace; orc::readAceBin(ace, path); auto* n = new orc_ce::Nuclide(); n->init(ace, true); return n; CE_CATCH(nullptr)
}
// aceCard header and fission-data flags (DataDecks/Tests/aceCard_iTest.f90): out = { AW, TZ }, flags = { precursorGroups, isFissile,
// hasNuPrompt, hasNuDelayed, hasNuTotal }, zaid[16]
int orc_ace_card_info(const char* path, double* out, int* flags, char* zaid) {
  CE_TRY orc_ce::AceCard ace; orc::readAceBin(ace, path);
  out[0] = ace.AW; out[1] = ace.TZ;
  flags[0] = ace.NXS[7]; flags[1] = ace.isFiss ? 1 : 0; flags[2] = ace.promptNUp != 0; flags[3] = ace.delayNUp != 0; flags[4] = ace.totalNUp != 0;
  strncpy(zaid, ace.ZAID.c_str(), 15); zaid[15] = 0;
  return 0; CE_CATCH(-1)
}
// elasticNeutronScatter of the nuclide: 1 if the angular law is isotropic (LOCB == 0), else 0
int orc_ce_nuclide_elastic_isotropic(void* h) { return ((orc_ce::Nuclide*)h)->elastic.angle.kind == orc_ce::AngleLaw::ISOTROPIC ? 1 : 0; }
int orc_ce_nuclide_mt_list(void* h, int* MTs, int* firstIdx) {
  auto* x = (orc_ce::Nuclide*)h;
  for (int i = 0; i < x->nMTinelastic; ++i) { MTs[i] = x->mtData[i].MT; firstIdx[i] = x->mtData[i].firstIdx; }
  return x->nMTinelastic;
}
// sampleOut of one reaction of the nuclide from a given RNG state: kind 0 elastic, 1 inelastic MT record `which` (0-based), 2 fission.
// out = { mu, phi, E_out, lambda }, returns the number of random numbers drawn (< 0 on error)
long orc_ce_nuclide_sample(void* h, int kind, int which, double E_in, uint64_t state, double* out) {
  CE_TRY auto* x = (orc_ce::Nuclide*)h; orc::RNG r; r.init((int64_t)state);
  double mu = 0, phi = 0, E = 0, lam = orc_ce::HUGE_LAMBDA;
  if (kind == 0) x->elastic.sampleOut(mu, phi, E, E_in, r);
  else if (kind == 1) x->mtData.at(which).kin.sampleOut(mu, phi, E, E_in, r);
  else x->fission.sampleOut(mu, phi, E, E_in, r, lam);
  out[0] = mu; out[1] = phi; out[2] = E; out[3] = lam;
  return (long)r.count; CE_CATCH(-1)
}
int orc_ce_nuclide_invert_inelastic(void* h, double E, uint64_t state) {
  CE_TRY auto* x = (orc_ce::Nuclide*)h; orc::RNG r; r.init((int64_t)state); return x->invertInelastic(E, r); CE_CATCH(-1)
}
double orc_ce_nuclide_mt_release(void* h, int which, double E) { CE_TRY return ((orc_ce::Nuclide*)h)->mtData.at(which).kin.release(E); CE_CATCH(std::nan("")) }
int orc_ce_nuclide_mt_cm(void* h, int which) { return ((orc_ce::Nuclide*)h)->mtData.at(which).kin.cmFrame ? 1 : 0; }
void* orc_ce_db_new() { return new orc_ce::Database(); }
void orc_ce_db_free(void* h) { delete (orc_ce::Database*)h; }
int orc_ce_db_add_nuclide(void* h, void* nuc) { auto* d = (orc_ce::Database*)h; d->nuclides.push_back(*(orc_ce::Nuclide*)nuc); return (int)d->nuclides.size(); }
int orc_ce_db_add_material(void* h, int n, const int* nucIdx, const double* dens) {
  auto* d = (orc_ce::Database*)h; orc_ce::Material m; m.nuclides.assign(nucIdx, nucIdx + n); m.dens.assign(dens, dens + n); d->materials.push_back(m); return (int)d->materials.size();
}
int orc_ce_db_finalise(void* h) { CE_TRY auto* d = (orc_ce::Database*)h; d->finalise(); d->initMajorant(); return (int)d->eGridUnion.size(); CE_CATCH(-1) }
int orc_ce_db_union(void* h, double* grid, double* maj) {
  auto* d = (orc_ce::Database*)h; std::copy(d->eGridUnion.begin(), d->eGridUnion.end(), grid); std::copy(d->majorant.begin(), d->majorant.end(), maj); return 0;
}
// batch lookups (the CPU side of the parity tests and of bench.py's cpu_baseline for the lookup kernel); OpenMP over particles
int orc_ce_db_total_n(void* h, long n, const double* E, const int* mat, double* out) {
  CE_TRY auto* d = (orc_ce::Database*)h;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) out[i] = d->totalMatXS(E[i], mat[i]);
  return 0; CE_CATCH(-1)
}
int orc_ce_db_macro_n(void* h, long n, const double* E, const int* mat, double* out8) {
  CE_TRY auto* d = (orc_ce::Database*)h;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) d->macroXSs(out8 + 8 * i, E[i], mat[i]);
  return 0; CE_CATCH(-1)
}
int orc_ce_db_majorant_n(void* h, long n, const double* E, double* out) {
  CE_TRY auto* d = (orc_ce::Database*)h;
  for (long i = 0; i < n; ++i) out[i] = d->majorantXS(E[i]);
  return 0; CE_CATCH(-1)
}
int orc_ce_db_index_n(void* h, int nucIdx, long n, const double* E, int* idx) {
  auto* d = (orc_ce::Database*)h; const auto& g = d->nuclides.at(nucIdx - 1).eGrid;
  for (long i = 0; i < n; ++i) idx[i] = orc_ce::binarySearch(g, E[i]);
  return 0;
}
}  // extern "C"

#ifdef ORC_MAIN
// scone_oracle <deck> [--omp N] [--pop P] [--inactive I] [--active A] [--seed S] [--tracking DT|ST|HT] [--math libm|sb]
int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: scone_oracle <deck> [--omp N] [--pop P] [--inactive I] [--active A] [--seed S] [--tracking DT|ST|HT] [--math libm|sb]\n"); return 2; }
  std::string deck = argv[1], ov;
  for (int i = 2; i + 1 < argc; i += 2) {
    std::string k = argv[i], v = argv[i + 1];
    if (k == "--omp") omp_set_num_threads(std::atoi(v.c_str()));
    else if (k == "--pop") ov += "pop " + v + "; ";
    else if (k == "--inactive") ov += "inactive " + v + "; ";
    else if (k == "--active") ov += "active " + v + "; ";
    else if (k == "--seed") ov += "seed " + v + "; ";
    else if (k == "--tracking") ov += "transportOperator { type transportOperator" + v + "; } ";
    else if (k == "--math") mathMode() = (v == "sb") ? MATH_SB : MATH_LIBM;
  }
  if (ov.find("seed") == std::string::npos) ov += "seed 20261017; ";
  void* e = orc_eigen_load(deck.c_str(), ov.c_str());
  if (!e) { std::fprintf(stderr, "error: %s\n", orc_last_error()); return 1; }
  auto* pp = (EigenPP*)e;
  try {
    pp->generateInitialState();
    auto t0 = std::chrono::steady_clock::now();
    pp->runCycles(false, pp->N_inactive);
    auto t1 = std::chrono::steady_clock::now();
    long seg0 = pp->nSegments;
    pp->runCycles(true, pp->N_active);
    auto t2 = std::chrono::steady_clock::now();
    double k, s; pp->activeAtch.getKeff(k, s);
    double ta = std::chrono::duration<double>(t2 - t1).count(), ti = std::chrono::duration<double>(t1 - t0).count();
    std::printf("{\"keff\": %.8f, \"keff_std\": %.8f, \"pop\": %d, \"inactive\": %d, \"active\": %d, \"threads\": %d, "
                "\"t_inactive_s\": %.4f, \"t_active_s\": %.4f, \"neutrons_per_s\": %.6e, \"segments_per_s\": %.6e}\n",
                k, s, pp->pop, pp->N_inactive, pp->N_active, omp_get_max_threads(), ti, ta,
                (double)pp->pop * pp->N_active / ta, (double)(pp->nSegments - seg0) / ta);
  } catch (const std::exception& ex) { std::fprintf(stderr, "fatal: %s\n", ex.what()); return 1; }
  return 0;
}
#endif
