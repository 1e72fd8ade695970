// Host side of the drop-in boundary: turns SCONE input dictionaries into the flat arrays of
// include/scone_b200.h -- the job the Fortran shim does by walking SCONE's already-built objects
// (surfaceShelf, cellShelf, universeShelf, geomGraph, baseMgNeutronDatabase, tallyAdmin).
// Here the same build rules are applied directly to the dictionaries:
//   Geometry/csg_class.f90:76-207                       build order + checks
//   Geometry/Surfaces/*  Geometry/Cells/*  Geometry/Universes/*   per-type init
//   Geometry/geomGraph_class.f90:172-345                shrunk / extended graph, unique IDs
//   NuclearData/materialMenu_mod.f90:161-196            matIdx = order in materials{}
//   NuclearData/mgNeutronData/baseMgNeutron/*           XS rows, majorant
//   Tallies/tallyAdmin_class.f90:179-254                clerk order -> memory addresses
// Product code (no oracle involvement); emits indices 1-based as SCONE holds them.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include "../../../include/scone_b200.h"
#include "dict.hpp"
#include "named_grids.hpp"

namespace sb {

constexpr double kINF = 9223372036854775808.0;
constexpr double kSURF_TOL = 1.0e-12;
constexpr double kFP_REL_TOL = 1.0e-7;
constexpr double kTWO_PI = 6.283185307179586476925286766559;
constexpr int kMAX_NEST = 12;

using MatMap = std::map<std::string, int>;

// materialMenu_mod init: names in order of appearance + special keywords
inline MatMap materialMenu(const Dict& nuclearData, std::vector<std::string>* names = nullptr) {
  MatMap m; int i = 0;
  for (auto& n : nuclearData.getDict("materials").keys("dict")) { m[n] = ++i; if (names) names->push_back(n); }
  m["void"] = SB_VOID_MAT; m["outside"] = SB_OUTSIDE_MAT; m["overlap"] = SB_OVERLAP_MAT;
  return m;
}

// ---------------------------------------------------------------------------------------------
struct FlatGeometry {
  std::vector<int> surfType, surfId; std::vector<double> surfPar;
  std::vector<int> cellOff{0}, cellSurf, cellFill, cellId;
  std::vector<int> uniType, uniId, uniIpar; std::vector<double> uniDpar;
  std::vector<std::string> uniName;
  std::vector<std::vector<int>> fills;
  std::vector<double> auxD; std::vector<int> auxI;
  std::vector<int> graphIdx, graphId;
  int rootIdx = 0, borderIdx = 0, uniqueCells = 0, nesting = 0;
  int bc[6] = {0, 0, 0, 0, 0, 0};
  std::vector<int> usedMats;

  sb_geom_flat view() const {
    sb_geom_flat g{};
    g.n_surf = (int)surfType.size(); g.surf_type = surfType.data(); g.surf_par = surfPar.data();
    g.n_cell = (int)cellOff.size() - 1; g.cell_off = cellOff.data(); g.cell_surf = cellSurf.data();
    g.n_uni = (int)uniType.size(); g.uni_type = uniType.data(); g.uni_ipar = uniIpar.data(); g.uni_dpar = uniDpar.data();
    g.n_aux_d = (int)auxD.size(); g.aux_d = auxD.data(); g.n_aux_i = (int)auxI.size(); g.aux_i = auxI.data();
    g.n_graph = (int)graphIdx.size(); g.graph_idx = graphIdx.data(); g.graph_id = graphId.data();
    g.root_idx = rootIdx; g.border_idx = borderIdx;
    for (int i = 0; i < 6; ++i) g.bc[i] = bc[i];
    return g;
  }

  // geometryStd % activeMats (geometryStd_class.f90:521-547)
  std::vector<int> activeMats() const {
    int N = (int)usedMats.size();
    if (N == 0) return {};
    int last = usedMats[N - 1];
    if (last == SB_VOID_MAT) { if (--N == 0) return {}; last = usedMats[N - 1]; }
    if (last == SB_UNDEF_MAT) { if (--N == 0) return {}; last = usedMats[N - 1]; }
    if (last == SB_OVERLAP_MAT) --N;
    return std::vector<int>(usedMats.begin(), usedMats.begin() + N);
  }

 private:
  friend FlatGeometry buildGeometry(const Dict&, const MatMap&);
};

namespace detail {

inline int idxOf(const std::vector<int>& ids, int id, const char* what) {
  for (size_t i = 0; i < ids.size(); ++i) if (ids[i] == id) return (int)i + 1;
  throw FatalError(what, "There is no entity with ID: " + std::to_string(id));
}

inline int charToFill(const std::string& name, const MatMap& mats, const char* where) {
  if (name.size() > 2 && name[0] == 'u' && name[1] == '<') {
    size_t pos = name.rfind('>');
    if (pos != std::string::npos) {
      std::string num = name.substr(2, pos - 2);
      if (!Dict::isInt(num)) throw FatalError(where, "Failed to convert " + name + " to universe ID");
      int f = Dict::toInt(num);
      if (f <= 0) throw FatalError(where, "Universe ID must be +ve is: " + num);
      return -f;
    }
  }
  auto it = mats.find(name);
  if (it == mats.end()) throw FatalError(where, "Unknown material: " + name);
  return it->second;
}

// SharedModules/genericProcedures.f90:1098-1140
inline void rotationMatrix(double* m, double phi, double theta, double psi) {
  if (phi < 0.0 || phi >= 360.0) throw FatalError("rotationMatrix", "Angle phi must be in <0;360)");
  if (theta < 0.0 || theta > 180.0) throw FatalError("rotationMatrix", "Angle theta must be in <0;180>");
  if (psi < 0.0 || psi >= 360.0) throw FatalError("rotationMatrix", "Angle psi must be in <0;360)");
  double conv = kTWO_PI / 360.0;
  double sp = std::sin(phi * conv), cp = std::cos(phi * conv), st = std::sin(theta * conv), ct = std::cos(theta * conv);
  double ss = std::sin(psi * conv), cs = std::cos(psi * conv);
  m[0] = cs * cp - ct * sp * ss;  m[1] = cs * sp + ct * cp * ss;  m[2] = ss * sp;
  m[3] = -ss * cp - ct * sp * cs; m[4] = -ss * sp + ct * cp * cs; m[5] = cs * st;
  m[6] = st * sp;                 m[7] = -st * cp;                m[8] = ct;
}

}  // namespace detail

inline FlatGeometry buildGeometry(const Dict& d, const MatMap& mats) {
  using namespace detail;
  FlatGeometry g;
  // ---- surfaces (surfaceShelf: index = order of appearance) -----------------------------------
  const Dict& sd = d.getDict("surfaces");
  for (auto& name : sd.keys("dict")) {
    const Dict& s = sd.getDict(name);
    std::string type = s.getWord("type");
    int id = s.getInt("id");
    if (id <= 0) throw FatalError("new_surface", "Surface ID must be +ve");
    for (int o : g.surfId) if (o == id) throw FatalError("surfaceShelf init", "Surfaces have the same ID: " + std::to_string(id));
    double p[SB_SURF_NPAR] = {0, 0, 0, 0, 0, 0, kSURF_TOL, 0};
    int t;
    auto vec3 = [&](const char* key) { auto v = s.getRealArray(key); if (v.size() != 3) throw FatalError("surface init", std::string(key) + " must have size 3"); return v; };
    if (type == "xPlane" || type == "yPlane" || type == "zPlane") { t = SB_SURF_XPLANE + (type[0] - 'x'); p[0] = s.getReal(std::string(1, type[0]) + "0"); }
    else if (type == "plane") {
      t = SB_SURF_PLANE;
      auto c = s.getRealArray("coeffs");
      if (c.size() != 4) throw FatalError("init (plane)", "4 plane coefficients must be given");
      if (c[0] == 0.0 && c[1] == 0.0 && c[2] == 0.0) throw FatalError("init (plane)", "Invalid plane normal");
      double nrm = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
      for (int i = 0; i < 4; ++i) p[i] = c[i] / nrm;
    } else if (type == "sphere") {
      t = SB_SURF_SPHERE;
      auto o = vec3("origin"); double r = s.getReal("radius");
      if (r <= 0.0) throw FatalError("init (sphere)", "Radius of sphere must be +ve");
      p[0] = o[0]; p[1] = o[1]; p[2] = o[2]; p[3] = r; p[4] = r * r; p[6] = 2.0 * r * kSURF_TOL;
    } else if (type == "xCylinder" || type == "yCylinder" || type == "zCylinder") {
      t = SB_SURF_XCYL + (type[0] - 'x');
      auto o = vec3("origin"); double r = s.getReal("radius");
      if (r <= 0.0) throw FatalError("build (cylinder)", "Radius of cylinder must be +ve");
      p[0] = o[0]; p[1] = o[1]; p[2] = o[2]; p[3] = r; p[4] = r * r; p[6] = 2.0 * r * kSURF_TOL;
    } else if (type == "box" || type == "xSquareCylinder" || type == "ySquareCylinder" || type == "zSquareCylinder") {
      t = (type == "box") ? SB_SURF_BOX : SB_SURF_XSQCYL + (type[0] - 'x');
      auto o = vec3("origin"), hw = vec3("halfwidth");
      for (int a = 0; a < 3; ++a) {
        bool active = (t == SB_SURF_BOX) || (a != t - SB_SURF_XSQCYL);
        if (active && hw[a] < 0.0) throw FatalError("init (box)", "halfwidth cannot have -ve values.");
        p[a] = o[a]; p[3 + a] = hw[a];
      }
    } else if (type == "xTruncCylinder" || type == "yTruncCylinder" || type == "zTruncCylinder") {      // truncCylinder_class.f90:108-163
      t = SB_SURF_XTCYL + (type[0] - 'x');
      auto o = vec3("origin"); double r = s.getReal("radius"), a = s.getReal("halfwidth");
      if (r <= 0.0) throw FatalError("init (truncCylinder)", "Radius must be +ve");
      if (a <= 0.0) throw FatalError("init (truncCylinder)", "Halfwidth must be +ve");
      p[0] = o[0]; p[1] = o[1]; p[2] = o[2]; p[3] = r; p[4] = r * r; p[5] = a;
    } else throw FatalError("new_surface", "Unrecognised type of a surface: " + type);
    g.surfType.push_back(t); g.surfId.push_back(id);
    g.surfPar.insert(g.surfPar.end(), p, p + SB_SURF_NPAR);
  }
  // ---- cells (cellShelf) ------------------------------------------------------------------------
  const Dict& cd = d.getDict("cells");
  for (auto& name : cd.keys("dict")) {
    const Dict& c = cd.getDict(name);
    if (c.getWord("type") != "simpleCell") throw FatalError("new_cell", "Unsupported type of cell: " + c.getWord("type"));
    int id = c.getInt("id");
    for (int o : g.cellId) if (o == id) throw FatalError("cellShelf init", "Cells have the same ID: " + std::to_string(id));
    std::vector<int> seen;
    for (int sid : c.getIntArray("surfaces")) {
      int idx = idxOf(g.surfId, std::abs(sid), "surfaceShelf getIdx");
      if (std::find(seen.begin(), seen.end(), idx) != seen.end()) throw FatalError("init (simpleCell)", "There are repeated surfaces in definition of cell");
      seen.push_back(idx);
      g.cellSurf.push_back(sid < 0 ? -idx : idx);
    }
    g.cellOff.push_back((int)g.cellSurf.size());
    std::string filling = c.getWord("filltype");
    int f;
    if (filling == "outside") f = SB_OUTSIDE_MAT;
    else if (filling == "mat") {
      auto it = mats.find(c.getWord("material"));
      if (it == mats.end()) throw FatalError("cellShelf init", "Material with name " + c.getWord("material") + " was not found.");
      f = it->second;
    } else if (filling == "uni") { f = c.getInt("universe"); if (f <= 0) throw FatalError("cellShelf init", "Universe ID must be +ve"); f = -f; }
    else throw FatalError("cellShelf init", "Unknown type of cell filling: " + filling);
    g.cellFill.push_back(f); g.cellId.push_back(id);
  }
  // ---- universes (universeShelf: index = order of appearance) --------------------------------------
  const Dict& ud = d.getDict("universes");
  for (auto& name : ud.keys("dict")) {
    const Dict& u = ud.getDict(name);
    std::string type = u.getWord("type");
    int ip[SB_UNI_NIPAR] = {0, 0, 0, 0, 0, 0, 0, 0};
    double dp[SB_UNI_NDPAR]; for (auto& x : dp) x = 0.0;
    int id = u.getInt("id");
    if (id <= 0) throw FatalError("setupBase", "Universe ID must be +ve");
    for (int o : g.uniId) if (o == id) throw FatalError("universeShelf init", "Universes have the same ID: " + std::to_string(id));
    if (u.isPresent("origin")) { auto t = u.getRealArray("origin"); if (t.size() != 3) throw FatalError("setupBase", "Origin must have size 3"); for (int i = 0; i < 3; ++i) dp[i] = t[i]; }
    if (u.isPresent("rotation")) {
      auto t = u.getRealArray("rotation");
      if (t.size() != 3) throw FatalError("setupBase", "3 rotation angles must be given");
      if (!(t[0] == 0.0 && t[1] == 0.0 && t[2] == 0.0)) { ip[0] = 1; rotationMatrix(dp + 3, t[0], t[1], t[2]); }
    }
    if (u.isPresent("global")) ip[1] = u.getBool("global") ? 1 : 0;
    std::vector<int> fill;
    int t;
    if (type == "rootUniverse") {
      t = SB_UNI_ROOT;
      if (u.isPresent("origin")) throw FatalError("init (rootUniverse)", "Origin is not allowed.");
      if (u.isPresent("rotation")) throw FatalError("init (rootUniverse)", "Rotation is not allowed.");
      int b = u.getInt("border");
      if (b <= 0) throw FatalError("init (rootUniverse)", "Border must be given as +ve ID");
      ip[2] = idxOf(g.surfId, b, "surfaceShelf getIdx");
      fill = {charToFill(u.getWord("fill"), mats, "init (rootUniverse)"), SB_OUTSIDE_MAT};
    } else if (type == "pinUniverse") {
      t = SB_UNI_PIN;
      auto radii = u.getRealArray("radii"); auto names = u.getWordArray("fills");
      if (radii.size() != names.size()) throw FatalError("init (pinUniverse)", "Size of radii and fills does not match");
      for (double r : radii) if (r < 0.0) throw FatalError("init (pinUniverse)", "Found -ve value of radius.");
      int N = (int)radii.size();
      int m0 = (int)(std::min_element(radii.begin(), radii.end()) - radii.begin());
      if (radii[m0] != 0.0) throw FatalError("init (pinUniverse)", "Did not found outermost element with radius 0.0.");
      std::swap(radii[m0], radii[N - 1]); std::swap(names[m0], names[N - 1]);
      radii[N - 1] = kINF * 1.1;
      for (int i = N - 2; i >= 0; --i) {
        int m = (int)(std::max_element(radii.begin(), radii.begin() + i + 1) - radii.begin());
        std::swap(radii[m], radii[i]); std::swap(names[m], names[i]);
      }
      for (int i = 0; i + 1 < N; ++i) if (radii[i] == radii[i + 1]) throw FatalError("init (pinUniverse)", "Duplicate value of radius");
      ip[2] = N; ip[3] = (int)g.auxD.size();
      for (int i = 0; i < N; ++i) g.auxD.push_back(radii[i] * radii[i]);
      for (int i = 0; i < N; ++i) g.auxD.push_back(2.0 * radii[i] * kSURF_TOL);     // cylinder surfTol
      for (auto& n : names) fill.push_back(charToFill(n, mats, "init (pinUniverse)"));
    } else if (type == "latUniverse") {
      t = SB_UNI_LAT;
      bool offset = u.getBool("offset", true);
      auto p = u.getRealArray("pitch"); auto s = u.getIntArray("shape");
      if (p.size() != 3) throw FatalError("init (latUniverse)", "Pitch must have size 3");
      if (s.size() != 3) throw FatalError("init (latUniverse)", "Shape must have size 3");
      for (int i = 0; i < 3; ++i) if (s[i] < 0) throw FatalError("init (latUniverse)", "Shape contains -ve entries");
      if (s[2] == 0) { s[2] = 1; p[2] = 2.0 * kINF; }
      for (int i = 0; i < 3; ++i) if (s[i] == 0) throw FatalError("init (latUniverse)", "Shape in X and Y axis cannot be 0.");
      for (int i = 0; i < 3; ++i) if (p[i] < 10 * kSURF_TOL) throw FatalError("init (latUniverse)", "Pitch size must be larger than 10*SURF_TOL");
      for (int i = 0; i < 3; ++i) {
        dp[12 + i] = p[i];
        dp[18 + i] = p[i] * 0.5 - p[i] * kSURF_TOL;              // a_bar
        dp[15 + i] = -(s[i] * 0.5 * p[i]);                       // corner
        dp[21 + i] = std::fabs(dp[15 + i]);                      // outline box halfwidth
        ip[2 + i] = s[i];
      }
      int nCells = s[0] * s[1] * s[2];
      ip[5] = nCells + 1;
      auto m = u.getIntArray("map");
      if ((int)m.size() != nCells) throw FatalError("init (latUniverse)", "Lattice map size not equal to size implied by shape.");
      int nx = s[0], ncol = s[1] * s[2];
      auto flip = [&](std::vector<int>& a) { for (int j = 0; j < ncol / 2; ++j) for (int i = 0; i < nx; ++i) std::swap(a[i + j * nx], a[i + (ncol - 1 - j) * nx]); };
      flip(m);
      for (int v : m) fill.push_back(-v);
      fill.push_back(charToFill(u.getWord("padMat"), mats, "init (latUniverse)"));
      ip[6] = offset ? 1 : 0;
      if (u.isPresent("offsetMap")) {
        if (!offset) throw FatalError("init (latUniverse)", "Cannot have both an offset map and no offset.");
        auto om = u.getIntArray("offsetMap");
        if ((int)om.size() != nCells) throw FatalError("init (latUniverse)", "Offset map size not equal to size implied by shape.");
        flip(om);
        for (int v : om) if (v != 0 && v != 1) throw FatalError("init (latUniverse)", "Invalid entry to the offset map.");
        om.push_back(0);
        ip[6] = 2; ip[7] = (int)g.auxI.size();
        g.auxI.insert(g.auxI.end(), om.begin(), om.end());
      }
    } else if (type == "cellUniverse") {
      t = SB_UNI_CELL;
      auto cells = u.getIntArray("cells");
      ip[2] = (int)cells.size(); ip[3] = (int)g.auxI.size(); ip[4] = u.getBool("checkOverlap", false) ? 1 : 0;
      for (int cid : cells) { int ci = idxOf(g.cellId, cid, "cellShelf getIdx"); g.auxI.push_back(ci); fill.push_back(g.cellFill[ci - 1]); }
      fill.push_back(SB_UNDEF_MAT); fill.push_back(SB_OVERLAP_MAT);
    } else throw FatalError("new_universe", "Unrecognised type of universe: " + type);
    g.uniType.push_back(t); g.uniId.push_back(id); g.uniName.push_back(name);
    g.uniIpar.insert(g.uniIpar.end(), ip, ip + SB_UNI_NIPAR);
    g.uniDpar.insert(g.uniDpar.end(), dp, dp + SB_UNI_NDPAR);
    g.fills.push_back(fill);
  }
  if (g.uniType.empty()) throw FatalError("uniFills init", "Given not +ve number of universes");
  // IDs -> indices (uniFills finishBuild)
  for (auto& f : g.fills) for (auto& x : f) if (x < 0) x = -idxOf(g.uniId, -x, "universeShelf getIdx");
  // ---- root, border, BC ----------------------------------------------------------------------------
  int rootId = d.isPresent("root") ? d.getInt("root") : ud.getDict("root").getInt("id");
  g.rootIdx = idxOf(g.uniId, rootId, "universeShelf getIdx");
  if (g.uniType[g.rootIdx - 1] != SB_UNI_ROOT) throw FatalError("init (csg)", "Root universe is not type `rootUniverse`");
  g.borderIdx = g.uniIpar[(g.rootIdx - 1) * SB_UNI_NIPAR + 2];
  {
    auto BC = d.getIntArray("boundary");
    int bt = g.surfType[g.borderIdx - 1];
    if (bt >= SB_SURF_XTCYL) {                                       // truncCylinder%setBC: { a_min, a_max }, the radial face is vacuum
      if (BC.size() < 2) throw FatalError("setBC", "Wrong size of BC string. Must be at least 2");
      for (int i = 0; i < 2; ++i) { if (BC[i] < 0 || BC[i] > 2) throw FatalError("setBC", "Unrecognised BC"); g.bc[i] = BC[i]; }
      for (int i = 2; i < 6; ++i) g.bc[i] = 0;
    } else if (bt >= SB_SURF_BOX) {
      if (BC.size() < 6) throw FatalError("setBC", "Wrong size of BC string. Must be at least 6");
      for (int i = 0; i < 6; ++i) { if (BC[i] < 0 || BC[i] > 2) throw FatalError("setBC", "Unrecognised BC"); g.bc[i] = BC[i]; }
      for (int a = 0; a < 3; ++a) if ((g.bc[2 * a] == 2) != (g.bc[2 * a + 1] == 2)) throw FatalError("setBC", "Periodic BC need to be applied to oposite surfaces");
    } else {
      if (BC.empty()) throw FatalError("setBC", "At least one entry in the BC string is required!");
      if (BC[0] != 0) throw FatalError("setBC", "this surface supports only VACUUM BCs");
    }
  }
  // ---- structure checks (uniFills) ---------------------------------------------------------------------
  struct Rec {
    const std::vector<std::vector<int>>& f;
    bool cyc(int idx, std::vector<int>& path) const {
      for (int s : path) if (s == idx) return true;
      path.push_back(idx);
      for (int x : f[idx - 1]) if (x < 0 && cyc(-x, path)) return true;
      path.pop_back(); return false;
    }
    int depth(int idx) const { int N = 1; for (int x : f[idx - 1]) if (x < 0) N = std::max(N, 1 + depth(-x)); return N; }
    bool outsideBelow(int idx) const { for (int x : f[idx - 1]) { if (x == SB_OUTSIDE_MAT) return true; if (x < 0 && outsideBelow(-x)) return true; } return false; }
    void used(int idx, std::vector<char>& u) const { u[idx - 1] = 1; for (int x : f[idx - 1]) if (x < 0) used(-x, u); }
    void count(int idx, std::vector<long>& c) const { c[idx - 1] += 1; for (int x : f[idx - 1]) if (x < 0) count(-x, c); }
  } rec{g.fills};
  { std::vector<int> path; if (rec.cyc(g.rootIdx, path)) throw FatalError("init (csg)", "There is recursion in the geometry nesting. Universe cannot contain itself below itself."); }
  g.nesting = rec.depth(g.rootIdx);
  if (g.nesting > kMAX_NEST) throw FatalError("init (csg)", "Nesting level > max nesting");
  for (int x : g.fills[g.rootIdx - 1]) if (x < 0 && rec.outsideBelow(-x)) throw FatalError("init (csg)", "Cell with outside fill is present below root universe");
  // ---- geometry graph ----------------------------------------------------------------------------------
  std::string gtype = d.getDict("graph").getWord("type");
  auto layout = [&](int& top, int idx) {
    const auto& f = g.fills[idx - 1];
    if ((size_t)(top - 1) + f.size() > g.graphIdx.size()) throw FatalError("layoutUniverse", "Overflow of the location array");
    for (size_t i = 0; i < f.size(); ++i) g.graphIdx[top - 1 + i] = f[i];
    top += (int)f.size();
  };
  if (gtype == "shrunk") {
    std::vector<char> used(g.fills.size(), 0); rec.used(g.rootIdx, used);
    size_t N = 0; for (size_t i = 0; i < g.fills.size(); ++i) if (used[i]) N += g.fills[i].size();
    g.graphIdx.assign(N, 0); g.graphId.assign(N, 0);
    std::map<int, int> layed; int top = 1; layed[g.rootIdx] = top; layout(top, g.rootIdx);
    for (size_t loc = 1; loc <= N; ++loc) {
      int fill = g.graphIdx[loc - 1];
      if (fill < 0) {
        auto it = layed.find(-fill); int rootID;
        if (it == layed.end()) { rootID = top; layed[-fill] = top; layout(top, -fill); } else rootID = it->second;
        g.graphId[loc - 1] = rootID;
      }
    }
    if (top != (int)N + 1) throw FatalError("buildShrunk", "Did not reach the end of the location array");
  } else if (gtype == "extended") {
    std::vector<long> cnt(g.fills.size(), 0); rec.count(g.rootIdx, cnt);
    size_t N = 0; for (size_t i = 0; i < g.fills.size(); ++i) N += g.fills[i].size() * (size_t)cnt[i];
    if (N > (size_t)std::numeric_limits<int>::max()) throw FatalError("buildExtended", "geometry graph too large");
    g.graphIdx.assign(N, 0); g.graphId.assign(N, 0);
    int top = 1; layout(top, g.rootIdx);
    for (size_t loc = 1; loc <= N; ++loc) {
      int fill = g.graphIdx[loc - 1];
      if (fill < 0) { int rootID = top; layout(top, -fill); g.graphId[loc - 1] = rootID; }
    }
    if (top != (int)N + 1) throw FatalError("buildExtended", "Did not reach the end of the location array");
  } else throw FatalError("init (geomGraph)", "Unknown geometry graph type: " + gtype);
  // unique IDs + used materials (setUniqueIDs)
  {
    int c = 0; std::vector<int> um;
    for (size_t i = 0; i < g.graphIdx.size(); ++i)
      if (g.graphIdx[i] > 0) { g.graphId[i] = ++c; if (std::find(um.begin(), um.end(), g.graphIdx[i]) == um.end()) um.push_back(g.graphIdx[i]); }
    g.uniqueCells = c; std::sort(um.begin(), um.end()); g.usedMats = um;
  }
  return g;
}

// ---------------------------------------------------------------------------------------------
struct FlatMgData {
  int nMat = 0, nG = 0; bool isP1 = false; double collisionXS = 0.0;
  std::vector<std::string> names;
  std::vector<double> data, P0, prod, P1, chi, nu, majorant; std::vector<int> fissile;
  sb_mg_flat view() const {
    sb_mg_flat d{};
    d.n_mat = nMat; d.n_g = nG; d.data = data.data(); d.P0 = P0.data(); d.prod = prod.data(); d.P1 = isP1 ? P1.data() : nullptr;
    d.chi = chi.data(); d.fissile = fissile.data(); d.majorant = majorant.data(); d.collision_xs = collisionXS;
    return d;
  }
};

// baseMgNeutronDatabase init + activate (baseMgNeutronDatabase_class.f90:343-503) and
// baseMgNeutronMaterial init (baseMgNeutronMaterial_class.f90:186-291)
inline FlatMgData buildMgData(const Dict& nuclearData, const std::string& handle, const std::string& baseDir, const std::vector<int>& activeMats) {
  FlatMgData m;
  const Dict& h = nuclearData.getDict("handles").getDict(handle);
  if (h.getWord("type") != "baseMgNeutronDatabase") throw FatalError("ndReg_activate", "MG data must be of type baseMgNeutronDatabase");
  if (h.isPresent("avgDist")) {
    double t = h.getReal("avgDist");
    if (t <= 0.0) throw FatalError("init (baseMgNeutronDatabase)", "Must have a finite, positive minimum average collision distance");
    m.collisionXS = 1.0 / t;
  }
  std::string key = h.getWord("PN");
  if (key != "P0" && key != "P1") throw FatalError("init (baseMgNeutronMaterial)", "scatterKey: " + key + " is wrong. Must be P0 or P1");
  m.isP1 = key == "P1";
  materialMenu(nuclearData, &m.names);
  m.nMat = (int)m.names.size();
  const Dict& md = nuclearData.getDict("materials");
  for (int im = 0; im < m.nMat; ++im) {
    std::string path = md.getDict(m.names[im]).getWord("xsFile");
    if (!path.empty() && path[0] != '/') path = baseDir + "/" + path;
    Dict x = Dict::fromFile(path);
    int nG = x.getInt("numberOfGroups");
    if (nG < 1) throw FatalError("init (baseMgNeutronMaterial)", "Number of groups is invalid");
    if (im == 0) m.nG = nG; else if (nG != m.nG) throw FatalError("init (baseMgNeutronDatabase)", "Inconsistent # of groups in materials");
    bool fissile = x.isPresent("fission");
    auto need = [&](const char* k, size_t n) { auto v = x.getRealArray(k); if (v.size() != n) throw FatalError("init (baseMgNeutronMaterial)", std::string(k) + " has wrong size"); return v; };
    auto P0 = need("P0", (size_t)nG * nG), prod = need("scatteringMultiplicity", (size_t)nG * nG);
    std::vector<double> scat(nG, 0.0);
    for (int gi = 0; gi < nG; ++gi) { double s = 0.0; for (int go = 0; go < nG; ++go) s += P0[go + nG * gi]; scat[gi] = s; }
    std::vector<double> P1(nG * nG, 0.0);
    if (m.isP1) { P1 = need("P1", (size_t)nG * nG); for (int i = 0; i < nG * nG; ++i) P1[i] = (P0[i] != 0.0) ? P1[i] / P0[i] * 3.0 : 0.0; }
    auto cap = need("capture", nG);
    std::vector<double> fis(nG, 0.0), nu(nG, 0.0), chi(nG, 0.0), kap(nG, 0.0);
    if (fissile) {
      fis = need("fission", nG); nu = need("nu", nG); chi = need("chi", nG);
      double S = 0.0; for (double c : chi) S += c;
      if (std::fabs(S - 1.0) > 0.01 * kFP_REL_TOL) for (double& c : chi) c = c / S;
      if (x.isPresent("kappa")) kap = need("kappa", nG); else kap.assign(nG, (double)202.27f);   // KAPPA_DEFAULT is a default-REAL literal
    }
    for (int g = 0; g < nG; ++g) {
      double tot = scat[g] + cap[g];
      if (fissile) tot = tot + fis[g];
      double row[6] = {tot, scat[g], cap[g], fissile ? fis[g] : 0.0, fissile ? nu[g] * fis[g] : 0.0, fissile ? kap[g] * fis[g] : 0.0};
      m.data.insert(m.data.end(), row, row + 6);
    }
    m.P0.insert(m.P0.end(), P0.begin(), P0.end()); m.prod.insert(m.prod.end(), prod.begin(), prod.end());
    m.P1.insert(m.P1.end(), P1.begin(), P1.end()); m.chi.insert(m.chi.end(), chi.begin(), chi.end());
    m.nu.insert(m.nu.end(), nu.begin(), nu.end());
    m.fissile.push_back(fissile ? 1 : 0);
  }
  m.majorant.assign(m.nG, 0.0);                                          // initMajorant
  for (int g = 0; g < m.nG; ++g) {
    double xs = 0.0;
    for (int idx : activeMats) {
      if (idx < 1 || idx > m.nMat) throw FatalError("initMajorant", "active material index out of range");
      xs = std::max(xs, m.data[((size_t)(idx - 1) * m.nG + g) * 6]);
    }
    m.majorant[g] = xs * 1.0;
  }
  return m;
}

// ---------------------------------------------------------------------------------------------
// tallyAdmin init for the clerks the device scores (collisionClerk with space/material/energy/multi maps,
// flux/macro responses). Keeps the storage the sb_clerk pointers refer to.
struct TallyDefs {
  std::vector<sb_clerk> clerks; std::vector<std::string> names;
  std::vector<std::vector<double>> boundsStore; std::vector<std::vector<int>> matStore;
  int normClerk = 0; double normVal = 1.0; long size = 0; std::vector<long> addr, width;
};

namespace detail {
inline void gridEqual(const std::string& type, double mini, double maxi, int N, double& first, double& step) {   // grid_class.f90:35-84
  if (N < 1) throw FatalError("init_equalSpaced", "Number of bins must be +ve");
  if (std::fabs((maxi - mini) / maxi) < kFP_REL_TOL) throw FatalError("init_equalSpaced", "Minimum value must be smaller then maximum above realtive FP tolerance");
  first = mini;
  if (type == "lin") step = (maxi - mini) / N;
  else { if (mini <= 0) throw FatalError("init_equalSpaced", "For logarithmic grid minimum must be +ve"); step = std::log(maxi / mini) / N; }
}
inline void addMap1D(TallyDefs& T, sb_clerk& c, const Dict& d, const MatMap& mats, int nMat) {
  if (c.n_maps >= SB_MAX_MAPS) throw FatalError("init (multiMap)", "too many maps in one multiMap for the device tallies");
  sb_map1d m{}; std::string t = d.getWord("type");
  if (t == "spaceMap" || t == "energyMap") {
    m.type = (t == "spaceMap") ? SB_MAP_SPACE : SB_MAP_ENERGY;
    if (t == "spaceMap") { std::string ax = d.getWord("axis"); if (ax == "x") m.axis = 0; else if (ax == "y") m.axis = 1; else if (ax == "z") m.axis = 2; else throw FatalError("init (spaceMap)", "Unrecognised axis: " + ax); }
    std::string g = d.getWord("grid");
    if (g == "lin" || (g == "log" && t == "energyMap")) { m.grid = (g == "lin") ? SB_GRID_LIN : SB_GRID_LOG; m.n_bins = d.getInt("N"); gridEqual(g, d.getReal("min"), d.getReal("max"), m.n_bins, m.first, m.step); }
    else if (g == "unstruct" || (g == "predef" && t == "energyMap")) {
      std::vector<double> b;
      if (g == "predef") {                                          // energyMap%build_predef (energyMap_class.f90:137-177)
        b = namedEnergyGrid(d.getWord("name"));
        if (b.empty()) throw FatalError("build_predef (energyMap)", "Grid " + d.getWord("name") + " is undefined!");
      } else b = d.getRealArray("bins");
      if (t == "energyMap") std::sort(b.begin(), b.end());
      if (b.size() < 2) throw FatalError("init_unstruct", "Empty array or array of size 1 was provided");
      for (size_t i = 1; i < b.size(); ++i) if (b[i] < b[i - 1]) throw FatalError("init_unstruct", "Provided grid is not sorted");
      m.grid = SB_GRID_UNSTRUCT; m.n_bins = (int)b.size() - 1; m.first = b[0];
      T.boundsStore.push_back(b);
    } else throw FatalError("init (map)", "'grid' keyword is not supported: " + g);
  } else if (t == "materialMap") {
    m.type = SB_MAP_MATERIAL;
    auto names = d.getWordArray("materials");
    std::string undef = d.getWord("undefBin", "false");
    bool track;
    if (undef == "yes" || undef == "y" || undef == "true" || undef == "TRUE" || undef == "T") track = true;
    else if (undef == "no" || undef == "n" || undef == "false" || undef == "FALSE" || undef == "F") track = false;
    else throw FatalError("init (materialMap)", undef + " is an unrecognised entry!");
    int N = (int)names.size();
    m.default_bin = track ? N + 1 : 0; m.n_bins = track ? N + 1 : N;
    std::vector<int> tab(nMat, m.default_bin);
    for (int i = 0; i < N; ++i) {
      auto it = mats.find(names[i]);
      if (it == mats.end()) throw FatalError("build (materialMap)", "Material " + names[i] + " does not exist in the input materials");
      if (it->second >= 1 && it->second <= nMat) tab[it->second - 1] = i + 1;
    }
    T.matStore.push_back(tab);
  } else throw FatalError("new_tallyMap", "tallyMap type not supported by the device tallies: " + t);
  c.maps[c.n_maps++] = m;
}
}  // namespace detail

inline TallyDefs buildTallies(const Dict& d, const MatMap& mats, int nMat) {
  TallyDefs T;
  T.boundsStore.reserve(64); T.matStore.reserve(64);
  long memLoc = 1;
  for (auto& n : d.keys("dict")) {
    const Dict& cd = d.getDict(n);
    std::string t = cd.getWord("type");
    if (t == "keffAnalogClerk" || t == "keffImplicitClerk") {      // k-eff clerks named by the deck itself: 3 / 5 bins, no maps, no responses
      sb_clerk c{};
      c.kind = (t == "keffAnalogClerk") ? SB_CLERK_KEFF_ANALOG : SB_CLERK_KEFF_IMPLICIT;
      c.handle_virtual = 1;
      if (t == "keffImplicitClerk" && !cd.getBool("handleVirtual", true)) throw FatalError("init (keffImplicitClerk)", "handleVirtual 0 is not supported by the device tallies");
      const int w = (c.kind == SB_CLERK_KEFF_ANALOG) ? 3 : 5;
      T.addr.push_back(memLoc); T.width.push_back(w);
      memLoc += w;
      T.clerks.push_back(c); T.names.push_back(n);
      continue;
    }
    if (t == "shannonEntropyClerk") {                               // shannonEntropyClerk_class.f90:75-92: one map, `cycles`
      sb_clerk c{};
      c.kind = SB_CLERK_SHANNON; c.handle_virtual = 1;
      const Dict& md = cd.getDict("map");
      if (md.getWord("type") == "multiMap") for (auto& mn : md.getWordArray("maps")) detail::addMap1D(T, c, md.getDict(mn), mats, nMat);
      else detail::addMap1D(T, c, md, mats, nMat);
      c.cycles = cd.getInt("cycles");
      if (c.cycles < 0) throw FatalError("init (shannonEntropyClerk)", "-ve number of cycles");
      long bins = 1; for (int i = 0; i < c.n_maps; ++i) bins *= c.maps[i].n_bins;
      T.addr.push_back(memLoc); T.width.push_back(1);
      memLoc += bins + 1 + c.cycles;
      T.clerks.push_back(c); T.names.push_back(n);
      continue;
    }
    if (t != "collisionClerk" && t != "trackClerk") throw FatalError("new_tallyClerk", "tallyClerk type not supported by the device tallies: " + t);
    if (cd.isPresent("filter")) throw FatalError("init (" + t + ")", "tally filters are not supported by the device tallies");
    sb_clerk c{};
    c.kind = (t == "trackClerk") ? SB_CLERK_TRACK : SB_CLERK_COLLISION;
    if (cd.isPresent("map")) {
      const Dict& md = cd.getDict("map");
      if (md.getWord("type") == "multiMap") for (auto& mn : md.getWordArray("maps")) detail::addMap1D(T, c, md.getDict(mn), mats, nMat);
      else detail::addMap1D(T, c, md, mats, nMat);
    }
    for (auto& rn : cd.getWordArray("response")) {
      if (c.n_resp >= SB_MAX_RESP) throw FatalError("init (collisionClerk)", "too many responses for the device tallies");
      const Dict& rd = cd.getDict(rn);
      std::string rt = rd.getWord("type");
      int mt;
      if (rt == "fluxResponse") mt = 0;
      else if (rt == "macroResponse") {
        mt = rd.getInt("MT");
        if (mt > 0) {
          switch (mt) { case 1: mt = -1; break; case 2: mt = -3; break; case 3: mt = -22; break; case 101: mt = -2; break; case 18: mt = -6; break; case 27: mt = -21; break; case 301: mt = -80; break;
            default: throw FatalError("build (macroResponse)", "MT outside the main data is not available for MG data"); }
        }
      } else throw FatalError("new_tallyResponse", "tallyResponse type not supported by the device tallies: " + rt);
      c.resp_mt[c.n_resp++] = mt;
    }
    c.handle_virtual = cd.getBool("handleVirtual", true) ? 1 : 0;
    long bins = 1; for (int i = 0; i < c.n_maps; ++i) bins *= c.maps[i].n_bins;
    T.addr.push_back(memLoc); T.width.push_back(c.n_resp);
    memLoc += bins * c.n_resp;
    T.clerks.push_back(c); T.names.push_back(n);
  }
  T.size = memLoc - 1;
  if (d.getInt("batchSize", 1) != 1) throw FatalError("init (tallyAdmin)", "batchSize other than 1 is not supported by the device tallies");
  // fix up pointers into the stores (stable because of reserve)
  size_t ib = 0, im = 0;
  for (auto& c : T.clerks) for (int i = 0; i < c.n_maps; ++i) {
    if (c.maps[i].type == SB_MAP_MATERIAL) c.maps[i].mat_bin = T.matStore[im++].data();
    else if (c.maps[i].grid == SB_GRID_UNSTRUCT) c.maps[i].bounds = T.boundsStore[ib++].data();
  }
  if (d.isPresent("norm")) {
    std::string nn = d.getWord("norm"); T.normVal = d.getReal("normVal");
    for (size_t i = 0; i < T.names.size(); ++i) if (T.names[i] == nn) T.normClerk = (int)i + 1;
    if (!T.normClerk) throw FatalError("init (tallyAdmin)", "norm clerk is not defined: " + nn);
  }
  return T;
}

}  // namespace sb
