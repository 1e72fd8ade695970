// ORACLE (test infrastructure, not product code).
// CPU restatement of SCONE's CSG geometry: surfaces, cells, universes, the
// geometry graph, coordList and the geometryStd run-time procedures.
// Object-per-entity, virtual dispatch, one particle at a time -- as the reference.
// Every function cites the reference lines it follows.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../scone_b200/csrc/host/dict.hpp"   // input grammar only (no arithmetic)
#include "../scone_b200/csrc/host/named_grids.hpp"   // input data only (the named energy-group structures)

namespace orc {

using sb::Dict;
using sb::FatalError;

// SharedModules/universalVariables.f90:22-27,43-47,58-60 ; numPrecision.f90:28
constexpr double INF = 9223372036854775808.0;   // 2^63
constexpr double SURF_TOL = 1.0e-12;
constexpr double NUDGE = 1.0e-8;
constexpr double FP_REL_TOL = 1.0e-7;
constexpr int OUTSIDE_MAT = 0;
constexpr int VOID_MAT = std::numeric_limits<int32_t>::max();
constexpr int UNDEF_MAT = VOID_MAT - 1;
constexpr int OVERLAP_MAT = VOID_MAT - 2;
constexpr int VACUUM_BC = 0, REFLECTIVE_BC = 1, PERIODIC_BC = 2;
constexpr int COLL_EV = 1, BOUNDARY_EV = 2, CROSS_EV = 3, LOST_EV = 4, FIELD_EV = 5;
constexpr int MAX_NEST = 12;
constexpr double TWO_PI = 6.283185307179586476925286766559;   // numPrecision.f90 PI*2

struct Vec3 {
  double v[3] = {0, 0, 0};
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
};

inline double fsign(double a, double b) { return std::copysign(std::fabs(a), b); }   // Fortran SIGN

// SharedModules/genericProcedures.f90:1098-1140  (ZXZ Euler angles in degrees)
inline void rotationMatrix(double m[3][3], double phi, double theta, double psi) {
  if (phi < 0.0 || phi >= 360.0) throw FatalError("rotationMatrix", "Angle phi must be in <0;360)");
  if (theta < 0.0 || theta > 180.0) throw FatalError("rotationMatrix", "Angle theta must be in <0;180>");
  if (psi < 0.0 || psi >= 360.0) throw FatalError("rotationMatrix", "Angle psi must be in <0;360)");
  double conv = TWO_PI / 360.0;
  double sp = std::sin(phi * conv), cp = std::cos(phi * conv);
  double st = std::sin(theta * conv), ct = std::cos(theta * conv);
  double ss = std::sin(psi * conv), cs = std::cos(psi * conv);
  m[0][0] = cs * cp - ct * sp * ss;  m[0][1] = cs * sp + ct * cp * ss;   m[0][2] = ss * sp;
  m[1][0] = -ss * cp - ct * sp * cs; m[1][1] = -ss * sp + ct * cp * cs;  m[1][2] = cs * st;
  m[2][0] = st * sp;                 m[2][1] = -st * cp;                 m[2][2] = ct;
}
inline Vec3 matvec(const double m[3][3], const Vec3& a) {
  Vec3 r;
  for (int i = 0; i < 3; ++i) r[i] = m[i][0] * a[0] + m[i][1] * a[1] + m[i][2] * a[2];
  return r;
}

// ===========================================================================
// Surfaces  (Geometry/Surfaces/surface_inter.f90:363-414)
// ===========================================================================
struct Surface {
  int id = -1;
  double tol = SURF_TOL;
  virtual ~Surface() = default;
  virtual std::string myType() const = 0;
  virtual double evaluate(const Vec3& r) const = 0;
  virtual double distance(const Vec3& r, const Vec3& u) const = 0;
  virtual bool going(const Vec3& r, const Vec3& u) const = 0;
  virtual void boundingBox(double aabb[6]) const = 0;
  virtual void setBC(const std::vector<int>& bc) {
    if (bc.empty()) throw FatalError("setBC", "At least one entry in the BC string is required");
    if (bc[0] != VACUUM_BC) throw FatalError("setBC", myType() + " supports only VACUUM BCs");
  }
  virtual void explicitBC(Vec3&, Vec3&) const {}
  virtual void transformBC(Vec3&, Vec3&) const {}
  // surface_inter.f90:363-377
  bool halfspace(const Vec3& r, const Vec3& u) const {
    double c = evaluate(r);
    bool hs = c > 0.0;
    if (std::fabs(c) < tol) hs = going(r, u);
    return hs;
  }
};

// QuadSurfaces/aPlane_class.f90
struct APlane : Surface {
  int axis = 0; double a0 = 0;
  std::string myType() const override { return axis == 0 ? "xPlane" : axis == 1 ? "yPlane" : "zPlane"; }
  double evaluate(const Vec3& r) const override { return r[axis] - a0; }
  double distance(const Vec3& r, const Vec3& u) const override {
    double ra = a0 - r[axis], ua = u[axis], d;
    if (std::fabs(ra) < tol) d = INF;
    else if (ua != 0.0) d = ra / ua;
    else d = INF;
    if (d <= 0.0 || d > INF) d = INF;
    return d;
  }
  bool going(const Vec3& r, const Vec3& u) const override {
    double ua = u[axis];
    bool hs = ua > 0.0;
    if (ua == 0.0) hs = (r[axis] - a0) >= 0.0;
    return hs;
  }
  void boundingBox(double b[6]) const override {
    for (int i = 0; i < 3; ++i) { b[i] = -INF; b[i + 3] = INF; }
    b[axis] = a0; b[axis + 3] = a0;
  }
};

// QuadSurfaces/plane_class.f90
struct Plane : Surface {
  double n[3] = {0, 0, 0}; double offset = 0;
  std::string myType() const override { return "plane"; }
  double evaluate(const Vec3& r) const override { return (r[0] * n[0] + r[1] * n[1] + r[2] * n[2]) - offset; }
  double distance(const Vec3& r, const Vec3& u) const override {
    double k = u[0] * n[0] + u[1] * n[1] + u[2] * n[2];
    double c = evaluate(r), d;
    if (k == 0.0 || std::fabs(c) < tol) d = INF;
    else { d = -c / k; if (d <= 0.0 || d > INF) d = INF; }
    return d;
  }
  bool going(const Vec3& r, const Vec3& u) const override {
    double proj = u[0] * n[0] + u[1] * n[1] + u[2] * n[2];
    bool hs = proj > 0.0;
    if (proj == 0.0) hs = evaluate(r) >= 0.0;
    return hs;
  }
  void boundingBox(double b[6]) const override { for (int i = 0; i < 3; ++i) { b[i] = -INF; b[i + 3] = INF; } }
};

// QuadSurfaces/sphere_class.f90
struct Sphere : Surface {
  double o[3] = {0, 0, 0}, r = 0, r_sq = 0;
  std::string myType() const override { return "sphere"; }
  double evaluate(const Vec3& p) const override {
    double d0 = p[0] - o[0], d1 = p[1] - o[1], d2 = p[2] - o[2];
    return (d0 * d0 + d1 * d1 + d2 * d2) - r_sq;
  }
  double distance(const Vec3& p, const Vec3& u) const override {
    double c = evaluate(p);
    double k = (p[0] - o[0]) * u[0] + (p[1] - o[1]) * u[1] + (p[2] - o[2]) * u[2];
    double delta = k * k - c, d;
    if (delta < 0.0) d = INF;
    else if (std::fabs(c) < tol) { if (k >= 0.0) d = INF; else d = -k + std::sqrt(delta); }
    else if (c < 0.0) d = -k + std::sqrt(delta);
    else { d = -k - std::sqrt(delta); if (d <= 0.0) d = INF; }
    return d;
  }
  bool going(const Vec3& p, const Vec3& u) const override {
    return ((p[0] - o[0]) * u[0] + (p[1] - o[1]) * u[1] + (p[2] - o[2]) * u[2]) >= 0.0;
  }
  void boundingBox(double b[6]) const override { for (int i = 0; i < 3; ++i) { b[i] = o[i] - r; b[i + 3] = o[i] + r; } }
};

// QuadSurfaces/cylinder_class.f90:104-289
struct Cylinder : Surface {
  int axis = 2, p0 = 0, p1 = 1;
  double o[3] = {0, 0, 0}, r = 0, r_sq = 0;
  void build(int id_, const std::string& type, const double origin[3], double radius) {
    if (id_ < 1) throw FatalError("cylinder build", "Invalid surface id");
    if (radius <= 0.0) throw FatalError("cylinder build", "Radius of cylinder must be +ve");
    if (type == "xCylinder") { axis = 0; p0 = 1; p1 = 2; }
    else if (type == "yCylinder") { axis = 1; p0 = 0; p1 = 2; }
    else if (type == "zCylinder") { axis = 2; p0 = 0; p1 = 1; }
    else throw FatalError("cylinder build", "Unknown type of cylinder: " + type);
    r = radius; r_sq = radius * radius;
    for (int i = 0; i < 3; ++i) o[i] = origin[i];
    id = id_;
    tol = 2.0 * r * SURF_TOL;
  }
  std::string myType() const override { return axis == 0 ? "xCylinder" : axis == 1 ? "yCylinder" : "zCylinder"; }
  double evaluate(const Vec3& p) const override {
    double d0 = p[p0] - o[p0], d1 = p[p1] - o[p1];
    return (d0 * d0 + d1 * d1) - r_sq;
  }
  double distance(const Vec3& p, const Vec3& u) const override {
    double c = evaluate(p);
    double k = (p[p0] - o[p0]) * u[p0] + (p[p1] - o[p1]) * u[p1];
    double a = 1.0 - u[axis] * u[axis];
    double delta = k * k - a * c, d;
    if (delta < 0.0 || a == 0.0) d = INF;
    else if (std::fabs(c) < tol) {
      if (k >= 0.0) d = INF;
      else { d = -k + std::sqrt(delta); d = d / a; }
    } else if (c < 0.0) { d = -k + std::sqrt(delta); d = d / a; }
    else { d = -k - std::sqrt(delta); d = d / a; if (d <= 0.0) d = INF; }
    return std::min(d, INF);
  }
  bool going(const Vec3& p, const Vec3& u) const override {
    return ((p[p0] - o[p0]) * u[p0] + (p[p1] - o[p1]) * u[p1]) >= 0.0;
  }
  void boundingBox(double b[6]) const override {
    b[p0] = o[p0] - r; b[p1] = o[p1] - r; b[p0 + 3] = o[p0] + r; b[p1 + 3] = o[p1] + r;
    b[axis] = -INF; b[axis + 3] = INF;
  }
};

// CompositeSurfaces/box_class.f90 and squareCylinder_class.f90.
// A squareCylinder is the same arithmetic restricted to its two in-plane axes,
// so one class carries a list of active axes (nax = 3 for a box, 2 otherwise).
struct BoxLike : Surface {
  int nax = 3; int ax[3] = {0, 1, 2};
  double o[3] = {0, 0, 0}, hw[3] = {0, 0, 0};    // indexed by active-axis slot
  int bc[6] = {0, 0, 0, 0, 0, 0};
  std::string type = "box";
  std::string myType() const override { return type; }
  double evaluate(const Vec3& r) const override {       // box_class.f90:134-146
    double c = -std::numeric_limits<double>::max();
    for (int i = 0; i < nax; ++i) c = std::max(c, std::fabs(r[ax[i]] - o[i]) - hw[i]);
    return c;
  }
  double distance(const Vec3& r, const Vec3& u) const override {   // box_class.f90:165-237
    const double FP_MISS_TOL = 1.0 + 10.0 * std::numeric_limits<double>::epsilon();
    double far = std::numeric_limits<double>::max(), near = -std::numeric_limits<double>::max();
    for (int i = 0; i < nax; ++i) {
      int a = ax[i];
      double rb = r[a] - o[i];
      double a_far = fsign(hw[i], u[a]);
      double a_near = -a_far;
      double test_near, test_far;
      if (u[a] != 0.0) {
        test_near = (a_near - rb) / u[a];
        test_far = (a_far - rb) / u[a];
      } else {
        test_near = fsign(INF, a_near - rb);
        test_far = fsign(INF, a_far - rb);
        if (test_near > test_far) std::swap(test_near, test_far);
      }
      far = std::min(far, test_far);
      near = std::max(near, test_near);
    }
    double d;
    if (far <= near * FP_MISS_TOL) d = INF;
    else if (std::fabs(evaluate(r)) < tol) d = (std::fabs(far) >= std::fabs(near)) ? far : near;
    else d = (near <= 0.0) ? far : near;
    if (d <= 0.0 || d > INF) d = INF;
    return d;
  }
  bool going(const Vec3& r, const Vec3& u) const override {   // box_class.f90:252-279
    int maxCom = 0; double best = 0;
    for (int i = 0; i < nax; ++i) {
      double val = std::fabs(r[ax[i]] - o[i]) - hw[i];
      if (i == 0 || val > best) { best = val; maxCom = i; }   // maxloc: first maximum
    }
    double rl = r[ax[maxCom]] - o[maxCom];
    double proj = u[ax[maxCom]] * fsign(1.0, rl);
    bool hs = proj > 0.0;
    if (proj == 0.0) hs = evaluate(r) >= 0.0;
    return hs;
  }
  void boundingBox(double b[6]) const override {
    for (int i = 0; i < 3; ++i) { b[i] = -INF; b[i + 3] = INF; }
    for (int i = 0; i < nax; ++i) { b[ax[i]] = o[i] - hw[i]; b[ax[i] + 3] = o[i] + hw[i]; }
  }
  void setBC(const std::vector<int>& BC) override {             // box_class.f90:342-372
    if (BC.size() < 6) throw FatalError("setBC", "Wrong size of BC string. Must be at least 6");
    for (int i = 0; i < 6; ++i) {
      if (BC[i] != VACUUM_BC && BC[i] != REFLECTIVE_BC && BC[i] != PERIODIC_BC)
        throw FatalError("setBC", "Unrecognised BC");
      bc[i] = BC[i];
    }
    for (int a = 0; a < 3; ++a)
      if ((bc[2 * a] == PERIODIC_BC) != (bc[2 * a + 1] == PERIODIC_BC))
        throw FatalError("setBC", "Periodic BC need to be applied to oposite surfaces");
  }
  void explicitBC(Vec3& r, Vec3& u) const override {            // box_class.f90:380-417
    for (int i = 0; i < nax; ++i) {
      int a = ax[i];
      double r0 = r[a] - o[i];
      if (std::fabs(r0) <= hw[i] * (1.0 - tol)) continue;
      int b = (r0 < 0.0) ? bc[2 * a] : bc[2 * a + 1];
      if (b == REFLECTIVE_BC) u[a] = -u[a];
      else if (b == PERIODIC_BC) r[a] = r[a] - 2.0 * fsign(hw[i], r0);
    }
  }
  void transformBC(Vec3& r, Vec3& u) const override {           // box_class.f90:432-487
    for (int i = 0; i < nax; ++i) {
      int a = ax[i];
      double a_bar = hw[i] * (1.0 - tol);
      int Ri = (int)std::ceil(std::fabs(r[a] - o[i]) / a_bar) / 2;
      for (int t = 1; t <= Ri; ++t) {
        double r0 = r[a] - o[i];
        int b = (r0 < 0.0) ? bc[2 * a] : bc[2 * a + 1];
        if (b == REFLECTIVE_BC) {
          double a0 = fsign(hw[i], r0) + o[i];
          double d = r[a] - a0;
          r[a] = r[a] - 2.0 * d;
          u[a] = -u[a];
        } else if (b == PERIODIC_BC) {
          double d = fsign(hw[i], r0);
          r[a] = r[a] - 2.0 * d;
        }
      }
    }
  }
};

// Geometry/Surfaces/surfaceFactory_func.f90 + per-class init
// CompositeSurfaces/truncCylinder_class.f90: finite cylinder along x, y or z; F(r) = max[ (rho^2 - R^2) / (2R), |r_ax - o_ax| - a ];
// BCs on the two axial faces only { a_min, a_max }, the radial face is always vacuum
struct TruncCylinder : Surface {
  int axis = 2, p0 = 0, p1 = 1;
  double o[3] = {0, 0, 0}, a = 0, r = 0;
  int BC[2] = {VACUUM_BC, VACUUM_BC};
  std::string myType() const override { return axis == 0 ? "xTruncCylinder" : axis == 1 ? "yTruncCylinder" : "zTruncCylinder"; }
  double evaluate(const Vec3& p) const override {                   // :204-219
    double d0 = p[p0] - o[p0], d1 = p[p1] - o[p1];
    double c = ((d0 * d0 + d1 * d1) - r * r) / r * 0.5;
    return std::max(c, std::fabs(p[axis] - o[axis]) - a);
  }
  double distance(const Vec3& p, const Vec3& u) const override {    // :231-309
    const double FP_MISS_TOL = 1.0 + 10.0 * std::numeric_limits<double>::epsilon();
    double d0 = p[p0] - o[p0], d1 = p[p1] - o[p1];
    double c1 = (d0 * d0 + d1 * d1) - r * r;
    double k = d0 * u[p0] + d1 * u[p1];
    double aa = 1.0 - u[axis] * u[axis];
    double delta = k * k - aa * c1;
    double far, near;
    if (delta <= 0.0 || aa == 0.0) { far = INF; near = std::copysign(INF, c1); }
    else {
      far = (-k + std::sqrt(delta)) / aa;
      near = (-k - std::sqrt(delta)) / aa;
      if (far < near) std::swap(far, near);
    }
    double rb = p[axis] - o[axis], tn, tf;
    if (u[axis] != 0.0) { tn = (-a - rb) / u[axis]; tf = (a - rb) / u[axis]; }
    else { tn = std::copysign(INF, -a - rb); tf = std::copysign(INF, a - rb); }
    if (tf < tn) std::swap(tf, tn);
    far = std::min(far, tf); near = std::max(near, tn);
    double c = std::max(c1 / r * 0.5, std::fabs(rb) - a), d;
    if (far <= near * FP_MISS_TOL) d = INF;
    else if (std::fabs(c) < tol) d = (std::fabs(far) >= std::fabs(near)) ? far : near;
    else d = (near <= 0.0) ? far : near;
    if (d <= 0.0 || d > INF) d = INF;
    return d;
  }
  bool going(const Vec3& p, const Vec3& u) const override {          // :320-359
    double rp0 = p[p0] - o[p0], rp1 = p[p1] - o[p1];
    double c1 = ((rp0 * rp0 + rp1 * rp1) - r * r) / r * 0.5;
    double rv = p[axis] - o[axis];
    double c2 = std::fabs(rv) - a, proj, c;
    if (c1 >= 2.0 * r * c2) { proj = u[p0] * rp0 + u[p1] * rp1; c = c1; }
    else { proj = u[axis] * rv; c = c2; }
    bool hs = proj > 0.0;
    if (proj == 0.0) hs = c >= 0.0;
    return hs;
  }
  void boundingBox(double b[6]) const override {
    b[axis] = o[axis] - a; b[axis + 3] = o[axis] + a;
    b[p0] = o[p0] - r; b[p1] = o[p1] - r; b[p0 + 3] = o[p0] + r; b[p1 + 3] = o[p1] + r;
  }
  void setBC(const std::vector<int>& bc) override {                 // :432-460
    if (bc.size() < 2) throw FatalError("setBC (truncCylinder)", "Wrong size of BC string. Must be at least 2");
    for (int i = 0; i < 2; ++i) {
      if (bc[i] != VACUUM_BC && bc[i] != REFLECTIVE_BC && bc[i] != PERIODIC_BC) throw FatalError("setBC (truncCylinder)", "Unrecognised BC");
      BC[i] = bc[i];
    }
  }
  void explicitBC(Vec3& p, Vec3& u) const override {                // :468-503
    double r0 = p[axis] - o[axis];
    if (std::fabs(r0) <= a - tol) return;
    int bc = (r0 < 0.0) ? BC[0] : BC[1];
    if (bc == REFLECTIVE_BC) u[axis] = -u[axis];
    else if (bc == PERIODIC_BC) p[axis] = p[axis] - 2.0 * std::copysign(a, r0);
  }
  void transformBC(Vec3& p, Vec3& u) const override {               // :511-562
    double a_bar = a - tol;
    int Ri = (int)std::ceil(std::fabs(p[axis] - o[axis]) / a_bar) / 2;
    for (int t = 1; t <= Ri; ++t) {
      double r0 = p[axis] - o[axis];
      int bc = (r0 < 0.0) ? BC[0] : BC[1];
      if (bc == REFLECTIVE_BC) {
        double a0 = std::copysign(a, r0) + o[axis];
        double d = p[axis] - a0;
        p[axis] = p[axis] - 2.0 * d;
        u[axis] = -u[axis];
      } else if (bc == PERIODIC_BC) {
        double d = std::copysign(a, r0);
        p[axis] = p[axis] - 2.0 * d;
      }
    }
  }
};

inline std::unique_ptr<Surface> newSurface(const Dict& d) {
  std::string type = d.getWord("type");
  int id = d.getInt("id");
  if (id <= 0) throw FatalError("new_surface", "ID must be +ve");
  if (type == "xPlane" || type == "yPlane" || type == "zPlane") {
    auto s = std::make_unique<APlane>();
    s->id = id;
    s->axis = type[0] - 'x';
    s->a0 = d.getReal(std::string(1, type[0]) + "0");
    return s;
  }
  if (type == "plane") {
    auto s = std::make_unique<Plane>();
    s->id = id;
    auto c = d.getRealArray("coeffs");
    if (c.size() != 4) throw FatalError("plane init", "4 plane coefficients must be given");
    if (c[0] == 0.0 && c[1] == 0.0 && c[2] == 0.0) throw FatalError("plane init", "Invalid plane normal");
    double nrm = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    for (auto& x : c) x = x / nrm;
    s->n[0] = c[0]; s->n[1] = c[1]; s->n[2] = c[2]; s->offset = c[3];
    return s;
  }
  if (type == "sphere") {
    auto s = std::make_unique<Sphere>();
    s->id = id;
    s->r = d.getReal("radius");
    auto o = d.getRealArray("origin");
    if (o.size() != 3) throw FatalError("sphere init", "Origin needs to have size 3");
    if (s->r <= 0.0) throw FatalError("sphere init", "Radius of sphere must be +ve");
    s->r_sq = s->r * s->r;
    for (int i = 0; i < 3; ++i) s->o[i] = o[i];
    s->tol = 2.0 * s->r * SURF_TOL;
    return s;
  }
  if (type == "xCylinder" || type == "yCylinder" || type == "zCylinder") {
    auto s = std::make_unique<Cylinder>();
    auto o = d.getRealArray("origin");
    if (o.size() != 3) throw FatalError("cylinder init", "Origin needs to have size 3");
    s->build(id, type, o.data(), d.getReal("radius"));
    return s;
  }
  if (type == "box" || type == "xSquareCylinder" || type == "ySquareCylinder" || type == "zSquareCylinder") {
    auto s = std::make_unique<BoxLike>();
    s->id = id; s->type = type;
    auto o = d.getRealArray("origin");
    auto h = d.getRealArray("halfwidth");
    if (o.size() != 3) throw FatalError("box init", "origin must have size 3");
    if (h.size() != 3) throw FatalError("box init", "halfwidth must have size 3");
    if (type == "box") { s->nax = 3; s->ax[0] = 0; s->ax[1] = 1; s->ax[2] = 2; }
    else {
      s->nax = 2;
      int axis = type[0] - 'x';
      int k = 0;
      for (int a = 0; a < 3; ++a) if (a != axis) s->ax[k++] = a;
    }
    for (int i = 0; i < s->nax; ++i) {
      s->o[i] = o[s->ax[i]]; s->hw[i] = h[s->ax[i]];
      if (s->hw[i] < 0.0) throw FatalError("box init", "halfwidth cannot have -ve values.");
    }
    return s;
  }
  if (type == "xTruncCylinder" || type == "yTruncCylinder" || type == "zTruncCylinder") {    // truncCylinder_class.f90:108-163
    auto s = std::make_unique<TruncCylinder>();
    s->id = id;
    auto o = d.getRealArray("origin");
    if (o.size() != 3) throw FatalError("init (truncCylinder)", "origin must have size 3");
    for (int i = 0; i < 3; ++i) s->o[i] = o[i];
    s->r = d.getReal("radius");
    if (s->r <= 0.0) throw FatalError("init (truncCylinder)", "Radius must be +ve");
    s->a = d.getReal("halfwidth");
    if (s->a <= 0.0) throw FatalError("init (truncCylinder)", "Halfwidth must be +ve");
    s->axis = type[0] - 'x';
    s->p0 = (s->axis == 0) ? 1 : 0; s->p1 = (s->axis == 2) ? 1 : 2;
    return s;
  }
  throw FatalError("new_surface", "Unrecognised / unsupported type of a surface: " + type);
}

// Geometry/Surfaces/surfaceShelf_class.f90 : surfIdx = order of appearance
struct SurfaceShelf {
  std::vector<std::unique_ptr<Surface>> surfs;
  std::map<int, int> idMap;
  void init(const Dict& d) {
    for (auto& name : d.keys("dict")) {
      auto s = newSurface(d.getDict(name));
      if (idMap.count(s->id)) throw FatalError("surfaceShelf init", "Surfaces have the same ID");
      idMap[s->id] = (int)surfs.size() + 1;
      surfs.push_back(std::move(s));
    }
  }
  int getIdx(int id) const {
    auto it = idMap.find(id);
    if (it == idMap.end()) throw FatalError("surfaceShelf getIdx", "There is no surface with ID: " + std::to_string(id));
    return it->second;
  }
  Surface* getPtr(int idx) const { return surfs.at(idx - 1).get(); }
};

// ===========================================================================
// Cells  (Geometry/Cells/simpleCell_class.f90:90-141, cellShelf_class.f90)
// ===========================================================================
struct SimpleCell {
  int id = 0;
  std::vector<int> surfIdx;            // signed
  std::vector<Surface*> ptr;
  bool inside(const Vec3& r, const Vec3& u) const {
    bool isIt = false;
    for (size_t i = 0; i < ptr.size(); ++i) {
      bool sense = surfIdx[i] > 0;
      bool hs = ptr[i]->halfspace(r, u);
      isIt = (hs == sense);
      if (!isIt) return isIt;
    }
    return isIt;
  }
  void distance(double& d, int& sIdx, const Vec3& r, const Vec3& u) const {
    d = INF; sIdx = 0;
    for (size_t i = 0; i < ptr.size(); ++i) {
      double t = ptr[i]->distance(r, u);
      if (t < d) { d = t; sIdx = std::abs(surfIdx[i]); }
    }
  }
};

struct CellShelf {
  std::vector<SimpleCell> cells;
  std::vector<int> fill;
  std::map<int, int> idMap;
  void init(const Dict& d, const SurfaceShelf& surfs, const std::map<std::string, int>& mats) {
    for (auto& name : d.keys("dict")) {
      const Dict& cd = d.getDict(name);
      std::string type = cd.getWord("type");
      if (type != "simpleCell") throw FatalError("new_cell", "Unsupported cell type in oracle: " + type);
      SimpleCell c;
      c.id = cd.getInt("id");
      for (int sid : cd.getIntArray("surfaces")) {
        int idx = surfs.getIdx(std::abs(sid));
        c.ptr.push_back(surfs.getPtr(idx));
        c.surfIdx.push_back(sid < 0 ? -idx : idx);
      }
      if (idMap.count(c.id)) throw FatalError("cellShelf init", "Cells have the same ID");
      idMap[c.id] = (int)cells.size() + 1;
      std::string filling = cd.getWord("filltype");
      int f;
      if (filling == "outside") f = OUTSIDE_MAT;
      else if (filling == "mat") {
        auto it = mats.find(cd.getWord("material"));
        if (it == mats.end()) throw FatalError("cellShelf init", "Material was not found: " + cd.getWord("material"));
        f = it->second;
      } else if (filling == "uni") {
        f = cd.getInt("universe");
        if (f <= 0) throw FatalError("cellShelf init", "Universe ID must be +ve");
        f = -f;
      } else throw FatalError("cellShelf init", "Unknown type of cell filling: " + filling);
      cells.push_back(std::move(c));
      fill.push_back(f);
    }
  }
  int getIdx(int id) const {
    auto it = idMap.find(id);
    if (it == idMap.end()) throw FatalError("cellShelf getIdx", "There is no cell with ID: " + std::to_string(id));
    return it->second;
  }
};

// ===========================================================================
// coord / coordList  (Geometry/coord_class.f90)
// ===========================================================================
struct Coord {
  Vec3 r, dir;
  bool isRotated = false;
  double rotMat[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  int uniIdx = 0, uniRootID = 0, localID = 0, cellIdx = 0;
};

Vec3 rotateVector(const Vec3& dir, double mu, double phi);   // defined in physics.hpp (math mode)

struct CoordList {
  int nesting = 0;
  Coord lvl[MAX_NEST];
  int matIdx = UNDEF_MAT;
  int uniqueID = -3;
  void init(const Vec3& r, const Vec3& u) { takeAboveGeom(); lvl[0].r = r; lvl[0].dir = u; nesting = 1; }
  bool isPlaced() const { return matIdx > 0 && uniqueID > 0 && nesting >= 1; }
  void takeAboveGeom() { nesting = 1; matIdx = UNDEF_MAT; uniqueID = -3; }
  void moveGlobal(double d) {                                       // coord_class.f90:341-348
    takeAboveGeom();
    for (int k = 0; k < 3; ++k) lvl[0].r[k] = lvl[0].r[k] + d * lvl[0].dir[k];
  }
  void moveLocal(double d, int n) {                                 // coord_class.f90:362-375
    if (n > nesting || n < 1) throw FatalError("decreaseLevel", "New nesting is invalid");
    nesting = n;
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) lvl[i].r[k] = lvl[i].r[k] + d * lvl[i].dir[k];
  }
  void point(const Vec3& d) {                                       // assignDirection, coord_class.f90:453-472
    lvl[0].dir = d;
    for (int i = 1; i < nesting; ++i) {
      if (lvl[i].isRotated) lvl[i].dir = matvec(lvl[i].rotMat, lvl[i - 1].dir);
      else lvl[i].dir = lvl[i - 1].dir;
    }
  }
  void rotate(double mu, double phi) {                              // coord_class.f90:386-408
    lvl[0].dir = rotateVector(lvl[0].dir, mu, phi);
    for (int i = 1; i < nesting; ++i) {
      if (lvl[i].isRotated) lvl[i].dir = matvec(lvl[i].rotMat, lvl[i - 1].dir);
      else lvl[i].dir = lvl[i - 1].dir;
    }
  }
};

// ===========================================================================
// Universes  (Geometry/Universes/*.f90)
// ===========================================================================
inline int charToFill(const std::string& name, const std::map<std::string, int>& mats, const char* where) {
  // universe_inter.f90: 'u<ID>' is a universe, anything else is a material name
  if (name.size() > 2 && name[0] == 'u' && name[1] == '<') {
    size_t pos = name.rfind('>');
    if (pos != std::string::npos) {
      std::string num = name.substr(2, pos - 2);
      if (!Dict::isInt(num)) throw FatalError(where, "Failed to convert " + name + " to universe ID");
      int f = Dict::toInt(num);
      if (f <= 0) throw FatalError(where, "Universe ID must be +ve");
      return -f;
    }
  }
  auto it = mats.find(name);
  if (it == mats.end()) throw FatalError(where, "Unknown material: " + name);
  return it->second;
}

struct Universe {
  int uniId = 0, uniIdx = 0;
  double origin[3] = {0, 0, 0};
  double rotMat[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  bool rot = false, globalTrans = false;
  virtual ~Universe() = default;
  virtual std::string myType() const = 0;
  virtual void findCell(int& localID, int& cellIdx, const Vec3& r, const Vec3& u) = 0;
  virtual void distance(double& d, int& surfIdx, const Coord& c) = 0;
  virtual void cross(Coord& c, int surfIdx) = 0;
  virtual Vec3 cellOffset(const Coord& c) const { (void)c; return Vec3(); }

  void setupBase(const Dict& d) {                                   // universe_inter.f90 setupBase
    int id = d.getInt("id");
    if (id <= 0) throw FatalError("setupBase", "Universe ID must be +ve");
    uniId = id;
    if (d.isPresent("origin")) {
      auto t = d.getRealArray("origin");
      if (t.size() != 3) throw FatalError("setupBase", "Origin must have size 3");
      for (int i = 0; i < 3; ++i) origin[i] = t[i];
    }
    if (d.isPresent("rotation")) {
      auto t = d.getRealArray("rotation");
      if (t.size() != 3) throw FatalError("setupBase", "3 rotation angles must be given");
      if (!(t[0] == 0.0 && t[1] == 0.0 && t[2] == 0.0)) { rot = true; rotationMatrix(rotMat, t[0], t[1], t[2]); }
    }
    if (d.isPresent("global")) globalTrans = d.getBool("global");
  }
  // universe_inter.f90:400-424
  void enter(Coord& n, const Vec3& r, const Vec3& u) {
    n = Coord();
    n.r = r; n.dir = u; n.uniIdx = uniIdx; n.isRotated = rot;
    if (rot) {
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) n.rotMat[i][j] = rotMat[i][j];
      n.r = matvec(rotMat, n.r);
      n.dir = matvec(rotMat, n.dir);
    }
    for (int k = 0; k < 3; ++k) n.r[k] = n.r[k] - origin[k];
    findCell(n.localID, n.cellIdx, n.r, n.dir);
  }
};

// rootUniverse_class.f90:127-181
struct RootUniverse : Universe {
  Surface* surf = nullptr; int surfIdx = 0;
  std::string myType() const override { return "rootUniverse"; }
  void findCell(int& localID, int& cellIdx, const Vec3& r, const Vec3& u) override {
    cellIdx = 0;
    localID = surf->halfspace(r, u) ? 2 : 1;
  }
  void distance(double& d, int& sIdx, const Coord& c) override { sIdx = surfIdx; d = surf->distance(c.r, c.dir); }
  void cross(Coord& c, int) override { findCell(c.localID, c.cellIdx, c.r, c.dir); }
};

// pinUniverse_class.f90:72-253
struct PinUniverse : Universe {
  static constexpr int MOVING_IN = -1, MOVING_OUT = -2;
  std::vector<double> r_sq;
  std::vector<Cylinder> annuli;
  std::string myType() const override { return "pinUniverse"; }
  void findCell(int& localID, int& cellIdx, const Vec3& r, const Vec3& u) override {
    double rs = r[0] * r[0] + r[1] * r[1];
    cellIdx = 0;
    double mul = (r[0] * u[0] + r[1] * u[1] >= 0.0) ? -1.0 : 1.0;
    int N = (int)r_sq.size();
    for (localID = 1; localID <= N; ++localID)
      if (rs < r_sq[localID - 1] + mul * annuli[localID - 1].tol) return;
    // falls through with localID = N + 1
  }
  void distance(double& d, int& sIdx, const Coord& c) override {
    int id = c.localID, N = (int)r_sq.size();
    if (id < 1 || id > N + 1) throw FatalError("distance (pinUniverse)", "Invalid local ID");
    double d_out = (id > N) ? INF : annuli[id - 1].distance(c.r, c.dir);
    double d_in = (id == 1) ? INF : annuli[id - 2].distance(c.r, c.dir);
    if (d_in < d_out) { sIdx = MOVING_IN; d = d_in; } else { sIdx = MOVING_OUT; d = d_out; }
  }
  void cross(Coord& c, int sIdx) override {
    if (sIdx == MOVING_IN) c.localID -= 1;
    else if (sIdx == MOVING_OUT) c.localID += 1;
    else throw FatalError("cross (pinUniverse)", "Unknown surface memento");
  }
};

// latUniverse_class.f90:106-414,489-506
struct LatUniverse : Universe {
  static constexpr int OUTLINE_SURF = -7;
  double pitch[3] = {0, 0, 0}, corner[3] = {0, 0, 0}, a_bar[3] = {0, 0, 0};
  int sizeN[3] = {0, 0, 0};
  BoxLike outline;
  int outLocalID = 0;
  bool offset = true;
  std::vector<int> offsetMap;
  std::string myType() const override { return "latUniverse"; }
  static void get_ijk(int ijk[3], int localID, const int sizeN[3]) {
    int temp = localID - 1;
    int base = temp / sizeN[0];
    ijk[0] = temp - sizeN[0] * base + 1;
    temp = base;
    base = temp / sizeN[1];
    ijk[1] = temp - sizeN[1] * base + 1;
    ijk[2] = base + 1;
  }
  void findCell(int& localID, int& cellIdx, const Vec3& r, const Vec3& u) override {
    int ijk[3]; double r_bar[3];
    for (int i = 0; i < 3; ++i) {
      ijk[i] = (int)std::floor((r[i] - corner[i]) / pitch[i]) + 1;
      r_bar[i] = r[i] - corner[i] - ijk[i] * pitch[i] + 0.5 * pitch[i];
    }
    for (int i = 0; i < 3; ++i) {
      if (std::fabs(r_bar[i]) > a_bar[i] && r_bar[i] * u[i] > 0.0) {
        int inc = (u[i] < 0.0) ? -1 : 1;
        ijk[i] += inc;
      }
    }
    bool out = false;
    for (int i = 0; i < 3; ++i) if (ijk[i] <= 0 || ijk[i] > sizeN[i]) out = true;
    if (out) localID = outLocalID;
    else localID = ijk[0] + sizeN[0] * (ijk[1] - 1 + sizeN[1] * (ijk[2] - 1));
    cellIdx = 0;
  }
  void distance(double& d, int& sIdx, const Coord& c) override {
    if (c.localID == outLocalID) { sIdx = OUTLINE_SURF; d = outline.distance(c.r, c.dir); return; }
    int ijk[3]; get_ijk(ijk, c.localID, sizeN);
    double r_bar[3], bounds[3];
    for (int i = 0; i < 3; ++i) {
      r_bar[i] = c.r[i] - corner[i];
      r_bar[i] = r_bar[i] - (ijk[i] - 0.5) * pitch[i];
      bounds[i] = fsign(pitch[i] * 0.5, c.dir[i]);
    }
    d = INF; int ax = 1;
    for (int i = 0; i < 3; ++i) {
      double test_d = (bounds[i] - r_bar[i]) / c.dir[i];
      if (test_d < d) { d = test_d; ax = i + 1; }
    }
    d = std::max(0.0, d);
    d = std::min(INF, d);
    sIdx = ax * 2;
    if (c.dir[ax - 1] < 0.0) sIdx -= 1;
    sIdx = -sIdx;
  }
  void cross(Coord& c, int) override { findCell(c.localID, c.cellIdx, c.r, c.dir); }
  Vec3 cellOffset(const Coord& c) const override {
    bool doOffset = offsetMap.empty() ? offset : (offsetMap[c.localID - 1] == 1);
    Vec3 o;
    if (doOffset && c.localID != outLocalID) {
      int ijk[3]; get_ijk(ijk, c.localID, sizeN);
      for (int i = 0; i < 3; ++i) o[i] = (ijk[i] - 0.5) * pitch[i] + corner[i];
    }
    return o;
  }
};

// cellUniverse_class.f90:203-437 ; the visit-count reordering is a search-order
// optimisation only (cells do not overlap), so cells are searched in input order.
struct CellUniverse : Universe {
  std::vector<int> cellIdxs;           // index into CellShelf (1-based)
  const CellShelf* shelf = nullptr;
  bool checkOverlap = false;
  std::string myType() const override { return "cellUniverse"; }
  void findCell(int& localID, int& cellIdx, const Vec3& r, const Vec3& u) override {
    int N = (int)cellIdxs.size();
    if (checkOverlap) {
      int found = 0, foundID = 0;
      for (int i = 1; i <= N; ++i)
        if (shelf->cells[cellIdxs[i - 1] - 1].inside(r, u)) { foundID = i; cellIdx = cellIdxs[i - 1]; ++found; }
      if (found == 0) { localID = N + 1; cellIdx = 0; }
      else if (found > 1) { localID = N + 2; cellIdx = 0; }
      else localID = foundID;
      return;
    }
    for (int i = 1; i <= N; ++i) {
      if (shelf->cells[cellIdxs[i - 1] - 1].inside(r, u)) { localID = i; cellIdx = cellIdxs[i - 1]; return; }
    }
    localID = N + 1; cellIdx = 0;
  }
  void distance(double& d, int& sIdx, const Coord& c) override {
    int N = (int)cellIdxs.size();
    if (c.localID == N + 1) throw FatalError("distance (cellUniverse)", "Particle is in undefined local cell");
    if (c.localID == N + 2) throw FatalError("distance (cellUniverse)", "Particle is in an overlapping local cell");
    shelf->cells[cellIdxs[c.localID - 1] - 1].distance(d, sIdx, c.r, c.dir);
  }
  void cross(Coord& c, int) override {
    for (int k = 0; k < 3; ++k) c.r[k] = c.r[k] + c.dir[k] * NUDGE;
    findCell(c.localID, c.cellIdx, c.r, c.dir);
  }
};

inline std::unique_ptr<Universe> newUniverse(std::vector<int>& fill, const Dict& d, const CellShelf& cells,
                                             const SurfaceShelf& surfs, const std::map<std::string, int>& mats) {
  std::string type = d.getWord("type");
  if (type == "rootUniverse") {
    auto u = std::make_unique<RootUniverse>();
    u->setupBase(d);
    if (d.isPresent("origin")) throw FatalError("init (rootUniverse)", "Origin is not allowed");
    if (d.isPresent("rotation")) throw FatalError("init (rootUniverse)", "Rotation is not allowed");
    int id = d.getInt("border");
    if (id <= 0) throw FatalError("init (rootUniverse)", "Border must be given as +ve ID");
    u->surfIdx = surfs.getIdx(id);
    u->surf = surfs.getPtr(u->surfIdx);
    fill.assign(2, 0);
    fill[1] = OUTSIDE_MAT;
    fill[0] = charToFill(d.getWord("fill"), mats, "init (rootUniverse)");
    return u;
  }
  if (type == "pinUniverse") {
    auto u = std::make_unique<PinUniverse>();
    u->setupBase(d);
    auto radii = d.getRealArray("radii");
    auto names = d.getWordArray("fills");
    if (radii.size() != names.size()) throw FatalError("init (pinUniverse)", "Size of radii and fills does not match");
    for (double r : radii) if (r < 0.0) throw FatalError("init (pinUniverse)", "Found -ve value of radius.");
    int N = (int)radii.size();
    int idx = (int)(std::min_element(radii.begin(), radii.end()) - radii.begin());
    if (radii[idx] != 0.0) throw FatalError("init (pinUniverse)", "Did not found outermost element with radius 0.0.");
    std::swap(radii[idx], radii[N - 1]); std::swap(names[idx], names[N - 1]);
    radii[N - 1] = INF * 1.1;
    for (int i = N - 2; i >= 0; --i) {       // selection sort as written (maxloc of 1..i+1)
      int m = (int)(std::max_element(radii.begin(), radii.begin() + i + 1) - radii.begin());
      std::swap(radii[m], radii[i]); std::swap(names[m], names[i]);
    }
    for (int i = 0; i + 1 < N; ++i) if (radii[i] == radii[i + 1]) throw FatalError("init (pinUniverse)", "Duplicate value of radius");
    u->r_sq.resize(N); u->annuli.resize(N);
    const double o[3] = {0, 0, 0};
    for (int i = 0; i < N; ++i) { u->r_sq[i] = radii[i] * radii[i]; u->annuli[i].build(1, "zCylinder", o, radii[i]); }
    fill.resize(N);
    for (int i = 0; i < N; ++i) fill[i] = charToFill(names[i], mats, "init (pinUniverse)");
    return u;
  }
  if (type == "latUniverse") {
    auto u = std::make_unique<LatUniverse>();
    u->setupBase(d);
    u->offset = d.getBool("offset", true);
    auto p = d.getRealArray("pitch");
    if (p.size() != 3) throw FatalError("init (latUniverse)", "Pitch must have size 3");
    auto s = d.getIntArray("shape");
    if (s.size() != 3) throw FatalError("init (latUniverse)", "Shape must have size 3");
    for (int i = 0; i < 3; ++i) { if (s[i] < 0) throw FatalError("init (latUniverse)", "Shape contains -ve entries"); u->pitch[i] = p[i]; u->sizeN[i] = s[i]; }
    if (u->sizeN[2] == 0) { u->sizeN[2] = 1; u->pitch[2] = 2.0 * INF; }
    for (int i = 0; i < 3; ++i) if (u->sizeN[i] == 0) throw FatalError("init (latUniverse)", "Shape in X and Y axis cannot be 0.");
    for (int i = 0; i < 3; ++i) if (u->pitch[i] < 10 * SURF_TOL) throw FatalError("init (latUniverse)", "Pitch size too small");
    for (int i = 0; i < 3; ++i) {
      u->a_bar[i] = u->pitch[i] * 0.5 - u->pitch[i] * SURF_TOL;
      u->corner[i] = -(u->sizeN[i] * 0.5 * u->pitch[i]);
    }
    int nCells = u->sizeN[0] * u->sizeN[1] * u->sizeN[2];
    u->outLocalID = nCells + 1;
    u->outline.nax = 3; u->outline.id = 1;
    for (int i = 0; i < 3; ++i) { u->outline.o[i] = 0.0; u->outline.hw[i] = std::fabs(u->corner[i]); }
    auto m = d.getIntArray("map");
    if ((int)m.size() != nCells) throw FatalError("init (latUniverse)", "Lattice map size not equal to size implied by shape");
    // flip up-down: rows (fastest index x) are reversed over the combined y*z index
    int nx = u->sizeN[0], ncol = u->sizeN[1] * u->sizeN[2];
    auto flip = [&](std::vector<int>& a) {
      for (int j = 0; j < ncol / 2; ++j)
        for (int i = 0; i < nx; ++i) std::swap(a[i + j * nx], a[i + (ncol - 1 - j) * nx]);
    };
    flip(m);
    int outFill = charToFill(d.getWord("padMat"), mats, "init (latUniverse)");
    fill.resize(nCells + 1);
    for (int i = 0; i < nCells; ++i) fill[i] = -m[i];
    fill[nCells] = outFill;
    if (d.isPresent("offsetMap")) {
      if (!u->offset) throw FatalError("init (latUniverse)", "Cannot have both an offset map and no offset.");
      auto om = d.getIntArray("offsetMap");
      if ((int)om.size() != nCells) throw FatalError("init (latUniverse)", "Offset map size mismatch");
      flip(om);
      for (int v : om) if (v != 0 && v != 1) throw FatalError("init (latUniverse)", "Invalid entry to the offset map");
      om.push_back(0);
      u->offsetMap = om;
    }
    return u;
  }
  if (type == "cellUniverse") {
    auto u = std::make_unique<CellUniverse>();
    u->setupBase(d);
    u->shelf = &cells;
    for (int cid : d.getIntArray("cells")) u->cellIdxs.push_back(cells.getIdx(cid));
    int N = (int)u->cellIdxs.size();
    fill.resize(N + 2);
    for (int i = 0; i < N; ++i) fill[i] = cells.fill[u->cellIdxs[i] - 1];
    fill[N] = UNDEF_MAT; fill[N + 1] = OVERLAP_MAT;
    u->checkOverlap = d.getBool("checkOverlap", false);
    return u;
  }
  throw FatalError("new_universe", "Unrecognised / unsupported type of universe: " + type);
}

// ===========================================================================
// csg + geomGraph  (Geometry/csg_class.f90:76-207, geomGraph_class.f90:136-345,
//                   Universes/uniFills_class.f90, universeShelf_class.f90:102-138)
// ===========================================================================
struct Location { int idx = 0; int id = 0; };

struct CSG {
  SurfaceShelf surfs;
  CellShelf cells;
  std::vector<std::unique_ptr<Universe>> unis;
  std::vector<std::string> uniNames;
  std::vector<std::vector<int>> fills;     // per universe idx (0-based), entries: mat (>=0) or -uniIdx
  std::map<int, int> uniIdMap;
  int rootIdx = 0, borderIdx = 0;
  std::vector<Location> graph;
  int uniqueCells = 0;
  std::vector<int> usedMats;
  int nesting = 0;

  void init(const Dict& d, const std::map<std::string, int>& mats) {
    surfs.init(d.getDict("surfaces"));
    cells.init(d.getDict("cells"), surfs, mats);
    const Dict& ud = d.getDict("universes");
    for (auto& name : ud.keys("dict")) {
      std::vector<int> f;
      auto u = newUniverse(f, ud.getDict(name), cells, surfs, mats);
      if (uniIdMap.count(u->uniId)) throw FatalError("universeShelf init", "Universes have the same ID: " + std::to_string(u->uniId));
      u->uniIdx = (int)unis.size() + 1;
      uniIdMap[u->uniId] = u->uniIdx;
      unis.push_back(std::move(u));
      uniNames.push_back(name);
      fills.push_back(std::move(f));
    }
    if (unis.empty()) throw FatalError("uniFills init", "Given not +ve number of universes");
    // translate universe IDs to indices (uniFills finishBuild)
    for (auto& f : fills)
      for (auto& x : f)
        if (x < 0) {
          auto it = uniIdMap.find(-x);
          if (it == uniIdMap.end()) throw FatalError("uniFills finishBuild", "There is no universe with ID: " + std::to_string(-x));
          x = -it->second;
        }
    int rootId;
    if (d.isPresent("root")) rootId = d.getInt("root");
    else rootId = ud.getDict("root").getInt("id");
    auto it = uniIdMap.find(rootId);
    if (it == uniIdMap.end()) throw FatalError("csg init", "There is no universe with ID: " + std::to_string(rootId));
    rootIdx = it->second;
    auto* root = dynamic_cast<RootUniverse*>(unis[rootIdx - 1].get());
    if (!root) throw FatalError("csg init", "Root universe is not type `rootUniverse`");
    borderIdx = root->surfIdx;
    surfs.getPtr(borderIdx)->setBC(d.getIntArray("boundary"));
    // structure checks
    std::vector<int> stack;
    if (hasCycles(rootIdx, stack)) throw FatalError("csg init", "There is recursion in the geometry nesting.");
    nesting = countDepth(rootIdx);
    if (nesting > MAX_NEST) throw FatalError("csg init", "Nesting level > max nesting");
    for (int x : fills[rootIdx - 1]) if (x < 0 && outsideBelow(-x)) throw FatalError("csg init", "Cell with outside fill is present below root universe");
    std::string gtype = d.getDict("graph").getWord("type");
    if (gtype == "shrunk") buildShrunk();
    else if (gtype == "extended") buildExtended();
    else throw FatalError("geomGraph init", "Unknown geometry graph type: " + gtype);
  }

  bool hasCycles(int idx, std::vector<int>& stack) const {
    for (int s : stack) if (s == idx) return true;
    stack.push_back(idx);
    for (int x : fills[idx - 1]) if (x < 0 && hasCycles(-x, stack)) return true;
    stack.pop_back();
    return false;
  }
  int countDepth(int idx) const {
    int dep = 0;
    for (int x : fills[idx - 1]) if (x < 0) dep = std::max(dep, countDepth(-x));
    return dep + 1;
  }
  bool outsideBelow(int idx) const {
    for (int x : fills[idx - 1]) {
      if (x == OUTSIDE_MAT) return true;
      if (x < 0 && outsideBelow(-x)) return true;
    }
    return false;
  }
  void collectUsed(int idx, std::vector<char>& used) const {
    used[idx - 1] = 1;
    for (int x : fills[idx - 1]) if (x < 0) collectUsed(-x, used);
  }
  long countInstancesBelow(int idx, std::vector<long>& cnt) const {
    cnt[idx - 1] += 1;
    for (int x : fills[idx - 1]) if (x < 0) countInstancesBelow(-x, cnt);
    return 0;
  }
  void layout(int& top, int idx) {
    const auto& f = fills[idx - 1];
    if (top - 1 + f.size() > graph.size()) throw FatalError("layoutUniverse", "Overflow of the location array");
    for (size_t i = 0; i < f.size(); ++i) graph[top - 1 + i].idx = f[i];
    top += (int)f.size();
  }
  void buildShrunk() {                                               // geomGraph_class.f90:172-236
    std::vector<char> used(unis.size(), 0);
    collectUsed(rootIdx, used);
    size_t N = 0;
    for (size_t i = 0; i < unis.size(); ++i) if (used[i]) N += fills[i].size();
    graph.assign(N, Location());
    std::map<int, int> layed;
    int top = 1; size_t loc = 1;
    layed[rootIdx] = top;
    layout(top, rootIdx);
    while (loc <= graph.size()) {
      int fill = graph[loc - 1].idx;
      if (fill < 0) {
        int rootID;
        auto it = layed.find(-fill);
        if (it == layed.end()) { rootID = top; layed[-fill] = top; layout(top, -fill); }
        else rootID = it->second;
        graph[loc - 1].id = rootID;
      }
      ++loc;
    }
    if (top != (int)N + 1) throw FatalError("buildShrunk", "Did not reach the end of the location array");
    setUniqueIDs();
  }
  void buildExtended() {                                             // geomGraph_class.f90:248-296
    std::vector<long> cnt(unis.size(), 0);
    countInstancesBelow(rootIdx, cnt);
    size_t N = 0;
    for (size_t i = 0; i < unis.size(); ++i) N += fills[i].size() * (size_t)cnt[i];
    graph.assign(N, Location());
    int top = 1; size_t loc = 1;
    layout(top, rootIdx);
    while (loc <= graph.size()) {
      int fill = graph[loc - 1].idx;
      if (fill < 0) { int rootID = top; layout(top, -fill); graph[loc - 1].id = rootID; }
      ++loc;
    }
    if (top != (int)N + 1) throw FatalError("buildExtended", "Did not reach the end of the location array");
    setUniqueIDs();
  }
  void setUniqueIDs() {                                              // geomGraph_class.f90:305-345
    int c = 0;
    std::vector<int> mats;
    for (auto& l : graph)
      if (l.idx > 0) { ++c; l.id = c; if (std::find(mats.begin(), mats.end(), l.idx) == mats.end()) mats.push_back(l.idx); }
    uniqueCells = c;
    std::sort(mats.begin(), mats.end());
    usedMats = mats;
  }
  void getFill(int& idx, int& id, int uniRootID, int localID) const {
    const Location& l = graph.at(uniRootID + localID - 2);
    idx = l.idx; id = l.id;
  }
};

// ===========================================================================
// geometryStd  (Geometry/geometryStd_class.f90)
// ===========================================================================
struct DistCache { int lvl = 0; double dist[MAX_NEST]; int surf[MAX_NEST]; };

struct GeometryStd {
  CSG geom;
  void init(const Dict& d, const std::map<std::string, int>& mats) { geom.init(d, mats); }

  void placeCoord(CoordList& c) const {                              // :119-147
    if (c.nesting < 1) throw FatalError("placeCoord", "CoordList is not initialised");
    c.takeAboveGeom();
    Vec3 r = c.lvl[0].r, dir = c.lvl[0].dir;
    geom.unis[geom.rootIdx - 1]->enter(c.lvl[0], r, dir);
    c.lvl[0].uniRootID = 1;
    diveToMat(c, 1);
  }
  void whatIsAt(int& matIdx, int& uniqueID, const Vec3& r, const Vec3* u = nullptr) const {   // :154-181
    Vec3 ul; ul[0] = 1.0;
    if (u) ul = *u;
    CoordList c; c.init(r, ul);
    placeCoord(c);
    matIdx = c.matIdx; uniqueID = c.uniqueID;
  }
  void bounds(double b[6]) const {                                   // :188-206
    geom.surfs.getPtr(geom.borderIdx)->boundingBox(b);
    for (int i = 0; i < 3; ++i) if (b[i] <= -INF && b[i + 3] >= INF) { b[i] = 0.0; b[i + 3] = 0.0; }
  }
  void diveToMat(CoordList& c, int start) const {                    // :565-619
    for (int i = start; i <= MAX_NEST; ++i) {
      int rootID = c.lvl[i - 1].uniRootID, localID = c.lvl[i - 1].localID, fill, id;
      geom.getFill(fill, id, rootID, localID);
      if (fill >= 0) { c.matIdx = fill; c.uniqueID = id; return; }
      if (i == MAX_NEST) break;
      fill = -fill;
      Universe* uni = geom.unis[c.lvl[i - 1].uniIdx - 1].get();
      Vec3 off = uni->cellOffset(c.lvl[i - 1]);
      uni = geom.unis[fill - 1].get();
      Vec3 r;
      if (uni->globalTrans) r = c.lvl[0].r;
      else for (int k = 0; k < 3; ++k) r[k] = c.lvl[i - 1].r[k] - off[k];
      c.nesting += 1;
      Vec3 dir = c.lvl[i - 1].dir;
      uni->enter(c.lvl[i], r, dir);
      c.lvl[i].uniRootID = id;
    }
    throw FatalError("diveToMat", "Failed to find material cell");
  }
  void closestDist(double& dist, int& surfIdx, int& lvl, const CoordList& c) const {   // :633-660
    dist = INF; surfIdx = 0; lvl = 0;
    for (int l = 1; l <= c.nesting; ++l) {
      double td; int ti;
      geom.unis[c.lvl[l - 1].uniIdx - 1]->distance(td, ti, c.lvl[l - 1]);
      if ((dist - td) >= dist * FP_REL_TOL) { dist = td; surfIdx = ti; lvl = l; }
    }
  }
  void closestDist_cache(double& dist, int& surfIdx, int& lvl, const CoordList& c, DistCache& cache) const {  // :678-717
    dist = INF; surfIdx = 0; lvl = 0;
    for (int l = 1; l <= c.nesting; ++l) {
      if (cache.lvl < l) {
        geom.unis[c.lvl[l - 1].uniIdx - 1]->distance(cache.dist[l - 1], cache.surf[l - 1], c.lvl[l - 1]);
        cache.lvl += 1;
      }
      double td = cache.dist[l - 1]; int ti = cache.surf[l - 1];
      if ((dist - td) >= dist * FP_REL_TOL) { dist = td; surfIdx = ti; lvl = l; }
    }
  }
  // move_noCache :214-277 / move_withCache :288-352 (no fields: fieldDist = INF)
  void move(CoordList& c, double& maxDist, int& event, DistCache* cache = nullptr) const {
    if (!c.isPlaced()) throw FatalError("move", "Coordinate list is not placed in the geometry");
    double dist; int surfIdx, level;
    if (cache) closestDist_cache(dist, surfIdx, level, c, *cache);
    else closestDist(dist, surfIdx, level, c);
    if (maxDist < dist) {
      c.moveLocal(maxDist, c.nesting);
      event = COLL_EV;
      if (cache) cache->lvl = 0;
    } else if (surfIdx == geom.borderIdx && level == 1) {
      c.moveGlobal(dist);
      event = BOUNDARY_EV;
      maxDist = dist;
      if (cache) cache->lvl = 0;
      geom.surfs.getPtr(geom.borderIdx)->explicitBC(c.lvl[0].r, c.lvl[0].dir);
      placeCoord(c);
    } else {
      c.moveLocal(dist, level);
      event = CROSS_EV;
      maxDist = dist;
      if (cache) {
        for (int l = 0; l < level - 1; ++l) cache->dist[l] = cache->dist[l] - dist;
        cache->lvl = level - 1;
      }
      geom.unis[c.lvl[level - 1].uniIdx - 1]->cross(c.lvl[level - 1], surfIdx);
      diveToMat(c, level);
    }
  }
  void moveGlobal(CoordList& c, double& maxDist, int& event) const {   // :363-394
    Surface* surf = geom.surfs.getPtr(geom.borderIdx);
    double dist = surf->distance(c.lvl[0].r, c.lvl[0].dir);
    if (maxDist < dist) { c.moveGlobal(maxDist); event = COLL_EV; }
    else { c.moveGlobal(dist); event = BOUNDARY_EV; surf->explicitBC(c.lvl[0].r, c.lvl[0].dir); maxDist = dist; }
    placeCoord(c);
  }
  void teleport(CoordList& c, double dist) const {                   // :492-514
    c.moveGlobal(dist);
    placeCoord(c);
    if (c.matIdx == OUTSIDE_MAT) {
      geom.surfs.getPtr(geom.borderIdx)->transformBC(c.lvl[0].r, c.lvl[0].dir);
      placeCoord(c);
    }
  }
  std::vector<int> activeMats() const {                              // :521-547
    const auto& um = geom.usedMats;
    int N = (int)um.size();
    if (N == 0) return {};
    int last = um[N - 1];
    if (last == VOID_MAT) { N -= 1; if (N == 0) return {}; last = um[N - 1]; }
    if (last == UNDEF_MAT) { N -= 1; if (N == 0) return {}; last = um[N - 1]; }
    if (last == OVERLAP_MAT) N -= 1;
    return std::vector<int>(um.begin(), um.begin() + N);
  }
};

}  // namespace orc
