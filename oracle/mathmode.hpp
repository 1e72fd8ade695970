// ORACLE (test infrastructure, not product code).
// Math mode of the oracle. The reference calls the Fortran intrinsics (glibc libm on the CPU): MATH_LIBM does the
// same. MATH_SB routes log / sin / cos through scone_b200/csrc/sb_math.h, the deterministic implementation the CUDA
// engine uses, so that oracle and engine follow bit-identical histories.
#pragma once
#include <cmath>

#include "../scone_b200/csrc/sb_math.h"

namespace orc {

enum MathMode { MATH_LIBM = 0, MATH_SB = 1 };
inline int& mathMode() { static int m = MATH_LIBM; return m; }
inline double mlog(double x) { return mathMode() == MATH_SB ? sbm::log(x) : std::log(x); }
inline void msincos(double x, double& s, double& c) {
  if (mathMode() == MATH_SB) sbm::sincos(x, &s, &c);
  else { s = std::sin(x); c = std::cos(x); }
}
inline double mcos(double x) { double s, c; msincos(x, s, c); return c; }

}  // namespace orc
