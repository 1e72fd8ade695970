// ORACLE (test infrastructure, not product code).
// CPU restatement of SCONE's 63-bit LCG.
//   RandomNumbers/RNG_class.f90:40-44   constants g, c, M = 2^63, norm = 2^-63
//   RandomNumbers/RNG_class.f90:147-168 get
//   RandomNumbers/RNG_class.f90:251-299 skip  (F. Brown's arbitrary-stride algorithm)
//   RandomNumbers/RNG_class.f90:305-316 stride (skip by 152917 * n)
// Pinned by tests/test_oracle_rng.py against RandomNumbers/Tests/RNG_test.f90:15-70.
#pragma once
#include <cstdint>

namespace orc {

struct RNG {
  static constexpr uint64_t G = 2806196910506780709ULL;
  static constexpr uint64_t C = 1ULL;
  static constexpr uint64_t MASK = 0x7fffffffffffffffULL;   // huge(0_int64)
  static constexpr int64_t STRIDE = 152917;

  uint64_t seed = 0;
  uint64_t initialSeed = 0;
  uint64_t count = 0;

  void init(int64_t s) { seed = (uint64_t)s; initialSeed = (uint64_t)s; }

  uint64_t getInt() {
    uint64_t s = (G * seed) & MASK;
    s = (s + C) & MASK;
    seed = s;
    ++count;
    return s;
  }
  // rand = seed * 2^-63 ; integer -> real conversion rounds to nearest
  double get() { return (double)(int64_t)getInt() * (1.0 / 9223372036854775808.0); }

  void skip(int64_t k_in) {
    uint64_t k;
    // -ve skip == skip by (period - |k|); period is 2^63
    if (k_in >= 0) k = (uint64_t)k_in;
    else k = (uint64_t)(INT64_MAX - (-k_in) + 1);
    k &= MASK;
    uint64_t Gk = 1, Ck = 0, h = G, L = C;
    while (k > 0) {
      if (k & 1ULL) {
        Gk = (Gk * h) & MASK;
        Ck = (Ck * h) & MASK;
        Ck = (Ck + L) & MASK;
      }
      L = (L * (h + 1)) & MASK;
      h = (h * h) & MASK;     // == tabulated pow_of_gsq(i), RNG_class.f90:61-123
      k >>= 1;
    }
    seed = (Gk * seed + Ck) & MASK;
  }
  void stride(int32_t n) { skip(STRIDE * (int64_t)n); }
  uint64_t currentState() const { return seed; }
};

}  // namespace orc
