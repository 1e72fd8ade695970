// SCONE's 63-bit LCG with O(log k) skip-ahead, usable from host and device code.
//   RandomNumbers/RNG_class.f90:40-44    g = 2806196910506780709, c = 1, M = 2^63, stride 152917
//   RandomNumbers/RNG_class.f90:147-168  get    (state * 2^-63, integer->real rounds to nearest)
//   RandomNumbers/RNG_class.f90:251-299  skip   (F. Brown's arbitrary-stride algorithm)
#pragma once
#include <stdint.h>

#include "sb_math.h"   // SB_HD

namespace sbd {

constexpr uint64_t RNG_G = 2806196910506780709ULL;
constexpr uint64_t RNG_MASK = 0x7fffffffffffffffULL;
constexpr int64_t RNG_STRIDE = 152917;

SB_HD double rng_get(uint64_t& s) {
  s = (RNG_G * s) & RNG_MASK;
  s = (s + 1ULL) & RNG_MASK;
  return (double)(int64_t)s * (1.0 / 9223372036854775808.0);
}
SB_HD uint64_t rng_skip(uint64_t s, int64_t k_in) {
  uint64_t k = (k_in >= 0) ? (uint64_t)k_in : (uint64_t)(INT64_MAX - (-k_in) + 1);
  k &= RNG_MASK;
  uint64_t Gk = 1, Ck = 0, h = RNG_G, L = 1;
  while (k > 0) {
    if (k & 1ULL) { Gk = (Gk * h) & RNG_MASK; Ck = (Ck * h) & RNG_MASK; Ck = (Ck + L) & RNG_MASK; }
    L = (L * (h + 1)) & RNG_MASK;
    h = (h * h) & RNG_MASK;
    k >>= 1;
  }
  return (Gk * s + Ck) & RNG_MASK;
}

}  // namespace sbd
