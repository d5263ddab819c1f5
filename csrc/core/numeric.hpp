// Host-side numeric helpers shared by the host backend and the tests: bf16/f16/e4m3 conversions that are
// bit-compatible with what the CUDA kernels produce (round-to-nearest-even, e4m3 saturating at +-448).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace mlslb {

inline float bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (uint16_t)(u >> 16);
}

inline float f16_to_f32(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1fu;
  uint32_t man = h & 0x3ffu;
  uint32_t u;
  if (exp == 0) {
    if (man == 0) {
      u = sign;
    } else {
      int e = -1;
      do {
        e++;
        man <<= 1;
      } while ((man & 0x400u) == 0);
      man &= 0x3ffu;
      u = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
    }
  } else if (exp == 31) {
    u = sign | 0x7f800000u | (man << 13);
  } else {
    u = sign | ((exp + 112u) << 23) | (man << 13);
  }
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline uint16_t f32_to_f16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  uint32_t sign = (u >> 16) & 0x8000u;
  int32_t exp = (int32_t)((u >> 23) & 0xffu) - 127 + 15;
  uint32_t man = u & 0x7fffffu;
  if (((u >> 23) & 0xffu) == 0xffu) return (uint16_t)(sign | 0x7c00u | (man ? 0x200u : 0));
  if (exp >= 31) return (uint16_t)(sign | 0x7c00u);
  if (exp <= 0) {
    if (exp < -10) return (uint16_t)sign;
    man |= 0x800000u;
    uint32_t shift = (uint32_t)(14 - exp);
    uint32_t half = man >> shift;
    uint32_t rem = man & ((1u << shift) - 1u);
    uint32_t mid = 1u << (shift - 1);
    if (rem > mid || (rem == mid && (half & 1u))) half++;
    return (uint16_t)(sign | half);
  }
  uint32_t half = ((uint32_t)exp << 10) | (man >> 13);
  uint32_t rem = man & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (half & 1u))) half++;
  return (uint16_t)(sign | half);
}

// OCP FP8 E4M3 (finite-only variant "e4m3fn": max 448, no infinities, NaN = 0x7f/0xff), satfinite conversion.
// definition of the format (slow: libm calls); the hot paths use the table / bit-twiddling versions below
inline float e4m3_to_f32_ref(uint8_t v) {
  uint32_t sign = v >> 7;
  uint32_t exp = (v >> 3) & 0xfu;
  uint32_t man = v & 0x7u;
  float r;
  if (exp == 0) {
    r = ldexpf((float)man, -9);              // subnormal: man * 2^-3 * 2^-6
  } else if (exp == 15 && man == 7) {
    r = NAN;
  } else {
    r = ldexpf(1.0f + (float)man / 8.0f, (int)exp - 7);
  }
  return sign ? -r : r;
}
inline uint8_t f32_to_e4m3_ref(float f) {
  uint8_t sign = std::signbit(f) ? 0x80 : 0;
  float a = fabsf(f);
  if (std::isnan(f)) return (uint8_t)(sign | 0x7f);
  if (a >= 448.0f) return (uint8_t)(sign | 0x7e);     // saturate to max finite
  if (a < ldexpf(1.0f, -10)) return sign;             // below half of the smallest subnormal (2^-9): RNE -> 0
  int e;
  float m = frexpf(a, &e);                            // a = m * 2^e, m in [0.5,1)
  int exp = e - 1;                                    // a = (2m) * 2^(e-1)
  if (exp < -6) {
    // subnormal: quantum 2^-9
    float q = a * 512.0f;
    float r = nearbyintf(q);
    uint32_t mi = (uint32_t)r;
    if (mi >= 8) return (uint8_t)(sign | (1u << 3));  // rounds up to the smallest normal
    return (uint8_t)(sign | mi);
  }
  float frac = (2.0f * m - 1.0f) * 8.0f;              // [0,8)
  float r = nearbyintf(frac);
  uint32_t mi = (uint32_t)r;
  uint32_t be = (uint32_t)(exp + 7);
  if (mi == 8) {
    mi = 0;
    be++;
  }
  if (be > 15 || (be == 15 && mi == 7)) return (uint8_t)(sign | 0x7e);
  return (uint8_t)(sign | (be << 3) | mi);
}

struct E4M3Table {
  float v[256];
  E4M3Table() {
    for (int i = 0; i < 256; ++i) v[i] = e4m3_to_f32_ref((uint8_t)i);
  }
};
inline const E4M3Table kE4M3Table{};   // built while the library loads; a function-local static would cost a guard per call
inline float e4m3_to_f32(uint8_t v) { return kE4M3Table.v[v]; }   // (an arithmetic decode measured 2-3x slower)
// Round-to-nearest-even conversion on the bit pattern; identical to f32_to_e4m3_ref for every float
// (csrc/tests/quant_codec_check.cpp compares all 2^32 inputs).
inline uint8_t f32_to_e4m3(float f) {
  // straight-line code (selects, no branches): loops over it vectorise
  uint32_t u;
  memcpy(&u, &f, 4);
  const uint32_t sign = (u >> 24) & 0x80u;
  u &= 0x7fffffffu;
  float a;
  memcpy(&a, &u, 4);
  const float t = a * 512.0f + 12582912.0f;              // |f| < 2^-6: subnormal grid, quantum 2^-9; 1.5 * 2^23 makes the
  uint32_t ti;                                            // sum round to an integer (RNE): 0 .. 8, 8 = smallest normal
  memcpy(&ti, &t, 4);
  const uint32_t sub = ti - 0x4b400000u;
  const uint32_t r = u + 0x7ffffu + ((u >> 20) & 1u);     // RNE at bit 20: 3 mantissa bits survive
  uint32_t code = (r >> 20) - 0x3c0u;                     // (exponent - 120) << 3 | mantissa
  code = code > 0x7eu ? 0x7eu : code;                     // also covers |f| >= 448: saturate to the largest finite value
  code = u < 0x3c800000u ? sub : code;
  code = u >= 0x43e00000u ? 0x7eu : code;
  code = u > 0x7f800000u ? 0x7fu : code;                  // NaN
  return (uint8_t)(sign | code);
}

}  // namespace mlslb
