// Block-scaled FP8 gradient compression format (host definition; the CUDA kernel in csrc/cuda implements the
// identical arithmetic).
//
// The reference ships only the glue for an external int8 "DFP" plugin (256 elements -> 268 bytes, error feedback,
// custom MPI_Op; reference quant/quant.c:96-211, tests/examples/mlsl_test/mlsl_test.cpp:586-600).  On Blackwell
// the natural transport format is FP8 E4M3 with one fp32 scale per 128-element block (132 bytes / 128 elements,
// ratio 3.88 vs fp32), accumulated in fp32 and re-quantised once, so the format is built in; QuantParams is kept
// for API compatibility (block_size = 132, elem_in_block = 128 are reported back).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

#include "numeric.hpp"

namespace mlslb {

constexpr int kQuantBlock = 128;                       // elements per scale
constexpr size_t kQuantBlockBytes = kQuantBlock + 4;   // payload + fp32 scale

// Quantise one block (exactly kQuantBlock floats, zero padded by the caller); returns the scale.
inline float quant_block(const float* v, uint8_t* q) {
  // largest magnitude through the bit patterns (an integer max reduction vectorises; NaN / inf patterns sort above every
  // finite value, so one comparison afterwards finds them)
  uint32_t umax = 0;
  for (int i = 0; i < kQuantBlock; ++i) {
    uint32_t u;
    memcpy(&u, &v[i], 4);
    u &= 0x7fffffffu;
    umax = u > umax ? u : umax;
  }
  float amax;
  memcpy(&amax, &umax, 4);
  if (umax == 0 || umax >= 0x7f800000u) {
    for (int i = 0; i < kQuantBlock; ++i) q[i] = 0;
    return 0.f;
  }
  float scale = amax / 448.0f;
  float inv = 448.0f / amax;
  for (int i = 0; i < kQuantBlock; ++i) q[i] = f32_to_e4m3(v[i] * inv);
  return scale;
}

// ---- MX flavour (MLSL_QUANT_MX=1): one power-of-two scale per 32 elements, stored as a biased exponent byte (ue8m0, the OCP
// microscaling layout Blackwell's block-scaled tensor formats use).  Four exponent bytes take the place of the one fp32 scale of
// a 128-element block, so the wire format has the same size.  scale = 2^ceil(log2(amax / 448)): the block always fits e4m3.
constexpr int kMxBlock = 32;

inline uint8_t mx_exp_of(float amax) {      // 0 = all-zero (or non-finite) block
  uint32_t ua;
  memcpy(&ua, &amax, 4);
  if ((ua & 0x7fffffffu) == 0 || (ua & 0x7fffffffu) >= 0x7f800000u) return 0;
  const float r = amax / 448.0f;
  uint32_t u;
  memcpy(&u, &r, 4);
  const uint32_t E = (u >> 23) & 0xffu, M = u & 0x7fffffu;
  if (E == 0) return 0;                      // amax below 448 * 2^-126: treated as zero
  const uint32_t eb = E + (M ? 1u : 0u);
  return (uint8_t)(eb > 254u ? 254u : eb);
}
inline float mx_scale_of(uint8_t eb) {
  if (!eb) return 0.f;
  const uint32_t u = (uint32_t)eb << 23;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline float mx_inv_of(uint8_t eb) {
  if (!eb) return 0.f;
  const uint32_t u = (uint32_t)(254u - eb) << 23;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
// Quantise one 128-element block as four 32-element MX sub-blocks; exps[4] receives the exponent bytes.
inline void quant_block_mx(const float* v, uint8_t* q, uint8_t* exps) {
  for (int s = 0; s < kQuantBlock / kMxBlock; ++s) {
    float amax = 0.f;
    bool bad = false;
    for (int i = 0; i < kMxBlock; ++i) {
      const float a = std::fabs(v[s * kMxBlock + i]);
      if (!(a == a) || std::isinf(a)) bad = true;
      amax = a > amax ? a : amax;
    }
    const uint8_t eb = bad ? 0 : mx_exp_of(amax);
    exps[s] = eb;
    const float inv = mx_inv_of(eb);
    for (int i = 0; i < kMxBlock; ++i) q[s * kMxBlock + i] = eb ? f32_to_e4m3(v[s * kMxBlock + i] * inv) : 0;
  }
}

}  // namespace mlslb
