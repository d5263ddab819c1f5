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

}  // namespace mlslb
