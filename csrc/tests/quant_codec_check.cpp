// Exhaustive check of the fast FP8 (E4M3) codec against its definition: all 2^32 float bit patterns (or every STRIDE-th
// with an argument), and the 256 decode table entries.   quant_codec_check [stride]
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "core/numeric.hpp"

int main(int argc, char** argv) {
  const unsigned long long stride = argc > 1 ? strtoull(argv[1], 0, 10) : 1;
  unsigned long long bad = 0, n = 0;
  for (int i = 0; i < 256; ++i) {
    const float a = mlslb::e4m3_to_f32((uint8_t)i), b = mlslb::e4m3_to_f32_ref((uint8_t)i);
    const float t = mlslb::kE4M3Table.v[i];
    if ((memcmp(&a, &b, 4) != 0 && !(a != a && b != b)) || (memcmp(&t, &b, 4) != 0 && !(t != t && b != b))) {
      printf("decode mismatch at code %02x: %g vs %g\n", i, a, b);
      ++bad;
    }
  }
  for (unsigned long long u = 0; u <= 0xffffffffull; u += stride, ++n) {
    const uint32_t bits = (uint32_t)u;
    float f;
    memcpy(&f, &bits, 4);
    if (mlslb::f32_to_e4m3(f) != mlslb::f32_to_e4m3_ref(f)) {
      if (bad < 10) printf("mismatch at %08x (%g): fast %02x ref %02x\n", bits, f, mlslb::f32_to_e4m3(f), mlslb::f32_to_e4m3_ref(f));
      ++bad;
    }
  }
  printf("%llu inputs checked, %llu mismatches\n", n, bad);
  return bad ? 1 : 0;
}
