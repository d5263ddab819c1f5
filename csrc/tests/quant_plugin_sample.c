/* Sample gradient-compression plug-in for Environment::SetQuantizationParams (lib_path + three function names).
 *
 * Interface (what the reference's quantization layer dlsym()s, reference quant/quant.c:57-65):
 *   int compress  (void* src, void* dst, size_t count, void* diff, int src_dtype, size_t comp_ratio, int method);
 *   int decompress(void* src, void* dst, size_t count);
 *   int reduce_sum(const void* in, void* inout, size_t block_count);
 * Both buffer functions are called IN PLACE (src == dst, a buffer of `count` floats whose head holds the blocks).
 *
 * Format of this sample: blocks of 256 elements = 12-byte header {float scale, 8 bytes reserved} + 256 int8
 * (268 bytes, the geometry the reference's test configures: block_size 268, elem_in_block 256).  `diff` is the
 * error-feedback residual: what rounding lost this time is added to the next call's input. */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define ELEMS 256
#define HEADER 12
#define BLOCK (HEADER + ELEMS)

static void encode(const float* v, size_t n, unsigned char* out) {
  float amax = 0.f;
  for (size_t i = 0; i < n; ++i) {
    float a = fabsf(v[i]);
    if (a > amax) amax = a;
  }
  float scale = amax > 0.f ? amax / 127.f : 1.f;
  memset(out, 0, BLOCK);
  memcpy(out, &scale, sizeof(float));
  int8_t* q = (int8_t*)(out + HEADER);
  for (size_t i = 0; i < n; ++i) {
    float r = nearbyintf(v[i] / scale);
    if (r > 127.f) r = 127.f;
    if (r < -127.f) r = -127.f;
    q[i] = (int8_t)r;
  }
}

static void decode(const unsigned char* in, float* v) {
  float scale;
  memcpy(&scale, in, sizeof(float));
  const int8_t* q = (const int8_t*)(in + HEADER);
  for (int i = 0; i < ELEMS; ++i) v[i] = (float)q[i] * scale;
}

int sample_compress(void* src, void* dst, size_t count, void* diff, int src_dtype, size_t comp_ratio, int method) {
  (void)comp_ratio;
  (void)method;
  if (src_dtype != 2) return 1; /* float32 only */
  const float* x = (const float*)src;
  float* res = (float*)diff;
  unsigned char* out = (unsigned char*)dst;
  size_t nblk = (count + ELEMS - 1) / ELEMS;
  for (size_t b = 0; b < nblk; ++b) { /* ascending: block b's output ends before block b+1's input starts */
    size_t lo = b * ELEMS, n = count - lo < ELEMS ? count - lo : ELEMS;
    float v[ELEMS], back[ELEMS];
    unsigned char blk[BLOCK];
    for (size_t i = 0; i < ELEMS; ++i) v[i] = i < n ? x[lo + i] + (res ? res[lo + i] : 0.f) : 0.f;
    encode(v, ELEMS, blk);
    decode(blk, back);
    if (res)
      for (size_t i = 0; i < n; ++i) res[lo + i] = v[i] - back[i];
    memcpy(out + b * BLOCK, blk, BLOCK);
  }
  return 0;
}

int sample_decompress(void* src, void* dst, size_t count) {
  const unsigned char* in = (const unsigned char*)src;
  float* y = (float*)dst;
  size_t nblk = (count + ELEMS - 1) / ELEMS;
  for (size_t b = nblk; b-- > 0;) { /* descending: block b's floats land behind every block still to be read */
    size_t lo = b * ELEMS, n = count - lo < ELEMS ? count - lo : ELEMS;
    unsigned char blk[BLOCK];
    float v[ELEMS];
    memcpy(blk, in + b * BLOCK, BLOCK);
    decode(blk, v);
    memcpy(y + lo, v, n * sizeof(float));
  }
  return 0;
}

int sample_reduce_sum(const void* in, void* inout, size_t block_count) {
  const unsigned char* a = (const unsigned char*)in;
  unsigned char* b = (unsigned char*)inout;
  for (size_t k = 0; k < block_count; ++k) {
    float va[ELEMS], vb[ELEMS];
    decode(a + k * BLOCK, va);
    decode(b + k * BLOCK, vb);
    for (int i = 0; i < ELEMS; ++i) vb[i] += va[i];
    encode(vb, ELEMS, b + k * BLOCK);
  }
  return 0;
}

/* The same three entry points under the names the reference's own functional test asks for (it was written against
 * Intel's dl_comp library, tests/examples/mlsl_test/mlsl_test.cpp:590-592), so that test can run with this plug-in. */
int dl_comp_compress_buffer(void* src, void* dst, size_t count, void* diff, int src_dtype, size_t comp_ratio, int method) {
  return sample_compress(src, dst, count, diff, src_dtype, comp_ratio, method);
}
int dl_comp_decompress_buffer(void* src, void* dst, size_t count) { return sample_decompress(src, dst, count); }
int dl_comp_compressed_buffer_reduce_sum(const void* in, void* inout, size_t block_count) {
  return sample_reduce_sum(in, inout, block_count);
}
