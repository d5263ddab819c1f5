// BASELINE config #1: a single AllReduce of COUNT=128 floats across the world; every rank contributes its index,
// so every element must equal (P-1)*P/2.  (Same check as reference mlsl_to_oneccl/mlsl_sample.cpp:24-71; uses
// only API that exists in the reference, so this file also compiles against the reference headers.)
#include <cstdio>

#include "mlsl.hpp"

using namespace MLSL;
#define COUNT 128

int main(int argc, char** argv) {
  Environment& env = Environment::GetEnv();
  env.Init(&argc, &argv);
  size_t rank = env.GetProcessIdx(), size = env.GetProcessCount();
  Distribution* dist = env.CreateDistribution(size, 1);
  float* buf = (float*)env.Alloc(COUNT * sizeof(float), 64);
  for (int i = 0; i < COUNT; ++i) buf[i] = (float)rank;
  CommReq* req = dist->AllReduce(buf, buf, COUNT, DT_FLOAT, RT_SUM, GT_DATA);
  env.Wait(req);
  float expected = (float)((size - 1) * size / 2);
  int bad = 0;
  for (int i = 0; i < COUNT; ++i)
    if (buf[i] != expected) ++bad;
  printf("[%zu] mlsl_sample: %s (expected %.0f, %d mismatches)\n", rank, bad ? "FAILED" : "PASSED", expected, bad);
  env.Free(buf);
  env.DeleteDistribution(dist);
  env.Finalize();
  return bad ? 1 : 0;
}
