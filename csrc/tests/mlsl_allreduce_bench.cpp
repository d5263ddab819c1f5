// Host-pointer AllReduce bandwidth sweep through the public API (Environment::Alloc buffers, fp32 SUM,
// Distribution::AllReduce + Environment::Wait; in place, or send -> recv with MLSL_BENCH_OUT_OF_PLACE=1 - the form
// bench.py times on the GPU).  Uses only calls that exist in the reference API, so the SAME source
// is compiled against the reference's headers/library for the `--impl reference` arm of bench.py and against ours
// for the CPU-plumbing comparison.  Host-timed (both are CPU libraries on this path), max over ranks.
//   mlsl_allreduce_bench <min_bytes> <max_bytes> <iters> <warmup> [factor=4] [exact_bytes...]
// MLSL_BENCH_OP=allreduce (default) | allgather | reducescatter | alltoall | bcast selects the collective; `bytes` is
// always the size of the LARGER of the two buffers (the gathered / scattered / exchanged whole), algbw = bytes / t.
// Prints one JSON object per size on rank 0.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mlsl.hpp"

using namespace MLSL;

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
  size_t minb = argc > 1 ? strtoull(argv[1], 0, 10) : 1024;
  size_t maxb = argc > 2 ? strtoull(argv[2], 0, 10) : (64u << 20);
  int iters = argc > 3 ? atoi(argv[3]) : 20;
  int warm = argc > 4 ? atoi(argv[4]) : 5;
  size_t factor = argc > 5 ? strtoull(argv[5], 0, 10) : 4;
  if (factor < 2) factor = 2;
  Environment& env = Environment::GetEnv();
  env.Init(&argc, &argv);
  size_t rank = env.GetProcessIdx(), P = env.GetProcessCount();
  Distribution* dist = env.CreateDistribution(P, 1);
  float* buf = (float*)env.Alloc(maxb, 4096);
  const char* oop = getenv("MLSL_BENCH_OUT_OF_PLACE");
  const char* opname = getenv("MLSL_BENCH_OP") ? getenv("MLSL_BENCH_OP") : "allreduce";
  enum { AR, AG, RS, A2A, BC } op = !strcmp(opname, "allgather") ? AG : !strcmp(opname, "reducescatter") ? RS
                                    : !strcmp(opname, "alltoall") ? A2A : !strcmp(opname, "bcast") ? BC : AR;
  float* out = ((oop && atoi(oop)) || op == AG || op == RS || op == A2A) ? (float*)env.Alloc(maxb, 4096) : buf;
  if (out != buf) memset(out, 0, maxb);
  auto run = [&](size_t count) -> CommReq* {   // count = elements of the whole
    switch (op) {
      case AG: return dist->AllGather(buf, count / P, out, DT_FLOAT, GT_DATA);
      case RS: return dist->ReduceScatter(buf, out, count / P, DT_FLOAT, RT_SUM, GT_DATA);
      case A2A: return dist->AlltoAll(buf, count / P, out, DT_FLOAT, GT_DATA);
      case BC: return dist->Bcast(buf, count, DT_FLOAT, 0, GT_DATA);
      default: return dist->AllReduce(buf, out, count, DT_FLOAT, RT_SUM, GT_DATA);
    }
  };
  double* tbuf = (double*)env.Alloc(64, 64);
  for (size_t i = 0; i < maxb / 4; ++i) buf[i] = 1.0f;
  std::vector<size_t> sizes;
  for (size_t b = minb; b <= maxb; b *= factor) sizes.push_back(b);
  if (sizes.empty() || sizes.back() != maxb) sizes.push_back(maxb);
  for (size_t bytes : sizes) {
    size_t count = bytes / 4;
    if (!count) continue;
    if (count < P) continue;
    for (int i = 0; i < warm; ++i) env.Wait(run(count));
    dist->Barrier(GT_DATA);
    double t0 = now_s();
    for (int i = 0; i < iters; ++i) env.Wait(run(count));
    double dt = (now_s() - t0) / iters;
    tbuf[0] = dt;
    env.Wait(dist->AllReduce(tbuf, tbuf, 1, DT_DOUBLE, RT_MAX, GT_DATA));
    dt = tbuf[0];
    // keep magnitudes bounded over many in-place sums
    for (size_t i = 0; i < count; ++i) buf[i] = 1.0f;
    if (rank == 0) {
      double algbw = bytes / dt / 1e9;
      double busbw = algbw * (P > 1 ? 2.0 * (P - 1) / P : 1.0);
      printf("{\"op\": \"%s\", \"bytes\": %zu, \"us\": %.3f, \"algbw_GBps\": %.4f, \"busbw_GBps\": %.4f, \"ranks\": %zu}\n", opname, bytes,
             dt * 1e6, algbw, busbw, P);
      fflush(stdout);
    }
  }
  if (out != buf) {
    if (op == AR && out[0] != (float)P) fprintf(stderr, "rank %zu: unexpected result %f (expected %zu)\n", rank, out[0], P);
    env.Free(out);
  }
  env.Free(buf);
  env.Free(tbuf);
  env.DeleteDistribution(dist);
  env.Finalize();
  return 0;
}
