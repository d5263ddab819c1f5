// The native path from C++: MLSL_BACKEND=cuda, device memory from Environment::Alloc, collectives ordered on the
// caller's CUDA stream.  One rank per GPU:
//     bin/mlslrun -n 8 bin/mlsl_example_cuda          (rank r uses GPU LOCAL_RANK = r; every GPU stays visible to every rank)
// Each rank fills a gradient buffer on its GPU, the fused all-reduce averages it over the data group (one kernel: pull
// from the peers over NVLink, reduce, scale by 1/N, push), a second all-reduce moves the same data as block-scaled FP8.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mlsl.hpp"

using namespace MLSL;

#define CUDA_OK(call)                                                                  \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      fprintf(stderr, "%s failed: %s\n", #call, cudaGetErrorString(e_));               \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

__global__ void fill(float* p, size_t n, float v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

int main(int argc, char** argv) {
  setenv("MLSL_BACKEND", "cuda", 0);
  Environment& env = Environment::GetEnv();
  env.Init(&argc, &argv);
  const size_t rank = env.GetProcessIdx(), world = env.GetProcessCount();
  cudaStream_t stream;
  CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  env.SetStream(stream);             // collectives start after the work already queued here ...
  env.SetWaitMode("stream");         // ... and Wait() orders this stream after them instead of blocking the CPU
  if (rank == 0) printf("backend: %s\n", env.DescribeBackend());

  const size_t n = (size_t)16 << 20;   // 64 MiB of fp32 gradients
  float* grad = (float*)env.Alloc(n * sizeof(float), 256);   // symmetric heap: peers address it directly
  float* avg = (float*)env.Alloc(n * sizeof(float), 256);
  Distribution* dist = env.CreateDistribution(world, 1);

  fill<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(grad, n, (float)(rank + 1));
  env.Wait(dist->AllReduceEx(grad, avg, n, DT_FLOAT, RT_SUM, GT_DATA, 1.0f / (float)world, CT_NONE));
  std::vector<float> host(4);
  CUDA_OK(cudaMemcpyAsync(host.data(), avg, 4 * sizeof(float), cudaMemcpyDeviceToHost, stream));
  CUDA_OK(cudaStreamSynchronize(stream));
  const float want = (float)(world + 1) / 2.0f;
  bool ok = std::fabs(host[0] - want) < 1e-5f;

  env.Wait(dist->AllReduceEx(grad, avg, n, DT_FLOAT, RT_SUM, GT_DATA, 1.0f / (float)world, CT_QUANTIZATION));
  CUDA_OK(cudaMemcpyAsync(host.data(), avg, 4 * sizeof(float), cudaMemcpyDeviceToHost, stream));
  CUDA_OK(cudaStreamSynchronize(stream));
  ok = ok && std::fabs(host[0] - want) < 0.07f * want;   // fp8 transport: a few per cent

  printf("[%zu] mean of ranks' gradients = %.4f (expected %.4f): %s\n", rank, host[0], want, ok ? "PASSED" : "FAILED");
  env.Free(grad);
  env.Free(avg);
  env.DeleteDistribution(dist);
  env.Finalize();
  CUDA_OK(cudaStreamDestroy(stream));
  return ok ? 0 : 1;
}
