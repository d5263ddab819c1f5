// Which CUDA runtime calls block the calling host thread while ANOTHER stream of the same context runs a kernel that
// spins for a peer?  Loop-back ranks (several ranks = threads of one process on one GPU, the single-GPU test mode of
// this repository) dead-lock on any such call: the spinning kernel waits for the rank whose host thread is stuck in it.
// Usage: probe_blocking [spin_ms] [holder]  - prints one line per API call with the time it took while the spinner was
// live.  With `holder` = 1 a second host thread sits in a pageable device-to-host copy queued BEHIND the spinning kernel
// (what `tensor.cpu()` of a rank does while its collective waits for a peer): the driver holds a reader lock for the
// whole copy, and every call that needs the writer side (cuEventCreate, ...) blocks - the dead-lock class found in the
// round-2 device fuzz test.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define CK(x)                                                                             \
  do {                                                                                    \
    cudaError_t e_ = (x);                                                                 \
    if (e_ != cudaSuccess) {                                                              \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
      exit(1);                                                                            \
    }                                                                                     \
  } while (0)

__global__ void k_spin(volatile int* flag, unsigned long long timeout_ns) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  while (*flag == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    if (t - t0 > timeout_ns) break;
  }
}
__global__ void k_first_use_a(int* p) { if (p) p[0] = 1; }
__global__ void k_first_use_stack(int* p) {
  volatile int big[4096];
  for (int i = 0; i < 4096; ++i) big[i] = i;
  if (p) p[0] = big[threadIdx.x];
}
__global__ void k_used_before(int* p) { if (p) p[1] = 2; }

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
  const double spin_ms = argc > 1 ? atof(argv[1]) : 1500.0;
  const bool holder = argc > 2 && atoi(argv[2]) != 0;
  CK(cudaSetDevice(0));
  volatile int* flag_h;
  int* flag_d;
  CK(cudaHostAlloc((void**)&flag_h, 64, cudaHostAllocMapped));
  CK(cudaHostGetDevicePointer((void**)&flag_d, (void*)flag_h, 0));
  cudaStream_t spin_s, s2;
  CK(cudaStreamCreateWithFlags(&spin_s, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
  int* dbuf;
  CK(cudaMalloc(&dbuf, 1 << 20));
  k_used_before<<<1, 32, 0, s2>>>(dbuf);
  CK(cudaStreamSynchronize(s2));
  cudaEvent_t pre_ev;
  CK(cudaEventCreateWithFlags(&pre_ev, cudaEventDisableTiming));
  std::vector<char> pageable(1 << 20, 1);
  char* pinned;
  CK(cudaHostAlloc((void**)&pinned, 1 << 20, cudaHostAllocDefault));

  struct Probe {
    const char* name;
    void (*fn)(cudaStream_t, int*, void*, void*, cudaEvent_t);
  };
  static std::vector<cudaEvent_t> evs;
  Probe probes[] = {
      {"cudaEventRecord(existing event)", [](cudaStream_t s, int*, void*, void*, cudaEvent_t e) { CK(cudaEventRecord(e, s)); }},
      {"cudaEventCreateWithFlags x1", [](cudaStream_t, int*, void*, void*, cudaEvent_t) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); evs.push_back(e); }},
      {"cudaEventCreateWithFlags x4096", [](cudaStream_t, int*, void*, void*, cudaEvent_t) { for (int i = 0; i < 4096; ++i) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); evs.push_back(e); } }},
      {"cudaEventCreate (timing) x256 + record", [](cudaStream_t s, int*, void*, void*, cudaEvent_t) { for (int i = 0; i < 256; ++i) { cudaEvent_t e; CK(cudaEventCreate(&e)); CK(cudaEventRecord(e, s)); evs.push_back(e); } }},
      {"cudaEventDestroy x512", [](cudaStream_t, int*, void*, void*, cudaEvent_t) { for (int i = 0; i < 512 && !evs.empty(); ++i) { CK(cudaEventDestroy(evs.back())); evs.pop_back(); } }},
      {"kernel launch (used before)", [](cudaStream_t s, int* d, void*, void*, cudaEvent_t) { k_used_before<<<1, 32, 0, s>>>(d); CK(cudaGetLastError()); }},
      {"kernel launch (first use, same module)", [](cudaStream_t s, int* d, void*, void*, cudaEvent_t) { k_first_use_a<<<1, 32, 0, s>>>(d); CK(cudaGetLastError()); }},
      {"kernel launch (first use, 16 KB stack/thread)", [](cudaStream_t s, int* d, void*, void*, cudaEvent_t) { k_first_use_stack<<<1, 32, 0, s>>>(d); CK(cudaGetLastError()); }},
      {"cudaMemcpyAsync H2D pageable 4 KB", [](cudaStream_t s, int* d, void* pg, void*, cudaEvent_t) { CK(cudaMemcpyAsync(d, pg, 4096, cudaMemcpyHostToDevice, s)); }},
      {"cudaMemcpyAsync H2D pageable 1 MB", [](cudaStream_t s, int* d, void* pg, void*, cudaEvent_t) { CK(cudaMemcpyAsync(d, pg, 1 << 20, cudaMemcpyHostToDevice, s)); }},
      {"cudaMemcpyAsync D2H pageable 4 KB + sync", [](cudaStream_t s, int* d, void* pg, void*, cudaEvent_t) { CK(cudaMemcpyAsync(pg, d, 4096, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s)); }},
      {"cudaMemcpyAsync H2D pinned 1 MB", [](cudaStream_t s, int* d, void*, void* pn, cudaEvent_t) { CK(cudaMemcpyAsync(d, pn, 1 << 20, cudaMemcpyHostToDevice, s)); }},
      {"cudaMemsetAsync 1 MB", [](cudaStream_t s, int* d, void*, void*, cudaEvent_t) { CK(cudaMemsetAsync(d, 0, 1 << 20, s)); }},
      {"cudaStreamSynchronize(own stream)", [](cudaStream_t s, int*, void*, void*, cudaEvent_t) { CK(cudaStreamSynchronize(s)); }},
      {"cudaMalloc 4 KB", [](cudaStream_t, int*, void*, void*, cudaEvent_t) { void* p; CK(cudaMalloc(&p, 4096)); }},
      {"cudaMalloc 64 MB", [](cudaStream_t, int*, void*, void*, cudaEvent_t) { void* p; CK(cudaMalloc(&p, 64 << 20)); }},
      {"cudaMallocAsync 1 MB", [](cudaStream_t s, int*, void*, void*, cudaEvent_t) { void* p; CK(cudaMallocAsync(&p, 1 << 20, s)); }},
      {"cudaHostAlloc 1 MB", [](cudaStream_t, int*, void*, void*, cudaEvent_t) { void* p; CK(cudaHostAlloc(&p, 1 << 20, cudaHostAllocDefault)); }},
      {"cudaStreamCreateWithFlags", [](cudaStream_t, int*, void*, void*, cudaEvent_t) { cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); }},
      {"cudaStreamCreateWithPriority(high)", [](cudaStream_t, int*, void*, void*, cudaEvent_t) { int lo, hi; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi)); cudaStream_t s; CK(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi)); }},
      {"cudaPointerGetAttributes", [](cudaStream_t, int* d, void*, void*, cudaEvent_t) { cudaPointerAttributes a; CK(cudaPointerGetAttributes(&a, d)); }},
      {"cudaStreamWaitEvent", [](cudaStream_t s, int*, void*, void*, cudaEvent_t e) { CK(cudaStreamWaitEvent(s, e, 0)); }},
      {"cudaFuncSetAttribute(max dyn smem)", [](cudaStream_t, int*, void*, void*, cudaEvent_t) { CK(cudaFuncSetAttribute(k_first_use_a, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536)); }},
      {"cudaMemGetInfo", [](cudaStream_t, int*, void*, void*, cudaEvent_t) { size_t f, t; CK(cudaMemGetInfo(&f, &t)); }},
      {"cudaFree 4 KB (expected to block)", [](cudaStream_t, int*, void*, void*, cudaEvent_t) { static void* p = nullptr; if (!p) { CK(cudaMalloc(&p, 4096)); } else { CK(cudaFree(p)); p = nullptr; } }},
  };
  printf("%-52s %12s\n", "call (while a peer kernel spins)", "host ms");
  for (auto& pr : probes) {
    if (!strncmp(pr.name, "cudaFree", 8)) pr.fn(s2, dbuf, pageable.data(), pinned, pre_ev);   // allocate outside the spin
    *flag_h = 0;
    k_spin<<<1, 32, 0, spin_s>>>(flag_d, (unsigned long long)(spin_ms * 1e6));
    CK(cudaGetLastError());
    std::thread hold;
    if (holder) {
      hold = std::thread([&] {
        static std::vector<char> dst(4096);
        cudaSetDevice(0);
        cudaMemcpyAsync(dst.data(), dbuf, 4096, cudaMemcpyDeviceToHost, spin_s);   // returns when the spinner ends
      });
      std::this_thread::sleep_for(std::chrono::milliseconds(50));
    }
    double t0 = now_ms();
    pr.fn(s2, dbuf, pageable.data(), pinned, pre_ev);
    double dt = now_ms() - t0;
    *flag_h = 1;
    if (holder) hold.join();
    CK(cudaStreamSynchronize(spin_s));
    CK(cudaStreamSynchronize(s2));
    printf("%-52s %12.3f %s\n", pr.name, dt, dt > spin_ms * 0.5 ? "  <-- BLOCKS until the kernel ends" : "");
    fflush(stdout);
  }
  return 0;
}
