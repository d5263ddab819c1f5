// Host-callable launchers of the sm_100a collective kernels (implemented in csrc/cuda/*.cu).
#pragma once
#include <cuda_runtime.h>

#include "core/common.hpp"
#include "core/runtime.hpp"
#include "cuda/dev_comm.cuh"

namespace mlslb {

constexpr int kCommThreads = 512;

// One source segment of a gather-like collective: copy `bytes` from peer's send buffer (+src_off, + the peer's
// published aux word * elem_size when use_aux) to my recv buffer + dst_off.
struct CopySeg {
  int peer;
  int use_aux;
  unsigned long long src_off, dst_off, bytes;
  // 2-D form (rows > 1): `rows` rows of `row_bytes` (bytes = rows * row_bytes), `src_stride` / `dst_stride` apart
  unsigned long long rows, row_bytes, src_stride, dst_stride;
};
struct CopyPlan {
  int nseg;
  int elem_size;
  int mc_mode;       // NVLS push instead of the pull, when the members' buffers turn out symmetric: 0 = never,
                     // 1 = bcast (mc_root multimem.st's mc_bytes of its buffer), 2 = all-gather (everyone pushes its shard)
  int mc_root;
  unsigned long long mc_bytes;
  int pairs_concurrent;   // 1: deal the channels to the segments (MLSL_ALLTOALL(V)_SPLIT=0), 0: every segment on all channels
  CopySeg seg[kMaxDevRanks];
  unsigned long long aux_out[kMaxDevRanks];   // word I publish to peer p in the opening handshake
};

// K8 barrier
cudaError_t launch_barrier(const DevComm& dc, cudaStream_t s);
// K1 all-reduce: fused reduce-scatter + all-gather over peer memory, scale epilogue
// p2p_cta / p2p_frac: hybrid for multicast groups - the last p2p_cta CTAs move the tail p2p_frac of the message with the
// peer-to-peer code while the others use multimem (0 = multicast only)
cudaError_t launch_allreduce(const DevComm& dc, DType dt, RedOp op, unsigned long long send_off,
                             unsigned long long recv_off, size_t count, float scale, int channels, int unroll, int p2p_cta,
                             float p2p_frac, cudaStream_t s);
// K1, latency path: messages <= kLLMaxBytes travel as (data, flag) pairs pushed straight into every peer's arena -
// one NVLink one-way trip, no handshake, no fence (csrc/cuda/kernels.cu: k_allreduce_ll)
constexpr size_t kLLMaxBytes = 8192;                                   // payload per rank
constexpr size_t kLLSlotBytes = 2 * kLLMaxBytes;                       // packed: 8 B data -> 16 B
constexpr size_t kLLRowBytes = 2 * (size_t)kMaxDevRanks * kLLSlotBytes;   // two parities x sources
constexpr int kLLRows = 32;                                            // group rows that own an arena
cudaError_t launch_allreduce_ll(const DevComm& dc, DType dt, RedOp op, const void* send, void* recv, size_t count,
                                float scale, cudaStream_t s);
// K1, mid sizes: multi-CTA flag-in-data kernel, one-shot or two-shot (csrc/cuda/kernels.cu: k_allreduce_mid).  The arena
// of a row holds two parities; one parity fits the larger of a one-shot (P copies of 2 x bytes) and a two-shot
// (reduce area + gather area, 2 x bytes each, slices rounded up) of the largest message.
constexpr int kMidThreads = 512;
constexpr int kMidCtas = 32;                                           // fixed grid: every launch counts on all of them
constexpr size_t kMidMaxBytes = (size_t)1 << 20;                       // largest message on this path
constexpr size_t kMidParityBytes = 4 * kMidMaxBytes + ((size_t)64 << 10);
constexpr size_t kMidRowBytes = 2 * kMidParityBytes;
constexpr int kMidRows = 8;                                            // group rows that own a mid arena
cudaError_t launch_allreduce_mid(const DevComm& dc, DType dt, RedOp op, const void* send, void* recv, size_t count,
                                 float scale, bool two_shot, int ctas, cudaStream_t s);
// K2/K5 reduce-scatter and reduce: out[i] = scale * op_p send_p[base + i], i < count (active ranks only)
cudaError_t launch_reduce_pull(const DevComm& dc, DType dt, RedOp op, unsigned long long send_off,
                               unsigned long long recv_off, size_t base, size_t count, float scale, bool active,
                               int channels, cudaStream_t s);
// K3/K4/K6/K7/K9 all-gather(v), bcast, all-to-all(v), gather, scatter, send/recv list
cudaError_t launch_pull_copy(const DevComm& dc, const CopyPlan& plan, unsigned long long send_off,
                             unsigned long long recv_off, int channels, bool bulk, cudaStream_t s);
// one-time kernel attributes (dynamic shared memory of the bulk-copy ring); call once per process before the first launch
cudaError_t init_kernel_attributes();
// K13 activation pack / unpack (strided (mb, fm, fmSize) gather/scatter), local
struct PackPlan {
  int n;
  BlockDesc b[kMaxDevRanks];
};
cudaError_t launch_pack_blocks(const PackPlan& plan, size_t local_fm_count, int elem_size, const void* src, void* dst,
                               bool unpack, size_t max_block_elems, cudaStream_t s);
// K11 fused fp8 block-quantised all-reduce with error feedback (fp32 in/out)
cudaError_t launch_allreduce_quant(const DevComm& dc, unsigned long long send_off, unsigned long long recv_off,
                                   unsigned long long stage_off, float* residual, size_t count, float scale,
                                   int channels, bool mx, cudaStream_t s);
size_t allreduce_quant_stage_bytes(size_t count);
// fused distributed update: reduce-scatter + optimizer + all-gather in one kernel
struct FusedUpdateArgs {
  int optimizer;
  float lr, momentum, beta1, beta2, eps, weight_decay, grad_scale, bc1, bc2;
  float* master;
  float* state1;
  float* state2;
};
cudaError_t launch_fused_update(const DevComm& dc, DType grad_dt, DType param_dt, unsigned long long grad_off,
                                unsigned long long param_off, size_t owned, const FusedUpdateArgs& a, int channels,
                                cudaStream_t s);
// K14: tcgen05 GEMM whose epilogue reduce-scatters the partial sums over peer memory (csrc/cuda/gemm_rs.cu)
size_t gemm_rs_stage_bytes(int M, int N);
int gemm_rs_channels(int M, int N, int max_channels);
const char* gemm_rs_check(int M, int N, int K, int P);   // nullptr when the shape is supported
// experimental: all-gather of row shards fused with the GEMM that consumes them (csrc/cuda/ag_gemm.cu)
const char* ag_gemm_check(int M, int N, int K, int P);
size_t ag_gemm_scratch_bytes(int M, int max_ctas);
void ag_gemm_grid(int M, int N, int P, int max_ctas, int* copy_ctas, int* total_ctas);
cudaError_t launch_ag_gemm(const DevComm& dc, unsigned long long x_off, const void* w, void* gathered, void* out, bool out_fp32,
                           int M, int N, int K, int copy_ctas, int total_ctas, void* scratch, cudaStream_t s);
// experimental cta_group::2 form (MLSL_GEMM_2CTA=1): 256 x 256 tiles on CTA pairs
const char* gemm_rs2_check(int M, int N, int K, int P);
int gemm_rs2_channels(int M, int N, int max_channels);
cudaError_t launch_gemm_rs2(const DevComm& dc, const void* a, const void* w, unsigned long long stage_off, void* out,
                            bool out_fp32, int M, int N, int K, int channels, cudaStream_t s);
cudaError_t launch_gemm_rs(const DevComm& dc, const void* a, const void* w, unsigned long long stage_off, void* out,
                           bool out_fp32, int M, int N, int K, int channels, cudaStream_t s);
// local elementwise helper (scale in place) for single-rank groups
cudaError_t launch_scale(DType dt, void* buf, size_t count, float scale, cudaStream_t s);
cudaError_t launch_scale_copy(DType dt, void* dst, const void* src, size_t count, float scale, cudaStream_t s);

}  // namespace mlslb
