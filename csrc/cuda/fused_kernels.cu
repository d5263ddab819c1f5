// Fused compute + collective kernels (sm_100a):
//   * k_allreduce_quant : gradient all-reduce with block-scaled FP8 (E4M3) transport and error feedback - quantise,
//                         reduce-scatter (dequant-accumulate in fp32), re-quantise, all-gather, dequantise, scale:
//                         ONE kernel, 1/4 of the NVLink bytes of an fp32 all-reduce.  Replaces the reference's
//                         dlopen()'d quantisation plugin + custom MPI_Op executed on the endpoint servers
//                         (reference quant/quant.c:96-211, eplib/cqueue.c:1977-1994,2283-2284).
//   * k_fused_update    : distributed weight update - gradient reduce-scatter + optimizer step on the owned shard +
//                         parameter all-gather (with the fp32->bf16 cast) as ONE kernel.  The reference needs
//                         ReduceScatter -> host optimizer loop -> AllGather (src/mlsl_impl.cpp:401-433,504-539).
#include <cuda_fp8.h>

#include "core/quant.hpp"
#include "cuda/kernels.hpp"

namespace mlslb {

static_assert(kQuantBlock == 128, "one warp handles one 128-element block (32 lanes x 4 elements)");

__host__ __device__ __forceinline__ size_t ru256(size_t v) { return (v + 255) & ~(size_t)255; }

// stage layout (per rank): [q1 | s1 | q2 | s2], block counts rounded up to whole superblocks (4 blocks = 512 elements:
// one warp pass with 16 bytes per lane)
__host__ __device__ __forceinline__ size_t quant_nblk4(size_t count) {
  const size_t nblk = (count + kQuantBlock - 1) / kQuantBlock;
  return (nblk + 3) & ~(size_t)3;
}
__host__ __device__ __forceinline__ size_t allreduce_quant_stage_bytes_dev(size_t count) {
  const size_t nb = quant_nblk4(count);
  return 2 * (ru256(nb * kQuantBlock) + ru256(nb * sizeof(float)));
}
size_t allreduce_quant_stage_bytes(size_t count) { return allreduce_quant_stage_bytes_dev(count); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ unsigned pack_e4m3x4(float a, float b, float c, float d) {
  unsigned lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
  unsigned hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
  return (lo & 0xffffu) | (hi << 16);
}
__device__ __forceinline__ float4 unpack_e4m3x4(unsigned w) {
  __half2_raw h01 = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w & 0xffffu), __NV_E4M3);
  __half2_raw h23 = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w >> 16), __NV_E4M3);
  float2 f01 = __half22float2(*reinterpret_cast<__half2*>(&h01));
  float2 f23 = __half22float2(*reinterpret_cast<__half2*>(&h23));
  return make_float4(f01.x, f01.y, f23.x, f23.y);
}
// quantise the 128 values a warp holds (4 per lane); returns the block scale (same arithmetic as quant_block())
__device__ __forceinline__ float quant_warp_block(const float4& v, unsigned& packed) {
  float amax = warp_max(fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  if (!(amax > 0.f) || !isfinite(amax)) {
    packed = 0;
    return 0.f;
  }
  const float scale = __fdiv_rn(amax, 448.0f);
  const float inv = __fdiv_rn(448.0f, amax);
  packed = pack_e4m3x4(__fmul_rn(v.x, inv), __fmul_rn(v.y, inv), __fmul_rn(v.z, inv), __fmul_rn(v.w, inv));
  return scale;
}

// MX flavour: biased exponent byte of the power-of-two scale of a sub-block with magnitude `amax` (same integer arithmetic as
// mx_exp_of() in core/quant.hpp), and the scale / inverse scale it stands for
__device__ __forceinline__ unsigned mx_exp_dev(float amax) {
  if (!(amax > 0.f) || !isfinite(amax)) return 0u;
  const unsigned u = __float_as_uint(__fdiv_rn(amax, 448.0f));
  const unsigned E = (u >> 23) & 0xffu, M = u & 0x7fffffu;
  if (E == 0u) return 0u;
  const unsigned eb = E + (M ? 1u : 0u);
  return eb > 254u ? 254u : eb;
}
__device__ __forceinline__ float mx_scale_dev(unsigned eb) { return eb ? __uint_as_float(eb << 23) : 0.f; }
__device__ __forceinline__ float mx_inv_dev(unsigned eb) { return eb ? __uint_as_float((254u - eb) << 23) : 0.f; }

// The local phases (quantise / dequantise) are HBM bound, so unlike the pure NVLink kernels this one wants the whole
// GPU: 1024-thread CTAs, up to one per SM.
constexpr int kQuantThreads = 1024;
template <bool kMx>
__global__ void __launch_bounds__(kQuantThreads) k_allreduce_quant(DevComm dc, unsigned long long send_off,
                                                                  unsigned long long recv_off,
                                                                  unsigned long long stage_off, float* residual,
                                                                  size_t count, float out_scale) {
  __shared__ PeerTable pt;
  __shared__ int s_symmetric;
  // peers need my staging area only; x / y are touched by this rank alone
  const unsigned long long t = comm_begin(dc, pt, stage_off, stage_off, NoAux());
  const int P = dc.nranks, me = dc.me;
  if (threadIdx.x == 0) {
    int sym = dc.mc != nullptr && !pt.failed && stage_off + allreduce_quant_stage_bytes_dev(count) <= dc.mc_bytes;
    for (int p = 0; p < P; ++p) sym &= (pt.send[p] - dc.slab[p]) == (long long)stage_off;
    s_symmetric = sym;
  }
  const float* x = reinterpret_cast<const float*>(dc.slab[me] + send_off);
  float* y = reinterpret_cast<float*>(dc.slab[me] + recv_off);
  const size_t nblk = (count + kQuantBlock - 1) / kQuantBlock;
  const size_t nblk4 = quant_nblk4(count), nsb = nblk4 / 4;
  const size_t qbytes = ru256(nblk4 * kQuantBlock), sbytes = ru256(nblk4 * sizeof(float));
  const size_t q2_off = qbytes + sbytes, s2_off = 2 * qbytes + sbytes;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const size_t C = gridDim.x, c = blockIdx.x;
  char* mystage = pt.send[me];

  // ---- phase 1: (x + residual) -> fp8 blocks in my staging area; residual <- quantisation error.  The only pass over
  //      the fp32 input: 4 + 4 bytes read, 4 + 1 written per element, one warp per 128-element block. --------------------
  //      Work is dealt by SUPERBLOCK: superblock sb belongs to channel sb % C in every phase and on every rank, so the
  //      per-channel handshakes below order exactly the producers and consumers of each block.
  for (size_t sb1 = c + (size_t)warp * C; sb1 < nsb; sb1 += C * nwarp)
  for (int j = 0; j < 4; ++j) {
    const size_t b = sb1 * 4 + j;
    const size_t e0 = b * kQuantBlock + (size_t)lane * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), r = v;
    if (e0 + 3 < count) {
      v = *reinterpret_cast<const float4*>(x + e0);
      r = *reinterpret_cast<const float4*>(residual + e0);
    } else {
      if (e0 < count) { v.x = x[e0]; r.x = residual[e0]; }
      if (e0 + 1 < count) { v.y = x[e0 + 1]; r.y = residual[e0 + 1]; }
      if (e0 + 2 < count) { v.z = x[e0 + 2]; r.z = residual[e0 + 2]; }
    }
    v.x = __fadd_rn(v.x, r.x); v.y = __fadd_rn(v.y, r.y); v.z = __fadd_rn(v.z, r.z); v.w = __fadd_rn(v.w, r.w);
    unsigned packed;
    float sc;
    if constexpr (kMx) {
      // 32-element sub-block = 8 lanes: its own power-of-two scale, one exponent byte each
      float am = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
      const unsigned eb = mx_exp_dev(am);
      sc = mx_scale_dev(eb);
      const float inv = mx_inv_dev(eb);
      packed = eb ? pack_e4m3x4(__fmul_rn(v.x, inv), __fmul_rn(v.y, inv), __fmul_rn(v.z, inv), __fmul_rn(v.w, inv)) : 0u;
      if ((lane & 7) == 0) *reinterpret_cast<unsigned char*>(mystage + qbytes + b * sizeof(float) + (lane >> 3)) = (unsigned char)eb;
    } else {
      sc = quant_warp_block(v, packed);
      if (lane == 0) *reinterpret_cast<float*>(mystage + qbytes + b * sizeof(float)) = sc;
    }
    *reinterpret_cast<unsigned*>(mystage + b * kQuantBlock + lane * 4) = packed;
    const float4 dq = unpack_e4m3x4(packed);
    r.x = __fsub_rn(v.x, __fmul_rn(dq.x, sc)); r.y = __fsub_rn(v.y, __fmul_rn(dq.y, sc));
    r.z = __fsub_rn(v.z, __fmul_rn(dq.z, sc)); r.w = __fsub_rn(v.w, __fmul_rn(dq.w, sc));
    if (e0 + 3 < count) {
      *reinterpret_cast<float4*>(residual + e0) = r;
    } else {
      if (e0 < count) residual[e0] = r.x;
      if (e0 + 1 < count) residual[e0 + 1] = r.y;
      if (e0 + 2 < count) residual[e0 + 2] = r.z;
    }
  }
  comm_sync(dc, pt, t, 1, true);

  // ---- phase 2: my slice of superblocks (4 blocks = 512 elements per warp pass, SIXTEEN bytes per lane on the wire):
  //      pull the peers' fp8, accumulate in fp32 in fixed peer order, re-quantise, PUSH the result into every member's
  //      gather area (one multimem.st when the staging areas are symmetric, P plain stores otherwise). -------------------
  const size_t sb_per = (nsb + P - 1) / P;
  const size_t slo = min(nsb, (size_t)me * sb_per), shi = min(nsb, slo + sb_per);
  const int sub = lane >> 3, l8 = lane & 7;                     // block inside the superblock, 16-byte piece inside the block
  const bool mcast = s_symmetric != 0;
  const size_t sbstart = slo + ((c + C - slo % C) % C);         // first superblock of my slice that belongs to this channel
  for (size_t sb = sbstart + (size_t)warp * C; sb < shi; sb += C * nwarp) {
    const size_t b = sb * 4 + sub;
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
    for (int p = 0; p < P; ++p) {
      const char* ps = pt.send[p];
      const uint4 w = ld16(ps + b * kQuantBlock + l8 * 16);
      float sc;
      if constexpr (kMx) {
        const unsigned ew = __ldcg(reinterpret_cast<const unsigned*>(ps + qbytes + b * sizeof(float)));
        sc = mx_scale_dev((ew >> (8 * (l8 >> 1))) & 0xffu);          // my 16 elements sit in sub-block l8 / 2
      } else {
        sc = __ldcg(reinterpret_cast<const float*>(ps + qbytes + b * sizeof(float)));
      }
      const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 dq = unpack_e4m3x4(ww[k]);
        acc[4 * k] = __fadd_rn(acc[4 * k], __fmul_rn(dq.x, sc));
        acc[4 * k + 1] = __fadd_rn(acc[4 * k + 1], __fmul_rn(dq.y, sc));
        acc[4 * k + 2] = __fadd_rn(acc[4 * k + 2], __fmul_rn(dq.z, sc));
        acc[4 * k + 3] = __fadd_rn(acc[4 * k + 3], __fmul_rn(dq.w, sc));
      }
    }
    // block amax over the 8 lanes that share the block (same arithmetic as quant_warp_block / quant_block())
    float amax = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) amax = fmaxf(amax, fabsf(acc[k]));
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    float sc2 = 0.f;
    if constexpr (kMx) {
      amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));     // the lane pair that shares a 32-element sub-block
      const unsigned eb = mx_exp_dev(amax);
      const float inv = mx_inv_dev(eb);
      if (eb) {
        q.x = pack_e4m3x4(__fmul_rn(acc[0], inv), __fmul_rn(acc[1], inv), __fmul_rn(acc[2], inv), __fmul_rn(acc[3], inv));
        q.y = pack_e4m3x4(__fmul_rn(acc[4], inv), __fmul_rn(acc[5], inv), __fmul_rn(acc[6], inv), __fmul_rn(acc[7], inv));
        q.z = pack_e4m3x4(__fmul_rn(acc[8], inv), __fmul_rn(acc[9], inv), __fmul_rn(acc[10], inv), __fmul_rn(acc[11], inv));
        q.w = pack_e4m3x4(__fmul_rn(acc[12], inv), __fmul_rn(acc[13], inv), __fmul_rn(acc[14], inv), __fmul_rn(acc[15], inv));
      }
      // the block's four exponent bytes, assembled on every lane of the 8-lane group (lanes 0, 2, 4, 6 hold them)
      const int g0 = lane & ~7;
      const unsigned ew = __shfl_sync(0xffffffffu, eb, g0) | (__shfl_sync(0xffffffffu, eb, g0 + 2) << 8) |
                          (__shfl_sync(0xffffffffu, eb, g0 + 4) << 16) | (__shfl_sync(0xffffffffu, eb, g0 + 6) << 24);
      sc2 = __uint_as_float(ew);                                    // travels in the fp32 scale slot of the block
    } else {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    }
    if (!kMx && amax > 0.f && isfinite(amax)) {
      sc2 = __fdiv_rn(amax, 448.0f);
      const float inv = __fdiv_rn(448.0f, amax);
      q.x = pack_e4m3x4(__fmul_rn(acc[0], inv), __fmul_rn(acc[1], inv), __fmul_rn(acc[2], inv), __fmul_rn(acc[3], inv));
      q.y = pack_e4m3x4(__fmul_rn(acc[4], inv), __fmul_rn(acc[5], inv), __fmul_rn(acc[6], inv), __fmul_rn(acc[7], inv));
      q.z = pack_e4m3x4(__fmul_rn(acc[8], inv), __fmul_rn(acc[9], inv), __fmul_rn(acc[10], inv), __fmul_rn(acc[11], inv));
      q.w = pack_e4m3x4(__fmul_rn(acc[12], inv), __fmul_rn(acc[13], inv), __fmul_rn(acc[14], inv), __fmul_rn(acc[15], inv));
    }
    // the four block scales of the superblock travel as one 16-byte store from lane 0
    uint4 s4;
    s4.x = __float_as_uint(__shfl_sync(0xffffffffu, sc2, 0));
    s4.y = __float_as_uint(__shfl_sync(0xffffffffu, sc2, 8));
    s4.z = __float_as_uint(__shfl_sync(0xffffffffu, sc2, 16));
    s4.w = __float_as_uint(__shfl_sync(0xffffffffu, sc2, 24));
    const size_t qo = q2_off + b * kQuantBlock + l8 * 16, so = s2_off + sb * 16;
    if (mcast) {
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dc.mc + stage_off + qo), "r"(q.x), "r"(q.y), "r"(q.z), "r"(q.w) : "memory");
      if (lane == 0)
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dc.mc + stage_off + so), "r"(s4.x), "r"(s4.y), "r"(s4.z), "r"(s4.w) : "memory");
    } else {
      for (int k = 0; k < P; ++k) {
        int p = me + k;
        if (p >= P) p -= P;
        st16(pt.send[p] + qo, q);
        if (lane == 0) st16(pt.send[p] + so, s4);
      }
    }
  }
  comm_sync(dc, pt, t, 2, true);

  // ---- phase 3: every slice's reduced blocks now sit in MY gather area: dequantise * out_scale -> y (local only).
  //      No closing handshake: after sync 2 nobody reads remote memory any more, and a peer can only write my gather area
  //      again in the NEXT launch's phase 2, which waits for my next sync 1 - i.e. for the end of this phase. ------------
  for (size_t sb = c + (size_t)warp * C; sb < nsb; sb += C * nwarp) {
    const size_t b = sb * 4 + sub;
    const uint4 w = *reinterpret_cast<const uint4*>(mystage + q2_off + b * kQuantBlock + l8 * 16);
    float sc;
    if constexpr (kMx) {
      const unsigned ew = *reinterpret_cast<const unsigned*>(mystage + s2_off + b * sizeof(float));
      sc = mx_scale_dev((ew >> (8 * (l8 >> 1))) & 0xffu);
    } else {
      sc = *reinterpret_cast<const float*>(mystage + s2_off + b * sizeof(float));
    }
    const unsigned ww[4] = {w.x, w.y, w.z, w.w};
    const size_t e0 = b * kQuantBlock + (size_t)l8 * 16;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 dq = unpack_e4m3x4(ww[k]);
      float4 o;
      o.x = __fmul_rn(__fmul_rn(dq.x, sc), out_scale); o.y = __fmul_rn(__fmul_rn(dq.y, sc), out_scale);
      o.z = __fmul_rn(__fmul_rn(dq.z, sc), out_scale); o.w = __fmul_rn(__fmul_rn(dq.w, sc), out_scale);
      const size_t e = e0 + 4 * k;
      if (e + 3 < count) {
        *reinterpret_cast<float4*>(y + e) = o;
      } else {
        if (e < count) y[e] = o.x;
        if (e + 1 < count) y[e + 1] = o.y;
        if (e + 2 < count) y[e + 2] = o.z;
      }
    }
  }
}

cudaError_t launch_allreduce_quant(const DevComm& dc, unsigned long long send_off, unsigned long long recv_off,
                                   unsigned long long stage_off, float* residual, size_t count, float scale,
                                   int channels, bool mx, cudaStream_t s) {
  if (mx) k_allreduce_quant<true><<<channels, kQuantThreads, 0, s>>>(dc, send_off, recv_off, stage_off, residual, count, scale);
  else k_allreduce_quant<false><<<channels, kQuantThreads, 0, s>>>(dc, send_off, recv_off, stage_off, residual, count, scale);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// fused distributed update
// ------------------------------------------------------------------------------------------------------------
template <typename T> struct Quad;   // 4 consecutive elements <-> float4
template <> struct Quad<float> {
  __device__ __forceinline__ static float4 load(const char* p) {
    uint4 v = ld16(p);
    return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
  }
  __device__ __forceinline__ static void store(char* p, const float4& f) {
    st16(p, make_uint4(__float_as_uint(f.x), __float_as_uint(f.y), __float_as_uint(f.z), __float_as_uint(f.w)));
  }
  __device__ __forceinline__ static float load1(const char* p) { return *(const float*)p; }
  __device__ __forceinline__ static void store1(char* p, float v) { *(float*)p = v; }
};
template <> struct Quad<__nv_bfloat16> {
  __device__ __forceinline__ static float4 load(const char* p) {
    uint2 v;
    asm volatile("ld.global.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16),
                       __uint_as_float(v.y & 0xffff0000u));
  }
  __device__ __forceinline__ static void store(char* p, const float4& f) {
    __nv_bfloat162 a = __floats2bfloat162_rn(f.x, f.y), b = __floats2bfloat162_rn(f.z, f.w);
    unsigned ua = *reinterpret_cast<unsigned*>(&a), ub = *reinterpret_cast<unsigned*>(&b);
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(ua), "r"(ub) : "memory");
  }
  __device__ __forceinline__ static float load1(const char* p) { return __bfloat162float(*(const __nv_bfloat16*)p); }
  __device__ __forceinline__ static void store1(char* p, float v) { *(__nv_bfloat16*)p = __float2bfloat16_rn(v); }
};

struct OptStep {
  FusedUpdateArgs a;
  // one scalar optimizer step; m1/m2 are updated in place
  __device__ __forceinline__ float operator()(float w, float g, float& m1, float& m2) const {
    if (a.optimizer == 0) {   // SGD with (optional) momentum and L2 weight decay, torch.optim.SGD semantics
      g = fmaf(a.weight_decay, w, g);
      if (a.state1) {
        m1 = fmaf(a.momentum, m1, g);
        g = m1;
      }
      return w - a.lr * g;
    }
    // AdamW (decoupled weight decay), torch.optim.AdamW semantics
    m1 = a.beta1 * m1 + (1.f - a.beta1) * g;
    m2 = a.beta2 * m2 + (1.f - a.beta2) * g * g;
    const float mh = m1 / a.bc1, vh = m2 / a.bc2;
    return w - a.lr * (mh / (sqrtf(vh) + a.eps) + a.weight_decay * w);
  }
};

template <typename GT, typename PT>
__global__ void __launch_bounds__(kCommThreads) k_fused_update(DevComm dc, unsigned long long grad_off,
                                                               unsigned long long param_off, size_t owned,
                                                               FusedUpdateArgs a) {
  __shared__ PeerTable pt;
  __shared__ int s_vec;
  const unsigned long long t = comm_begin(dc, pt, grad_off, param_off, NoAux());
  const int P = dc.nranks, me = dc.me;
  if (threadIdx.x == 0) {
    unsigned long long bits = (unsigned long long)(owned * sizeof(GT)) | (unsigned long long)(owned * sizeof(PT)) |
                              (unsigned long long)a.master | (unsigned long long)a.state1 | (unsigned long long)a.state2;
    for (int p = 0; p < P; ++p) bits |= (unsigned long long)pt.send[p] | (unsigned long long)pt.recv[p];
    s_vec = (bits & 15ull) == 0;
  }
  __syncthreads();
  const OptStep step{a};
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
  const size_t base = (size_t)me * owned;
  const size_t nquad = s_vec ? owned / 4 : 0;
  for (size_t q = gtid; q < nquad; q += gsz) {
    const size_t e = q * 4;
    float4 g = Quad<GT>::load(pt.send[me] + (base + e) * sizeof(GT));
    for (int k = 1; k < P; ++k) {
      int p = me + k;
      if (p >= P) p -= P;
      const float4 o = Quad<GT>::load(pt.send[p] + (base + e) * sizeof(GT));
      g.x += o.x; g.y += o.y; g.z += o.z; g.w += o.w;
    }
    g.x *= a.grad_scale; g.y *= a.grad_scale; g.z *= a.grad_scale; g.w *= a.grad_scale;
    float4 w = a.master ? *reinterpret_cast<const float4*>(a.master + e)
                        : Quad<PT>::load(pt.recv[me] + (base + e) * sizeof(PT));
    float4 m1 = a.state1 ? *reinterpret_cast<const float4*>(a.state1 + e) : make_float4(0, 0, 0, 0);
    float4 m2 = a.state2 ? *reinterpret_cast<const float4*>(a.state2 + e) : make_float4(0, 0, 0, 0);
    w.x = step(w.x, g.x, m1.x, m2.x); w.y = step(w.y, g.y, m1.y, m2.y);
    w.z = step(w.z, g.z, m1.z, m2.z); w.w = step(w.w, g.w, m1.w, m2.w);
    if (a.master) *reinterpret_cast<float4*>(a.master + e) = w;
    if (a.state1) *reinterpret_cast<float4*>(a.state1 + e) = m1;
    if (a.state2) *reinterpret_cast<float4*>(a.state2 + e) = m2;
    for (int k = 0; k < P; ++k) {
      int p = me + k;
      if (p >= P) p -= P;
      Quad<PT>::store(pt.recv[p] + (base + e) * sizeof(PT), w);
    }
  }
  for (size_t e = nquad * 4 + gtid; e < owned; e += gsz) {
    float g = 0.f;
    for (int p = 0; p < P; ++p) g += Quad<GT>::load1(pt.send[p] + (base + e) * sizeof(GT));
    g *= a.grad_scale;
    float w = a.master ? a.master[e] : Quad<PT>::load1(pt.recv[me] + (base + e) * sizeof(PT));
    float m1 = a.state1 ? a.state1[e] : 0.f, m2 = a.state2 ? a.state2[e] : 0.f;
    w = step(w, g, m1, m2);
    if (a.master) a.master[e] = w;
    if (a.state1) a.state1[e] = m1;
    if (a.state2) a.state2[e] = m2;
    for (int p = 0; p < P; ++p) Quad<PT>::store1(pt.recv[p] + (base + e) * sizeof(PT), w);
  }
  comm_sync(dc, pt, t, 1, true);
}

cudaError_t launch_fused_update(const DevComm& dc, DType grad_dt, DType param_dt, unsigned long long grad_off,
                                unsigned long long param_off, size_t owned, const FusedUpdateArgs& a, int channels,
                                cudaStream_t s) {
  if (grad_dt == DType::F32 && param_dt == DType::F32)
    k_fused_update<float, float><<<channels, kCommThreads, 0, s>>>(dc, grad_off, param_off, owned, a);
  else if (grad_dt == DType::F32 && param_dt == DType::BF16)
    k_fused_update<float, __nv_bfloat16><<<channels, kCommThreads, 0, s>>>(dc, grad_off, param_off, owned, a);
  else if (grad_dt == DType::BF16 && param_dt == DType::BF16)
    k_fused_update<__nv_bfloat16, __nv_bfloat16><<<channels, kCommThreads, 0, s>>>(dc, grad_off, param_off, owned, a);
  else if (grad_dt == DType::BF16 && param_dt == DType::F32)
    k_fused_update<__nv_bfloat16, float><<<channels, kCommThreads, 0, s>>>(dc, grad_off, param_off, owned, a);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace mlslb
