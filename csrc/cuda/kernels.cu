// Peer-memory collective kernels for sm_100a (NVLink 5 / NVSwitch).
//
// These replace every MPI call site of the reference (MPI_Iallreduce / Ireduce_scatter_block /
// Iallgather(v) / Ibcast / Ireduce / Ialltoall(v) / Igather / Iscatter / Barrier in reference src/comm_ep.cpp:768-1378
// and eplib/cqueue.c:1930-2098).  One kernel = one collective: the handshake, the data movement out of / into the
// peers' buffers and the scale epilogue all happen inside it; channels (CTAs) play the role of the reference's
// endpoints (a message is split across them).
#include "cuda/kernels.hpp"

#include <cstdlib>
#include <type_traits>

namespace mlslb {

// ------------------------------------------------------------------------------------------------------------
// K8: barrier
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_barrier(DevComm dc) {
  __shared__ PeerTable pt;
  unsigned long long t = comm_begin(dc, pt, 0, 0, NoAux());
  comm_sync(dc, pt, t, 1, false);
}

cudaError_t launch_barrier(const DevComm& dc, cudaStream_t s) {
  k_barrier<<<1, 32, 0, s>>>(dc);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// K1: all-reduce, "fused two-shot": CTA c of rank r owns sub-slice (r, c) of the message.  It pulls that sub-slice
// from every peer's send buffer, reduces in registers, applies the scale and pushes the result into every peer's
// receive buffer.  Reads and writes of one 16-byte chunk are done by the same thread, so in-place operation needs
// no extra barrier: 2 handshakes per collective, NVLink busy in both directions for the whole kernel.
// ------------------------------------------------------------------------------------------------------------
// NVLS flavour (kNvls): when every member passed the SAME slab offsets (symmetric buffers) the pull + reduce of a
// 16-byte chunk is a single `multimem.ld_reduce` (the NVSwitch adds the N copies) and the push to all peers a single
// `multimem.st` (the switch replicates it): 1 load + 1 store per chunk instead of N + N, and each GPU's links carry
// ~S(1+1/N) bytes per direction instead of 2S(N-1)/N.
template <typename T> struct Multimem;   // ld_reduce(.add) / st of one 16-byte vector through a multicast address
template <> struct Multimem<float> {
  __device__ __forceinline__ static uint4 ld_reduce_add(const void* p) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
  }
  __device__ __forceinline__ static void st(void* p, const uint4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
};
template <> struct Multimem<__nv_bfloat16> {
  __device__ __forceinline__ static uint4 ld_reduce_add(const void* p) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
  }
  __device__ __forceinline__ static void st(void* p, const uint4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
};
template <> struct Multimem<__half> {
  __device__ __forceinline__ static uint4 ld_reduce_add(const void* p) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
  }
  __device__ __forceinline__ static void st(void* p, const uint4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f16x2 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
};
template <typename T> struct HasMultimem { static constexpr bool value = false; };
template <> struct HasMultimem<float> { static constexpr bool value = true; };
template <> struct HasMultimem<__nv_bfloat16> { static constexpr bool value = true; };
template <> struct HasMultimem<__half> { static constexpr bool value = true; };

// Work distribution ("pass-major"): the message is cut into passes of P * grid * U 16-byte vectors; inside a pass rank r
// owns the r-th P-th (the final, shorter pass is split evenly too).  All ranks therefore walk the message front to
// back TOGETHER: at any moment the whole group touches one ~25 MiB window of the source and of the destination instead
// of P slices spread over the entire message (a 1 GiB single launch in slice-major order ran at 579 GB/s bus bandwidth
// on 8 GPUs against 767 GB/s for 256 MiB - every GPU's TLB then has to cover 2 GiB of peer / multicast mappings).
// Software pipeline: the loads of pass p+1 are issued before the stores of pass p, so a thread always has loads in
// flight while its stores drain (two register sets, the loop is unrolled by two so both are statically indexed).
template <typename T, typename Op, int U, bool kNvls>
struct ArPass {
  using VT = VecTraits<T>;
  using Acc = typename VT::Acc;
  static constexpr int N = VT::N;
  struct Regs {
    Acc acc[U][N];
    uint4 raw[U];
    unsigned valid;   // bit u: slot u holds a vector
  };
  // range of this rank inside pass `p`: vectors [lo, hi)
  __device__ __forceinline__ static void range(size_t nvec, size_t pass_vecs, size_t p, int P, int me, size_t& lo, size_t& hi) {
    const size_t p0 = p * pass_vecs;
    const size_t pv = min(pass_vecs, nvec - p0);
    const size_t per = (pv + P - 1) / P;
    lo = p0 + min(pv, (size_t)me * per);
    hi = p0 + min(pv, (size_t)(me + 1) * per);
  }
  __device__ __forceinline__ static void load(Regs& r, const DevComm& dc, const PeerTable& pt, const char* msrc, size_t lo, size_t hi,
                                              size_t gtid, size_t gsz) {
    const int P = dc.nranks, me = dc.me;
    r.valid = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = lo + gtid + (size_t)u * gsz;
      if (i < hi) {
        r.valid |= 1u << u;
        if constexpr (kNvls) r.raw[u] = Multimem<T>::ld_reduce_add(msrc + i * 16);
        else r.raw[u] = ld16(pt.send[me] + i * 16);
      }
    }
    if constexpr (!kNvls) {
#pragma unroll
      for (int u = 0; u < U; ++u) VT::unpack(r.raw[u], r.acc[u]);
      for (int q = 1; q < P; ++q) {
        int p = me + q;
        if (p >= P) p -= P;
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const size_t i = lo + gtid + (size_t)u * gsz;
          if (i < hi) v[u] = ld16(pt.send[p] + i * 16);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (r.valid >> u & 1) {
            Acc b[N];
            VT::unpack(v[u], b);
#pragma unroll
            for (int k = 0; k < N; ++k) r.acc[u][k] = Op::apply(r.acc[u][k], b[k]);
          }
        }
      }
    }
  }
  __device__ __forceinline__ static void store(Regs& r, const DevComm& dc, const PeerTable& pt, char* mdst, size_t lo, size_t gtid,
                                               size_t gsz, float scale) {
    const int P = dc.nranks, me = dc.me;
    const bool do_scale = scale != 1.0f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (r.valid >> u & 1) {
        const size_t i = lo + gtid + (size_t)u * gsz;
        if constexpr (kNvls) {
          if (do_scale) {
            VT::unpack(r.raw[u], r.acc[u]);
#pragma unroll
            for (int k = 0; k < N; ++k) r.acc[u][k] = VT::scale(r.acc[u][k], scale);
            r.raw[u] = VT::pack(r.acc[u]);
          }
          Multimem<T>::st(mdst + i * 16, r.raw[u]);
        } else {
          if (do_scale) {
#pragma unroll
            for (int k = 0; k < N; ++k) r.acc[u][k] = VT::scale(r.acc[u][k], scale);
          }
          const uint4 o = VT::pack(r.acc[u]);
          for (int q = 0; q < P; ++q) {
            int p = me + q;
            if (p >= P) p -= P;
            st16(pt.recv[p] + i * 16, o);
          }
        }
      }
    }
  }
};

// the pass loop over vectors [v0, v1) of the message, run by a sub-grid of `gsz` threads (this thread: `gtid`)
template <typename T, typename Op, int U, bool kNvls>
__device__ __forceinline__ void ar_run_passes(const DevComm& dc, const PeerTable& pt, const char* msrc, char* mdst, size_t v0, size_t v1,
                                              size_t gtid, size_t gsz, float scale) {
  using AP = ArPass<T, Op, U, kNvls>;
  const int P = dc.nranks, me = dc.me;
  const size_t nvec = v1 - v0;
  const size_t pass_vecs = (size_t)P * gsz * U;
  const size_t npass = (nvec + pass_vecs - 1) / pass_vecs;
  typename AP::Regs ra, rb;
  size_t lo_a = 0, hi_a = 0, lo_b = 0, hi_b = 0;
  if (npass) {
    AP::range(nvec, pass_vecs, 0, P, me, lo_a, hi_a);
    AP::load(ra, dc, pt, msrc, v0 + lo_a, v0 + hi_a, gtid, gsz);
  }
  for (size_t p = 0; p < npass; p += 2) {
    if (p + 1 < npass) {
      AP::range(nvec, pass_vecs, p + 1, P, me, lo_b, hi_b);
      AP::load(rb, dc, pt, msrc, v0 + lo_b, v0 + hi_b, gtid, gsz);
    }
    AP::store(ra, dc, pt, mdst, v0 + lo_a, gtid, gsz, scale);
    if (p + 1 < npass) {
      if (p + 2 < npass) {
        AP::range(nvec, pass_vecs, p + 2, P, me, lo_a, hi_a);
        AP::load(ra, dc, pt, msrc, v0 + lo_a, v0 + hi_a, gtid, gsz);
      }
      AP::store(rb, dc, pt, mdst, v0 + lo_b, gtid, gsz, scale);
    }
  }
}

template <typename T, typename Op, int U, bool kNvls>
__global__ void __launch_bounds__(kCommThreads) k_allreduce(DevComm dc, unsigned long long send_off,
                                                            unsigned long long recv_off, size_t count, float scale,
                                                            int p2p_cta, float p2p_frac) {
  using VT = VecTraits<T>;
  using Acc = typename VT::Acc;
  constexpr int N = VT::N;
  __shared__ PeerTable pt;
  __shared__ int s_aligned;
  __shared__ int s_symmetric;
  const unsigned long long t = comm_begin(dc, pt, send_off, recv_off, NoAux());
  const int P = dc.nranks, me = dc.me;
  if (threadIdx.x == 0) {
    unsigned long long bits = 0;
    int sym = 1;
    for (int p = 0; p < P; ++p) {
      bits |= (unsigned long long)pt.send[p] | (unsigned long long)pt.recv[p];
      sym &= (pt.send[p] - dc.slab[p]) == (long long)send_off && (pt.recv[p] - dc.slab[p]) == (long long)recv_off;
    }
    s_aligned = (bits & 15ull) == 0;
    s_symmetric = sym && dc.mc != nullptr && !pt.failed && send_off + count * sizeof(T) <= dc.mc_bytes &&
                  recv_off + count * sizeof(T) <= dc.mc_bytes;
  }
  __syncthreads();
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
  const bool do_scale = scale != 1.0f;
  if (s_aligned) {
    const size_t nvec = count / N;
    bool done = false;
    if constexpr (kNvls && HasMultimem<T>::value) {
      if (s_symmetric) {
        // Hybrid: the multicast path of an 8-GPU NVSwitch domain levels off at ~540 GB/s of link traffic per GPU and
        // direction whatever the grid (measured: 48..148 CTAs, 1..4 vectors per thread all land on 478 GB/s algorithm
        // bandwidth) - below what the links carry for plain peer traffic (770 GB/s).  So the last `p2p_cta` CTAs move the
        // tail `p2p_vecs` vectors of the message with the peer-to-peer two-shot code at the same time.
        const size_t p2p_vecs = p2p_cta > 0 && (unsigned)p2p_cta < gridDim.x ? min(nvec, (size_t)((double)nvec * p2p_frac)) : 0;
        const size_t mc_vecs = nvec - p2p_vecs;
        const unsigned mc_cta = gridDim.x - (p2p_vecs ? p2p_cta : 0);
        if (blockIdx.x < mc_cta) {
          ar_run_passes<T, Op, U, true>(dc, pt, dc.mc + send_off, dc.mc + recv_off, 0, mc_vecs,
                                        (size_t)blockIdx.x * blockDim.x + threadIdx.x, (size_t)mc_cta * blockDim.x, scale);
        } else {
          ar_run_passes<T, Op, U, false>(dc, pt, nullptr, nullptr, mc_vecs, nvec,
                                         (size_t)(blockIdx.x - mc_cta) * blockDim.x + threadIdx.x, (size_t)p2p_cta * blockDim.x, scale);
        }
        done = true;
      }
    }
    if (!done) ar_run_passes<T, Op, U, false>(dc, pt, nullptr, nullptr, 0, nvec, gtid, gsz, scale);
    // tail elements (less than one vector): the last rank's first threads, plain peer accesses
    if (me == P - 1) {
      for (size_t i = nvec * N + gtid; i < count; i += gsz) {
        Acc a = VT::load1(pt.send[0] + i * sizeof(T));
        for (int p = 1; p < P; ++p) a = Op::apply(a, VT::load1(pt.send[p] + i * sizeof(T)));
        if (do_scale) a = VT::scale(a, scale);
        for (int p = 0; p < P; ++p) VT::store1(pt.recv[p] + i * sizeof(T), a);
      }
    }
  } else {
    // unaligned views: element-wise, slice-major (rare: sub-views of foreign tensors are staged to aligned scratch)
    size_t per = (count + P - 1) / P;
    const size_t lo = min(count, (size_t)me * per), hi = min(count, lo + per);
    for (size_t i = lo + gtid; i < hi; i += gsz) {
      Acc a = VT::load1(pt.send[0] + i * sizeof(T));
      for (int p = 1; p < P; ++p) a = Op::apply(a, VT::load1(pt.send[p] + i * sizeof(T)));
      if (do_scale) a = VT::scale(a, scale);
      for (int p = 0; p < P; ++p) VT::store1(pt.recv[p] + i * sizeof(T), a);
    }
  }
  comm_sync(dc, pt, t, 1, true);
}

// ------------------------------------------------------------------------------------------------------------
// K1, latency path ("LL"): one CTA; thread i owns bytes [8i, 8i+8) of the message.  It stores {d0, flag, d1, flag}
// (16 bytes, each (data, flag) half is one atomic 8-byte unit) into slot [parity][me][i] of EVERY member's arena, then
// polls the P slots of its own arena until both flags carry the current ticket, reduces in fixed rank order and
// writes the result.  No opening handshake, no fence, no closing handshake: the critical path is one NVLink one-way
// trip.  Two parities suffice: a rank can enter collective t+2 only after it has received every peer's t+1 data,
// which a peer sends only after it finished reading t.  Send/recv may be ANY device-accessible pointers (the data
// goes through registers), so small foreign tensors need no staging either.
// ------------------------------------------------------------------------------------------------------------
template <typename T, typename Op>
__global__ void __launch_bounds__(1024) k_allreduce_ll(DevComm dc, const char* send, char* recv, size_t bytes, float scale) {
  using VT = VecTraits<T>;
  using Acc = typename VT::Acc;
  constexpr int N = VT::N, H = N / 2 > 0 ? N / 2 : 1;          // elements per 8 bytes
  __shared__ unsigned long long s_ticket;
  if (threadIdx.x == 0) {
    unsigned long long* seq = reinterpret_cast<unsigned long long*>(dc.slab[dc.me] + dc.seq_off);   // channel 0
    s_ticket = *seq + 1;
    *seq = s_ticket;
  }
  __syncthreads();
  const unsigned flag = (unsigned)s_ticket;
  const int P = dc.nranks, me = dc.me;
  const size_t n8 = (bytes + 7) / 8;
  const size_t par = (size_t)(s_ticket & 1ull) * kMaxDevRanks * kLLSlotBytes;
  const bool al_in = ((unsigned long long)send & 7ull) == 0, al_out = ((unsigned long long)recv & 7ull) == 0;
  for (size_t i = threadIdx.x; i < n8; i += blockDim.x) {
    // my 8 bytes (zero padded at the tail; byte by byte when the user's view is not 8-byte aligned)
    unsigned d0 = 0, d1 = 0;
    if (al_in && i * 8 + 8 <= bytes) {
      const uint2 v = *reinterpret_cast<const uint2*>(send + i * 8);
      d0 = v.x;
      d1 = v.y;
    } else {
      unsigned char tmp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (size_t b = i * 8; b < bytes && b < i * 8 + 8; ++b) tmp[b - i * 8] = (unsigned char)send[b];
      d0 = tmp[0] | (tmp[1] << 8) | (tmp[2] << 16) | ((unsigned)tmp[3] << 24);
      d1 = tmp[4] | (tmp[5] << 8) | (tmp[6] << 16) | ((unsigned)tmp[7] << 24);
    }
    for (int q = 0; q < P; ++q) {
      int p = me + q;
      if (p >= P) p -= P;
      char* slot = dc.slab[p] + dc.ll_off + par + (size_t)me * kLLSlotBytes + i * 16;
      asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(slot), "r"(d0), "r"(flag), "r"(d1), "r"(flag) : "memory");
    }
    // collect
    Acc acc[N];
    bool first = true, ok = true;
    unsigned long long t0 = 0;
    for (int p = 0; p < P && ok; ++p) {
      const char* slot = dc.slab[me] + dc.ll_off + par + (size_t)p * kLLSlotBytes + i * 16;
      uint4 v;
      unsigned spins = 0;
      for (;;) {
        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(slot) : "memory");
        if (v.y == flag && v.w == flag) break;
        if ((++spins & 0x3ff) == 0) {
          if (!t0) t0 = globaltimer_ns();
          if (*(volatile int*)dc.err != 0 || (dc.timeout_ns && globaltimer_ns() - t0 > dc.timeout_ns)) {
            *(volatile int*)dc.err = 1000 + dc.me;
            ok = false;
            break;
          }
        }
      }
      Acc b[N];
      VT::unpack(make_uint4(v.x, v.z, 0u, 0u), b);
      if (first) {
#pragma unroll
        for (int k = 0; k < H; ++k) acc[k] = b[k];
        first = false;
      } else {
#pragma unroll
        for (int k = 0; k < H; ++k) acc[k] = Op::apply(acc[k], b[k]);
      }
    }
    if (!ok) break;
    if (scale != 1.0f) {
#pragma unroll
      for (int k = 0; k < H; ++k) acc[k] = VT::scale(acc[k], scale);
    }
#pragma unroll
    for (int k = H; k < N; ++k) acc[k] = acc[0];
    const uint4 r = VT::pack(acc);
    if (al_out && i * 8 + 8 <= bytes) {
      *reinterpret_cast<uint2*>(recv + i * 8) = make_uint2(r.x, r.y);
    } else {
      const unsigned w[2] = {r.x, r.y};
      for (size_t b = i * 8; b < bytes && b < i * 8 + 8; ++b) recv[b] = (char)((w[(b - i * 8) >> 2] >> (((b - i * 8) & 3) * 8)) & 0xff);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// K1, mid sizes (8 KiB .. 1 MiB): the same flag-in-data idea on a fixed grid of kMidCtas CTAs.  Messages travel as 16-byte
// {d0, flag, d1, flag} stores into a per-row arena (two parities), receivers spin on the data itself: no opening
// handshake, no fence, no closing handshake - against 3 handshake words + fence.sys + flag round trip of the bandwidth
// kernel (12.9 us at 64 KiB, 21.8 us at 1 MiB on 8 GPUs in round 1).
//   one-shot  ((P-1) * bytes small): every rank pushes its whole vector to every peer, each reduces all P copies in rank
//             order (bit-identical results everywhere): ONE NVLink one-way trip.
//   two-shot  (larger): unit j of slice q goes to rank q only; the owner reduces its slice in rank order, scales, and
//             pushes the result to every peer: two trips, 2 * 2 * bytes * (P-1)/P per direction instead of 2 * bytes * (P-1).
// The flag is the launch number of this kernel on the row (all kMidCtas CTAs count every launch, so one value per
// launch); parity = flag & 1.  Safe for the same reason as k_allreduce_ll: a rank can only start launch t+2 after it
// received every peer's t+1 data, which a peer sends only after its launch t has completely finished reading.
// send / recv may be any device-accessible pointers at any alignment (the data goes through registers).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint2 ll_load8(const char* p, size_t off, size_t bytes, bool aligned) {
  if (aligned && off + 8 <= bytes) return *reinterpret_cast<const uint2*>(p + off);
  unsigned char tmp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t b = off; b < bytes && b < off + 8; ++b) tmp[b - off] = (unsigned char)p[b];
  return make_uint2(tmp[0] | (tmp[1] << 8) | (tmp[2] << 16) | ((unsigned)tmp[3] << 24),
                    tmp[4] | (tmp[5] << 8) | (tmp[6] << 16) | ((unsigned)tmp[7] << 24));
}
__device__ __forceinline__ void ll_store8(char* p, size_t off, size_t bytes, bool aligned, uint2 v) {
  if (aligned && off + 8 <= bytes) {
    *reinterpret_cast<uint2*>(p + off) = v;
    return;
  }
  const unsigned w[2] = {v.x, v.y};
  for (size_t b = off; b < bytes && b < off + 8; ++b) p[b] = (char)((w[(b - off) >> 2] >> (((b - off) & 3) * 8)) & 0xff);
}
__device__ __forceinline__ void ll_push(char* slot, uint2 d, unsigned flag) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(slot), "r"(d.x), "r"(flag), "r"(d.y), "r"(flag) : "memory");
}
// spin until both halves of the slot carry `flag`; false on watchdog / poison
__device__ __forceinline__ bool ll_poll(const DevComm& dc, const char* slot, unsigned flag, uint2& d) {
  uint4 v;
  unsigned spins = 0;
  unsigned long long t0 = 0;
  for (;;) {
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(slot) : "memory");
    if (v.y == flag && v.w == flag) break;
    if ((++spins & 0x3ff) == 0) {
      if (!t0) t0 = globaltimer_ns();
      if (*(volatile int*)dc.err != 0) return false;
      if (dc.timeout_ns && globaltimer_ns() - t0 > dc.timeout_ns) {
        volatile int* e = (volatile int*)dc.err;
        const unsigned long long off = (unsigned long long)(slot - dc.slab[dc.me]);
        e[1] = (int)blockIdx.x;
        e[2] = (int)threadIdx.x;
        e[3] = (int)(off & 0xffffffffull);
        e[4] = (int)(off >> 32);
        e[5] = (int)v.y;
        e[6] = (int)v.w;
        __threadfence_system();
        e[0] = 1000 + dc.me;
        return false;
      }
    }
  }
  d = make_uint2(v.x, v.z);
  return true;
}

template <typename T, typename Op, bool kTwoShot>
__global__ void __launch_bounds__(kMidThreads) k_allreduce_mid(DevComm dc, const char* send, char* recv, size_t bytes, float scale) {
  using VT = VecTraits<T>;
  using Acc = typename VT::Acc;
  constexpr int N = VT::N, H = N / 2 > 0 ? N / 2 : 1;          // elements per 8 bytes
  __shared__ unsigned long long s_ticket;
  if (threadIdx.x == 0) {
    unsigned long long* seq = reinterpret_cast<unsigned long long*>(dc.slab[dc.me] + dc.mid_seq_off) + blockIdx.x;
    s_ticket = *seq + 1;
    *seq = s_ticket;
  }
  __syncthreads();
  const unsigned flag = (unsigned)s_ticket;
  const int P = dc.nranks, me = dc.me;
  const size_t n8 = (bytes + 7) / 8;
  const size_t par = (size_t)(s_ticket & 1ull) * kMidParityBytes;
  const bool al_in = ((unsigned long long)send & 7ull) == 0, al_out = ((unsigned long long)recv & 7ull) == 0;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
  auto to_acc = [](uint2 d, Acc* a) {
    Acc b[N];
    VT::unpack(make_uint4(d.x, d.y, 0u, 0u), b);
#pragma unroll
    for (int k = 0; k < H; ++k) a[k] = b[k];
  };
  auto from_acc = [](Acc* a) {
#pragma unroll
    for (int k = H; k < N; ++k) a[k] = a[0];
    const uint4 r = VT::pack(a);
    return make_uint2(r.x, r.y);
  };
  if constexpr (!kTwoShot) {
    // arena of rank x, this parity: slot [source p][unit i]
    for (size_t i = gtid; i < n8; i += gsz) {
      const uint2 mine = ll_load8(send, i * 8, bytes, al_in);
      for (int q = 1; q < P; ++q) {
        int p = me + q;
        if (p >= P) p -= P;
        ll_push(dc.slab[p] + dc.mid_off + par + ((size_t)me * n8 + i) * 16, mine, flag);
      }
    }
    for (size_t i = gtid; i < n8; i += gsz) {
      Acc acc[N];
      bool ok = true;
      for (int p = 0; p < P && ok; ++p) {
        uint2 d;
        if (p == me) d = ll_load8(send, i * 8, bytes, al_in);
        else ok = ll_poll(dc, dc.slab[me] + dc.mid_off + par + ((size_t)p * n8 + i) * 16, flag, d);
        Acc b[N];
        to_acc(d, b);
        if (p == 0) {
#pragma unroll
          for (int k = 0; k < H; ++k) acc[k] = b[k];
        } else {
#pragma unroll
          for (int k = 0; k < H; ++k) acc[k] = Op::apply(acc[k], b[k]);
        }
      }
      if (!ok) return;
      if (scale != 1.0f) {
#pragma unroll
        for (int k = 0; k < H; ++k) acc[k] = VT::scale(acc[k], scale);
      }
      ll_store8(recv, i * 8, bytes, al_out, from_acc(acc));
    }
  } else {
    const size_t per = (n8 + P - 1) / P;                         // units per slice
    auto len = [&](int r) { size_t lo = (size_t)r * per; return lo >= n8 ? (size_t)0 : min(per, n8 - lo); };
    const size_t rs_base = par, ag_base = par + (size_t)P * per * 16;
    const size_t my_len = len(me);
    // phase 1: unit j of slice q -> rank q's reduce area, slot [source me][j]
    for (size_t j = gtid; j < per; j += gsz) {
      for (int q = 1; q < P; ++q) {
        int p = me + q;
        if (p >= P) p -= P;
        if (j < len(p)) ll_push(dc.slab[p] + dc.mid_off + rs_base + ((size_t)me * per + j) * 16, ll_load8(send, ((size_t)p * per + j) * 8, bytes, al_in), flag);
      }
    }
    // phase 2: reduce my slice in rank order, scale, publish it to every peer's gather area, slot [owner me][j]
    for (size_t j = gtid; j < my_len; j += gsz) {
      Acc acc[N];
      bool ok = true;
      for (int p = 0; p < P && ok; ++p) {
        uint2 d;
        if (p == me) d = ll_load8(send, ((size_t)me * per + j) * 8, bytes, al_in);
        else ok = ll_poll(dc, dc.slab[me] + dc.mid_off + rs_base + ((size_t)p * per + j) * 16, flag, d);
        Acc b[N];
        to_acc(d, b);
        if (p == 0) {
#pragma unroll
          for (int k = 0; k < H; ++k) acc[k] = b[k];
        } else {
#pragma unroll
          for (int k = 0; k < H; ++k) acc[k] = Op::apply(acc[k], b[k]);
        }
      }
      if (!ok) return;
      if (scale != 1.0f) {
#pragma unroll
        for (int k = 0; k < H; ++k) acc[k] = VT::scale(acc[k], scale);
      }
      const uint2 r = from_acc(acc);
      for (int q = 1; q < P; ++q) {
        int p = me + q;
        if (p >= P) p -= P;
        ll_push(dc.slab[p] + dc.mid_off + ag_base + ((size_t)me * per + j) * 16, r, flag);
      }
      ll_store8(recv, ((size_t)me * per + j) * 8, bytes, al_out, r);
    }
    // phase 3: collect the other slices
    for (size_t j = gtid; j < per; j += gsz) {
      for (int q = 1; q < P; ++q) {
        int p = me + q;
        if (p >= P) p -= P;
        if (j < len(p)) {
          uint2 d;
          if (!ll_poll(dc, dc.slab[me] + dc.mid_off + ag_base + ((size_t)p * per + j) * 16, flag, d)) return;
          ll_store8(recv, ((size_t)p * per + j) * 8, bytes, al_out, d);
        }
      }
    }
  }
}

template <typename T, typename Op>
static cudaError_t launch_mid_t(const DevComm& dc, const void* send, void* recv, size_t count, float scale, bool two_shot,
                                int ctas, cudaStream_t s) {
  const size_t bytes = count * sizeof(T);
  if (two_shot) k_allreduce_mid<T, Op, true><<<ctas, kMidThreads, 0, s>>>(dc, (const char*)send, (char*)recv, bytes, scale);
  else k_allreduce_mid<T, Op, false><<<ctas, kMidThreads, 0, s>>>(dc, (const char*)send, (char*)recv, bytes, scale);
  return cudaGetLastError();
}

template <typename T, typename Op>
static cudaError_t launch_ll_t(const DevComm& dc, const void* send, void* recv, size_t count, float scale, cudaStream_t s) {
  const size_t bytes = count * sizeof(T);
  const size_t n8 = (bytes + 7) / 8;
  int threads = (int)std::min<size_t>(1024, std::max<size_t>(32, (n8 + 31) / 32 * 32));
  k_allreduce_ll<T, Op><<<1, threads, 0, s>>>(dc, (const char*)send, (char*)recv, bytes, scale);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// K2 / K5: pull-reduce.  out[i] = scale * op_p send_p[base + i] for i < count, written to MY recv buffer only.
// ReduceScatter: every rank active with base = me * count.  Reduce: only the root is active (base 0); the
// others just keep their send buffers alive until the closing handshake.
// ------------------------------------------------------------------------------------------------------------
// NVLS flavour (kNvls): when every member published the SAME send offset, the P loads of a vector collapse into one
// `multimem.ld_reduce` - the switch adds the copies, the shard arrives reduced (inbound bytes / (P-1)).
template <typename T, typename Op, int U, bool kNvls>
__global__ void __launch_bounds__(kCommThreads) k_reduce_pull(DevComm dc, unsigned long long send_off,
                                                              unsigned long long recv_off, size_t base_elems,
                                                              size_t count, float scale, int active) {
  using VT = VecTraits<T>;
  using Acc = typename VT::Acc;
  constexpr int N = VT::N;
  __shared__ PeerTable pt;
  __shared__ int s_aligned;
  __shared__ int s_symmetric;
  const unsigned long long t = comm_begin(dc, pt, send_off, recv_off, NoAux());
  const int P = dc.nranks, me = dc.me;
  if (active) {
    if (threadIdx.x == 0) {
      unsigned long long bits = (unsigned long long)pt.recv[me] | (unsigned long long)(base_elems * sizeof(T));
      int sym = 1;
      for (int p = 0; p < P; ++p) {
        bits |= (unsigned long long)pt.send[p];
        sym &= (pt.send[p] - dc.slab[p]) == (long long)send_off;
      }
      s_aligned = (bits & 15ull) == 0;
      s_symmetric = sym && dc.mc != nullptr && !pt.failed && send_off + (base_elems + count) * sizeof(T) <= dc.mc_bytes;
    }
    __syncthreads();
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
    const bool do_scale = scale != 1.0f;
    char* out = pt.recv[me];
    size_t nvec = s_aligned ? count / N : 0;
    if constexpr (kNvls && HasMultimem<T>::value) {
      if (s_aligned && s_symmetric) {
        const char* msrc = dc.mc + send_off + base_elems * sizeof(T);
        for (size_t b0 = gtid; b0 < nvec; b0 += gsz * U) {
          uint4 v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const size_t i = b0 + (size_t)u * gsz;
            if (i < nvec) v[u] = Multimem<T>::ld_reduce_add(msrc + i * 16);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const size_t i = b0 + (size_t)u * gsz;
            if (i < nvec) {
              if (do_scale) {
                Acc a[N];
                VT::unpack(v[u], a);
#pragma unroll
                for (int k = 0; k < N; ++k) a[k] = VT::scale(a[k], scale);
                v[u] = VT::pack(a);
              }
              st16(out + i * 16, v[u]);
            }
          }
        }
        // the vector body is done: leave only the scalar tail to the generic code below
        for (size_t i = nvec * N + gtid; i < count; i += gsz) {
          Acc a = VT::load1(pt.send[me] + (base_elems + i) * sizeof(T));
          for (int q = 1; q < P; ++q) {
            int p = me + q;
            if (p >= P) p -= P;
            a = Op::apply(a, VT::load1(pt.send[p] + (base_elems + i) * sizeof(T)));
          }
          if (do_scale) a = VT::scale(a, scale);
          VT::store1(out + i * sizeof(T), a);
        }
        nvec = 0;
        count = 0;
      }
    }
    for (size_t b0 = gtid; b0 < nvec; b0 += gsz * U) {
      Acc acc[U][N];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t i = b0 + (size_t)u * gsz;
        if (i < nvec) VT::unpack(ld16(pt.send[me] + (base_elems + i * N) * sizeof(T)), acc[u]);
      }
      for (int q = 1; q < P; ++q) {
        int p = me + q;
        if (p >= P) p -= P;
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const size_t i = b0 + (size_t)u * gsz;
          if (i < nvec) v[u] = ld16(pt.send[p] + (base_elems + i * N) * sizeof(T));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const size_t i = b0 + (size_t)u * gsz;
          if (i < nvec) {
            Acc b[N];
            VT::unpack(v[u], b);
#pragma unroll
            for (int k = 0; k < N; ++k) acc[u][k] = Op::apply(acc[u][k], b[k]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t i = b0 + (size_t)u * gsz;
        if (i < nvec) {
          if (do_scale) {
#pragma unroll
            for (int k = 0; k < N; ++k) acc[u][k] = VT::scale(acc[u][k], scale);
          }
          st16(out + i * N * sizeof(T), VT::pack(acc[u]));
        }
      }
    }
    for (size_t i = nvec * N + gtid; i < count; i += gsz) {
      Acc a = VT::load1(pt.send[me] + (base_elems + i) * sizeof(T));
      for (int q = 1; q < P; ++q) {
        int p = me + q;
        if (p >= P) p -= P;
        a = Op::apply(a, VT::load1(pt.send[p] + (base_elems + i) * sizeof(T)));
      }
      if (do_scale) a = VT::scale(a, scale);
      VT::store1(out + i * sizeof(T), a);
    }
  }
  comm_sync(dc, pt, t, 1, false);
}

// ------------------------------------------------------------------------------------------------------------
// K3/K4/K6/K7/K9: pull-copy.  Every gather-like collective is "copy these byte ranges out of the peers' send
// buffers into my receive buffer"; only local memory is written, so no write fence is needed before the closing
// handshake.  For the *v collectives the source offset inside the peer's buffer comes from the peer (aux word).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mc_st16(void* p, const uint4& v) {   // raw 16 bytes to every member through the switch
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// NVLS push (bcast / all-gather): possible when every member published the same buffer offset, which all members
// see identically after the opening handshake - so they all take the same branch.  The pusher streams its bytes with
// `multimem.st` (one store, the switch replicates it: the root of a bcast sends the message ONCE instead of P-1 times),
// the closing handshake carries the fence.  Returns false when the pull path has to run.
__device__ __forceinline__ bool pull_copy_try_mc(const DevComm& dc, const PeerTable& pt, const CopyPlan& plan,
                                                 unsigned long long send_off, unsigned long long recv_off, size_t gtid, size_t gsz) {
  if (plan.mc_mode == 0 || dc.mc == nullptr || pt.failed) return false;
  const int P = dc.nranks, me = dc.me;
  // bcast publishes the one buffer as `send`; all-gather needs the receive buffers symmetric (the shard may live anywhere)
  bool sym = true;
  for (int p = 0; p < P; ++p) {
    if (plan.mc_mode == 1) sym &= (pt.send[p] - dc.slab[p]) == (long long)send_off;
    else sym &= (pt.recv[p] - dc.slab[p]) == (long long)recv_off;
  }
  const unsigned long long dst_off = plan.mc_mode == 1 ? send_off : recv_off + (unsigned long long)me * plan.mc_bytes;
  if ((plan.mc_mode == 1 ? send_off + plan.mc_bytes : recv_off + (unsigned long long)P * plan.mc_bytes) > dc.mc_bytes) return false;
  const char* src = plan.mc_mode == 1 ? pt.send[me] : pt.send[me];
  if (!sym || ((dst_off | (unsigned long long)src | (plan.mc_mode == 2 ? plan.mc_bytes : 0ull)) & 15ull) != 0) return false;
  if (plan.mc_mode == 1 && me != plan.mc_root) return true;          // receivers only wait for the closing handshake
  const size_t nvec = plan.mc_bytes / 16;
  char* mdst = dc.mc + dst_off;
  constexpr int U = 4;
  for (size_t b0 = gtid; b0 < nvec; b0 += gsz * U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = b0 + (size_t)u * gsz;
      if (i < nvec) v[u] = ld16(src + i * 16);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = b0 + (size_t)u * gsz;
      if (i < nvec) mc_st16(mdst + i * 16, v[u]);
    }
  }
  for (size_t i = nvec * 16 + gtid; i < plan.mc_bytes; i += gsz)     // bcast tail: plain stores to every member
    for (int p = 0; p < P; ++p) (dc.slab[p] + dst_off)[i] = src[i];
  return true;
}

__global__ void __launch_bounds__(kCommThreads) k_pull_copy(DevComm dc, CopyPlan plan, unsigned long long send_off,
                                                            unsigned long long recv_off) {
  __shared__ PeerTable pt;
  struct AuxOut {
    const CopyPlan* pl;
    __device__ __forceinline__ unsigned long long operator()(int p) const { return pl->aux_out[p]; }
  } auxfn{&plan};
  const unsigned long long t = comm_begin(dc, pt, send_off, recv_off, auxfn);
  const int me = dc.me;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
  if (pull_copy_try_mc(dc, pt, plan, send_off, recv_off, gtid, gsz)) {
    comm_sync(dc, pt, t, 1, true);
    return;
  }
  constexpr int U = 4;
  // pairs_concurrent: CTA c works on segment c mod nseg only, with the CTAs that share that residue as its sub-grid
  const bool deal = plan.pairs_concurrent && (int)gridDim.x >= plan.nseg && plan.nseg > 1;
  const int my_sg = deal ? (int)(blockIdx.x % plan.nseg) : -1;
  const size_t sub_ctas = deal ? (gridDim.x - my_sg + plan.nseg - 1) / plan.nseg : gridDim.x;
  const size_t stid = deal ? (size_t)(blockIdx.x / plan.nseg) * blockDim.x + threadIdx.x : gtid;
  const size_t ssz = deal ? sub_ctas * blockDim.x : gsz;
  for (int sgi = 0; sgi < plan.nseg; ++sgi) {
    if (deal && sgi != my_sg) continue;
    // rotate the segment order by rank so the peers are not all hammering the same source at the same time
    int sidx = sgi + me;
    while (sidx >= plan.nseg) sidx -= plan.nseg;
    const CopySeg sg = plan.seg[sidx];
    const char* src = pt.send[sg.peer] + sg.src_off + (sg.use_aux ? pt.aux[sg.peer] * (unsigned long long)plan.elem_size : 0ull);
    char* dst = pt.recv[me] + sg.dst_off;
    if (src == dst || sg.bytes == 0) continue;
    const bool aligned = ((((unsigned long long)src) | ((unsigned long long)dst)) & 15ull) == 0;
    size_t done = 0;
    const size_t gtid = stid, gsz = ssz;       // this segment's sub-grid (shadows the whole-grid indices)
    if (sg.rows > 1) {
      // strided rectangle (activation blocks pulled straight out of / into unpacked tensors)
      const bool al2 = aligned && ((sg.row_bytes | sg.src_stride | sg.dst_stride) & 15ull) == 0;
      if (al2) {
        const size_t rv = sg.row_bytes / 16, nvec = sg.rows * rv;
        for (size_t b0 = gtid; b0 < nvec; b0 += gsz * U) {
          uint4 v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const size_t i = b0 + (size_t)u * gsz;
            if (i < nvec) {
              const size_t row = i / rv, c = i - row * rv;
              v[u] = ld16(src + row * sg.src_stride + c * 16);
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const size_t i = b0 + (size_t)u * gsz;
            if (i < nvec) {
              const size_t row = i / rv, c = i - row * rv;
              st16(dst + row * sg.dst_stride + c * 16, v[u]);
            }
          }
        }
      } else {
        for (size_t i = gtid; i < sg.bytes; i += gsz) {
          const size_t row = i / sg.row_bytes, c = i - row * sg.row_bytes;
          dst[row * sg.dst_stride + c] = src[row * sg.src_stride + c];
        }
      }
      continue;
    }
    if (aligned) {
      const size_t nvec = sg.bytes / 16;
      for (size_t b0 = gtid; b0 < nvec; b0 += gsz * U) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const size_t i = b0 + (size_t)u * gsz;
          if (i < nvec) v[u] = ld16(src + i * 16);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const size_t i = b0 + (size_t)u * gsz;
          if (i < nvec) st16(dst + i * 16, v[u]);
        }
      }
      done = nvec * 16;
    }
    for (size_t i = done + gtid; i < sg.bytes; i += gsz) dst[i] = src[i];
  }
  comm_sync(dc, pt, t, 1, false);
}

// Large gather-like collectives: the same plan, moved by the copy engine instead of by threads.  One elected thread per
// CTA drives a ring of kBulkSlots x 16 KiB shared-memory pieces: `cp.async.bulk` peer global -> shared (mbarrier
// complete_tx), then shared -> local global (bulk_group); the store of a piece is issued kBulkLag steps after its load,
// so ~half the ring is always inbound over NVLink and half outbound to HBM, and no registers or LSU slots are spent
// on the payload (SASS: UBLKCP).  Pieces are dealt round-robin to the CTAs.  Segments that are not 16-byte aligned on
// both sides, and the sub-16-byte tails, are copied by the CTA's threads as before.
constexpr int kBulkThreads = 128;
constexpr int kBulkSlots = 8, kBulkLag = 4;
constexpr unsigned kBulkPiece = 16384;
constexpr size_t kBulkSmem = (size_t)kBulkSlots * kBulkPiece + 1024;

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(kBulkThreads, 1) k_pull_copy_bulk(DevComm dc, CopyPlan plan, unsigned long long send_off,
                                                                    unsigned long long recv_off) {
  extern __shared__ unsigned char bulk_smem_raw[];
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(bulk_smem_raw) + 127) & ~(uintptr_t)127);
  __shared__ PeerTable pt;
  __shared__ unsigned long long s_bar[kBulkSlots];
  struct AuxOut {
    const CopyPlan* pl;
    __device__ __forceinline__ unsigned long long operator()(int p) const { return pl->aux_out[p]; }
  } auxfn{&plan};
  const unsigned long long t = comm_begin(dc, pt, send_off, recv_off, auxfn);
  const int me = dc.me;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
  if (pull_copy_try_mc(dc, pt, plan, send_off, recv_off, gtid, gsz)) {
    comm_sync(dc, pt, t, 1, true);
    return;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kBulkSlots; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&s_bar[i])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // resolve the segments once (rotated by rank, see k_pull_copy)
  const char* seg_src[kMaxDevRanks];
  char* seg_dst[kMaxDevRanks];
  unsigned long long seg_bulk[kMaxDevRanks];       // bytes the copy engine moves (0: the threads copy the segment)
  long long pieces_before[kMaxDevRanks + 1];
  pieces_before[0] = 0;
  for (int sgi = 0; sgi < plan.nseg; ++sgi) {
    int sidx = sgi + me;
    while (sidx >= plan.nseg) sidx -= plan.nseg;
    const CopySeg sg = plan.seg[sidx];
    const char* src = pt.send[sg.peer] + sg.src_off + (sg.use_aux ? pt.aux[sg.peer] * (unsigned long long)plan.elem_size : 0ull);
    char* dst = pt.recv[me] + sg.dst_off;
    const bool skip = src == dst || sg.bytes == 0;
    const bool aligned = ((((unsigned long long)src) | ((unsigned long long)dst)) & 15ull) == 0;
    seg_src[sgi] = src;
    seg_dst[sgi] = dst;
    seg_bulk[sgi] = skip || !aligned ? 0ull : (sg.bytes & ~15ull);
    pieces_before[sgi + 1] = pieces_before[sgi] + (long long)((seg_bulk[sgi] + kBulkPiece - 1) / kBulkPiece);
  }
  const long long total = pieces_before[plan.nseg];
  if (threadIdx.x == 0) {
    const long long G = gridDim.x, c = blockIdx.x;
    const long long mine = total > c ? (total - c + G - 1) / G : 0;      // pieces c, c+G, c+2G, ...
    auto locate = [&](long long k, const char*& src, char*& dst, unsigned& bytes) {
      const long long gp = c + k * G;
      int sgi = 0;
      while (gp >= pieces_before[sgi + 1]) ++sgi;
      const unsigned long long off = (unsigned long long)(gp - pieces_before[sgi]) * kBulkPiece;
      src = seg_src[sgi] + off;
      dst = seg_dst[sgi] + off;
      const unsigned long long left = seg_bulk[sgi] - off;
      bytes = (unsigned)(left < kBulkPiece ? left : kBulkPiece);
    };
    for (long long i = 0; i < mine + kBulkLag; ++i) {
      if (i >= kBulkSlots && i - kBulkSlots < mine)
        asm volatile("cp.async.bulk.wait_group %0;" ::"n"(kBulkSlots - 1 - kBulkLag) : "memory");   // slot's last store done
      if (i < mine) {
        const char* src;
        char* dst;
        unsigned bytes;
        locate(i, src, dst, bytes);
        const int slot = (int)(i % kBulkSlots);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&s_bar[slot])), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_addr(ring + (size_t)slot * kBulkPiece)),
                     "l"(src), "r"(bytes), "r"(smem_addr(&s_bar[slot]))
                     : "memory");
      }
      const long long j = i - kBulkLag;
      if (j >= 0 && j < mine) {
        const char* src;
        char* dst;
        unsigned bytes;
        locate(j, src, dst, bytes);
        const int slot = (int)(j % kBulkSlots);
        const unsigned parity = (unsigned)((j / kBulkSlots) & 1);
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra DONE_%=;\n"
            "bra WAIT_%=;\n"
            "DONE_%=:\n"
            "}\n" ::"r"(smem_addr(&s_bar[slot])), "r"(parity)
            : "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr(ring + (size_t)slot * kBulkPiece)),
                     "r"(bytes)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    // the other threads: unaligned segments and sub-vector tails, grid-strided over all CTAs' helper threads
    const size_t hid = (size_t)blockIdx.x * (blockDim.x - 1) + (threadIdx.x - 1), hsz = (size_t)gridDim.x * (blockDim.x - 1);
    for (int sgi = 0; sgi < plan.nseg; ++sgi) {
      int sidx = sgi + me;
      while (sidx >= plan.nseg) sidx -= plan.nseg;
      const unsigned long long bytes = plan.seg[sidx].bytes;
      if (seg_src[sgi] == seg_dst[sgi]) continue;
      for (size_t i = seg_bulk[sgi] + hid; i < bytes; i += hsz) seg_dst[sgi][i] = seg_src[sgi][i];
    }
  }
  comm_sync(dc, pt, t, 1, false);
}

cudaError_t init_kernel_attributes() {
  return cudaFuncSetAttribute(k_pull_copy_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBulkSmem);
}

cudaError_t launch_pull_copy(const DevComm& dc, const CopyPlan& plan, unsigned long long send_off,
                             unsigned long long recv_off, int channels, bool bulk, cudaStream_t s) {
  if (bulk) k_pull_copy_bulk<<<channels, kBulkThreads, kBulkSmem, s>>>(dc, plan, send_off, recv_off);
  else k_pull_copy<<<channels, kCommThreads, 0, s>>>(dc, plan, send_off, recv_off);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------------------
template <typename T, typename Op>
static cudaError_t launch_ar_t(const DevComm& dc, unsigned long long so, unsigned long long ro, size_t count,
                               float scale, int channels, int unroll, int p2p_cta, float p2p_frac, cudaStream_t s) {
  // vectors each thread moves per pass: all in flight at once (U), two register sets (software pipeline) -> U <= 4
  const size_t per_thread = (count * sizeof(T) / (size_t)dc.nranks) / ((size_t)channels * kCommThreads * 16);
  constexpr bool kCanNvls = HasMultimem<T>::value && std::is_same<Op, OpSum>::value;
  int U = unroll > 0 ? unroll : (per_thread <= 1 ? 1 : (per_thread <= 3 ? 2 : 4));
  if (kCanNvls && dc.mc != nullptr) {
    if (U >= 4) k_allreduce<T, Op, 4, kCanNvls><<<channels, kCommThreads, 0, s>>>(dc, so, ro, count, scale, p2p_cta, p2p_frac);
    else if (U >= 2) k_allreduce<T, Op, 2, kCanNvls><<<channels, kCommThreads, 0, s>>>(dc, so, ro, count, scale, p2p_cta, p2p_frac);
    else k_allreduce<T, Op, 1, kCanNvls><<<channels, kCommThreads, 0, s>>>(dc, so, ro, count, scale, p2p_cta, p2p_frac);
    return cudaGetLastError();
  }
  if (U >= 4) k_allreduce<T, Op, 4, false><<<channels, kCommThreads, 0, s>>>(dc, so, ro, count, scale, 0, 0.f);
  else if (U >= 2) k_allreduce<T, Op, 2, false><<<channels, kCommThreads, 0, s>>>(dc, so, ro, count, scale, 0, 0.f);
  else k_allreduce<T, Op, 1, false><<<channels, kCommThreads, 0, s>>>(dc, so, ro, count, scale, 0, 0.f);
  return cudaGetLastError();
}
template <typename T, typename Op>
static cudaError_t launch_rp_t(const DevComm& dc, unsigned long long so, unsigned long long ro, size_t base,
                               size_t count, float scale, bool active, int channels, cudaStream_t s) {
  const size_t per_thread = (count * sizeof(T)) / ((size_t)channels * kCommThreads * 16);
  constexpr bool kCanNvls = HasMultimem<T>::value && std::is_same<Op, OpSum>::value;
  if (kCanNvls && dc.mc != nullptr) {
    if (per_thread <= 1)
      k_reduce_pull<T, Op, 1, kCanNvls><<<channels, kCommThreads, 0, s>>>(dc, so, ro, base, count, scale, active ? 1 : 0);
    else
      k_reduce_pull<T, Op, 4, kCanNvls><<<channels, kCommThreads, 0, s>>>(dc, so, ro, base, count, scale, active ? 1 : 0);
    return cudaGetLastError();
  }
  if (per_thread <= 1)
    k_reduce_pull<T, Op, 1, false><<<channels, kCommThreads, 0, s>>>(dc, so, ro, base, count, scale, active ? 1 : 0);
  else
    k_reduce_pull<T, Op, 4, false><<<channels, kCommThreads, 0, s>>>(dc, so, ro, base, count, scale, active ? 1 : 0);
  return cudaGetLastError();
}

#define MLSLB_DISPATCH_T_OP(DT, OP, CALL)                                                     \
  switch (DT) {                                                                               \
    case DType::F32: MLSLB_DISPATCH_OP(float, OP, CALL) break;                                \
    case DType::F64: MLSLB_DISPATCH_OP(double, OP, CALL) break;                               \
    case DType::I32: MLSLB_DISPATCH_OP(int, OP, CALL) break;                                  \
    case DType::U8:                                                                           \
    case DType::F8E4M3: MLSLB_DISPATCH_OP(unsigned char, OP, CALL) break;                     \
    case DType::BF16: MLSLB_DISPATCH_OP(__nv_bfloat16, OP, CALL) break;                       \
    case DType::F16: MLSLB_DISPATCH_OP(__half, OP, CALL) break;                               \
  }
#define MLSLB_DISPATCH_OP(T, OP, CALL)                                                        \
  switch (OP) {                                                                               \
    case RedOp::SUM: { using TT = T; using OO = OpSum; return CALL; }                         \
    case RedOp::MIN: { using TT = T; using OO = OpMin; return CALL; }                         \
    case RedOp::MAX: { using TT = T; using OO = OpMax; return CALL; }                         \
  }

cudaError_t launch_allreduce(const DevComm& dc, DType dt, RedOp op, unsigned long long send_off,
                             unsigned long long recv_off, size_t count, float scale, int channels, int unroll, int p2p_cta,
                             float p2p_frac, cudaStream_t s) {
  MLSLB_DISPATCH_T_OP(dt, op, (launch_ar_t<TT, OO>(dc, send_off, recv_off, count, scale, channels, unroll, p2p_cta, p2p_frac, s)))
  return cudaErrorInvalidValue;
}

cudaError_t launch_allreduce_ll(const DevComm& dc, DType dt, RedOp op, const void* send, void* recv, size_t count,
                                float scale, cudaStream_t s) {
  MLSLB_DISPATCH_T_OP(dt, op, (launch_ll_t<TT, OO>(dc, send, recv, count, scale, s)))
  return cudaErrorInvalidValue;
}

cudaError_t launch_allreduce_mid(const DevComm& dc, DType dt, RedOp op, const void* send, void* recv, size_t count,
                                 float scale, bool two_shot, int ctas, cudaStream_t s) {
  MLSLB_DISPATCH_T_OP(dt, op, (launch_mid_t<TT, OO>(dc, send, recv, count, scale, two_shot, ctas, s)))
  return cudaErrorInvalidValue;
}

cudaError_t launch_reduce_pull(const DevComm& dc, DType dt, RedOp op, unsigned long long send_off,
                               unsigned long long recv_off, size_t base, size_t count, float scale, bool active,
                               int channels, cudaStream_t s) {
  MLSLB_DISPATCH_T_OP(dt, op, (launch_rp_t<TT, OO>(dc, send_off, recv_off, base, count, scale, active, channels, s)))
  return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------------------------
// K13: activation pack / unpack.  A block is an (mbCount x fmCount*fmSize) rectangle; one minibatch row of it is
// contiguous on both sides, so each row is a straight vector copy.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_blocks(PackPlan plan, size_t local_fm_count, int es, const char* src,
                                                     char* dst, int unpack) {
  const int b = blockIdx.y;
  if (b >= plan.n) return;
  const BlockDesc k = plan.b[b];
  const size_t row_bytes = k.fm_cnt * k.fm_size * (size_t)es;
  const size_t total = k.mb_cnt * row_bytes;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
  const size_t local0 = (k.mb_off * local_fm_count + k.fm_off) * k.fm_size * (size_t)es;
  const size_t local_stride = local_fm_count * k.fm_size * (size_t)es;
  const size_t comm0 = k.buf_off * (size_t)es;
  const bool vec = ((row_bytes | local0 | local_stride | comm0 | (size_t)src | (size_t)dst) & 15) == 0;
  if (vec) {
    const size_t rv = row_bytes / 16, nv = total / 16;
    for (size_t i = gtid; i < nv; i += gsz) {
      const size_t mb = i / rv, c = i - mb * rv;
      const size_t lo = local0 + mb * local_stride + c * 16, co = comm0 + mb * row_bytes + c * 16;
      if (!unpack) *(uint4*)(dst + co) = *(const uint4*)(src + lo);
      else *(uint4*)(dst + lo) = *(const uint4*)(src + co);
    }
  } else {
    for (size_t i = gtid; i < total; i += gsz) {
      const size_t mb = i / row_bytes, c = i - mb * row_bytes;
      const size_t lo = local0 + mb * local_stride + c, co = comm0 + mb * row_bytes + c;
      if (!unpack) dst[co] = src[lo];
      else dst[lo] = src[co];
    }
  }
}

cudaError_t launch_pack_blocks(const PackPlan& plan, size_t local_fm_count, int elem_size, const void* src, void* dst,
                               bool unpack, size_t max_block_elems, cudaStream_t s) {
  size_t bytes = max_block_elems * (size_t)elem_size;
  int gx = (int)std::min<size_t>(148, std::max<size_t>(1, bytes / (256 * 64)));
  dim3 grid(gx, plan.n);
  k_pack_blocks<<<grid, 256, 0, s>>>(plan, local_fm_count, elem_size, (const char*)src, (char*)dst, unpack ? 1 : 0);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// local scale (single-rank groups keep the "result = scale * sum" contract without any peer traffic)
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_scale(T* buf, size_t count, float scale) {
  using VT = VecTraits<T>;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
    VT::store1(buf + i, VT::scale(VT::load1(buf + i), scale));
}

cudaError_t launch_scale(DType dt, void* buf, size_t count, float scale, cudaStream_t s) {
  int grid = (int)std::min<size_t>(296, std::max<size_t>(1, count / 1024));
  switch (dt) {
    case DType::F32: k_scale<float><<<grid, 256, 0, s>>>((float*)buf, count, scale); break;
    case DType::F64: k_scale<double><<<grid, 256, 0, s>>>((double*)buf, count, scale); break;
    case DType::BF16: k_scale<__nv_bfloat16><<<grid, 256, 0, s>>>((__nv_bfloat16*)buf, count, scale); break;
    case DType::F16: k_scale<__half><<<grid, 256, 0, s>>>((__half*)buf, count, scale); break;
    default: break;
  }
  return cudaGetLastError();
}

}  // namespace mlslb

namespace mlslb {
// dst = scale * src (single-rank groups: the whole "collective" is this one kernel)
template <typename T>
__global__ void __launch_bounds__(256) k_scale_copy(T* __restrict__ dst, const T* __restrict__ src, size_t count, float scale) {
  using VT = VecTraits<T>;
  using Acc = typename VT::Acc;
  constexpr int N = VT::N;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
  const bool vec = ((((size_t)dst) | ((size_t)src)) & 15) == 0;
  const size_t nvec = vec ? count / N : 0;
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  // four independent 16-byte loads in flight per thread before the first store: a single load per iteration leaves
  // ~32 KB outstanding per SM, short of what HBM3e needs to stay busy (82 % of the copy peak measured that way)
  constexpr int U = 4;
  const size_t nblk = nvec / (U * gsz) * (U * gsz);
  for (size_t base = gtid; base < nblk; base += U * gsz) {
    uint4 v[U];
#pragma unroll
    for (int j = 0; j < U; ++j) v[j] = __ldcs(s4 + base + (size_t)j * gsz);
#pragma unroll
    for (int j = 0; j < U; ++j) {
      Acc a[N];
      VT::unpack(v[j], a);
#pragma unroll
      for (int k = 0; k < N; ++k) a[k] = VT::scale(a[k], scale);
      __stcs(d4 + base + (size_t)j * gsz, VT::pack(a));
    }
  }
  for (size_t i = nblk + gtid; i < nvec; i += gsz) {
    Acc a[N];
    VT::unpack(__ldcs(s4 + i), a);
#pragma unroll
    for (int k = 0; k < N; ++k) a[k] = VT::scale(a[k], scale);
    __stcs(d4 + i, VT::pack(a));
  }
  for (size_t i = nvec * N + gtid; i < count; i += gsz) VT::store1(dst + i, VT::scale(VT::load1(src + i), scale));
}

cudaError_t launch_scale_copy(DType dt, void* dst, const void* src, size_t count, float scale, cudaStream_t s) {
  int grid = (int)std::min<size_t>(148 * 8, std::max<size_t>(1, count / 2048));
  switch (dt) {
    case DType::F32: k_scale_copy<float><<<grid, 256, 0, s>>>((float*)dst, (const float*)src, count, scale); break;
    case DType::F64: k_scale_copy<double><<<grid, 256, 0, s>>>((double*)dst, (const double*)src, count, scale); break;
    case DType::BF16: k_scale_copy<__nv_bfloat16><<<grid, 256, 0, s>>>((__nv_bfloat16*)dst, (const __nv_bfloat16*)src, count, scale); break;
    case DType::F16: k_scale_copy<__half><<<grid, 256, 0, s>>>((__half*)dst, (const __half*)src, count, scale); break;
    case DType::I32: k_scale_copy<int><<<grid, 256, 0, s>>>((int*)dst, (const int*)src, count, 1.0f); break;
    default: k_scale_copy<unsigned char><<<grid, 256, 0, s>>>((unsigned char*)dst, (const unsigned char*)src, count, 1.0f); break;
  }
  return cudaGetLastError();
}
}  // namespace mlslb
