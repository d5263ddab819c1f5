// Device-side communication primitives shared by every collective kernel (sm_100a).
//
// Model: each rank owns one "slab" of device memory that every peer of the NVSwitch domain has mapped (CUDA IPC,
// or plain pointers for in-process ranks).  The first bytes of a slab hold signal pads: for every (group row, lane)
// an array PadSlot[channel][peer].  A collective kernel is launched with the same grid on every member of the group;
// CTA c of rank r only ever talks to CTA c of the peers ("channel" c - the GPU analogue of the reference's endpoint
// c, eplib/), so no grid-wide synchronisation exists anywhere:
//
//   begin()  : ticket = ++seq[channel]            (device resident -> kernels are CUDA-graph replayable)
//              write {send_off, recv_off, aux, flag=4*ticket} into every peer's pad slot [c][me]  (st.release.sys)
//              wait for every peer's slot in MY pad  (ld.acquire.sys), pick up their buffer offsets
//   ... move / reduce data straight out of / into the peers' buffers over NVLink ...
//   sync(k)  : flag = 4*ticket + k handshake between the same CTAs (k = 1..3), release/acquire at .sys scope
//
// Flags only ever grow, so nothing is reset between collectives; a new group on a recycled row starts from zeroed
// pads (host side).  Every spin has a %globaltimer deadline: on expiry the kernel records an error word in
// host-mapped memory and stops waiting, so a missing peer can never wedge the GPU (fail-fast).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "core/common.hpp"

namespace mlslb {

struct alignas(32) PadSlot {
  unsigned long long flag, a, b, aux;
};

struct DevComm {
  int nranks;                 // group size
  int me;                     // my index in the group
  unsigned pad_off;           // byte offset of this (row, lane) pad array inside every slab
  unsigned seq_off;           // byte offset of my private per-channel ticket counters
  unsigned long long timeout_ns;
  int* err;                   // host-mapped error word (0 = ok)
  char* slab[kMaxDevRanks];   // slab base of every member (group order) in MY address space
  char* mc;                   // multicast mapping of the slabs (NVLS) or nullptr
  unsigned long long mc_bytes;   // the multicast object covers slab offsets [0, mc_bytes) (chunks the heap grew by are not in it)
  unsigned ll_off;            // byte offset of this row's low-latency arena inside every slab (0 = none)
  unsigned mid_off;           // byte offset of this row's mid-size flag-in-data arena (0 = none)
  unsigned mid_seq_off;       // byte offset of my launch counters of the mid kernel (one per CTA)
};

struct PeerTable {            // lives in shared memory
  char* send[kMaxDevRanks];
  char* recv[kMaxDevRanks];
  unsigned long long aux[kMaxDevRanks];
  unsigned long long ticket;
  int failed;
};

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// 16-byte accesses that stay out of L1 (peer lines are only ever cached in the reader's L1, never its L2 -
// keeping them out makes every load observe the owner's L2/HBM).
__device__ __forceinline__ uint4 ld16(const void* p) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st16(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

__device__ __forceinline__ PadSlot* pad_slot(const DevComm& dc, int owner, int channel, int writer) {
  return reinterpret_cast<PadSlot*>(dc.slab[owner] + dc.pad_off) + channel * kMaxDevRanks + writer;
}

// A spin costs one relaxed (volatile-class) system-scope load per iteration; the deadline check only starts after
// the first miss.
template <typename Pred>
__device__ __forceinline__ bool spin_on(const DevComm& dc, const unsigned long long* word, Pred ok) {
  if (ok(ld_relaxed_sys(word))) return true;
  unsigned long long t0 = globaltimer_ns();
  unsigned spins = 0;
  while (!ok(ld_relaxed_sys(word))) {
    if ((++spins & 0x3ff) == 0) {
      if (*(volatile int*)dc.err != 0) return false;               // another CTA / the host already gave up
      if (dc.timeout_ns && globaltimer_ns() - t0 > dc.timeout_ns) {
        // post-mortem for the host: which word (slab offset), what it held, which CTA / waiting thread
        volatile int* e = (volatile int*)dc.err;
        const unsigned long long off = (unsigned long long)((const char*)word - dc.slab[dc.me]);
        const unsigned long long seen = ld_relaxed_sys(word);
        e[1] = (int)blockIdx.x;
        e[2] = (int)threadIdx.x;
        e[3] = (int)(off & 0xffffffffull);
        e[4] = (int)(off >> 32);
        e[5] = (int)(seen & 0xffffffffull);
        e[6] = (int)(seen >> 32);
        __threadfence_system();
        e[0] = 1000 + dc.me;
        return false;
      }
    }
  }
  return true;
}

constexpr unsigned long long kTagShift = 40;                       // payload: 40 bits (1 TiB of slab / 2^40 elements)
constexpr unsigned long long kPayloadMask = (1ull << kTagShift) - 1;

// Opening handshake.  Every word a rank publishes carries the low 24 bits of the ticket in its top bits, so each
// word validates itself: three relaxed stores per peer, no release fence (= no NVLink round trip) on the critical
// path.  The data the peers are about to read was produced by EARLIER kernels of the publishing rank, i.e. it is
// already performed in that GPU's L2 before this kernel could start.  `aux_for_peer(p)` is a per-destination word
// (the *v collectives publish their send offsets with it).  Returns the ticket; fills `pt`.
template <typename AuxFn>
__device__ __forceinline__ unsigned long long comm_begin(const DevComm& dc, PeerTable& pt, unsigned long long send_off,
                                                         unsigned long long recv_off, AuxFn aux_for_peer) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    unsigned long long* seq = reinterpret_cast<unsigned long long*>(dc.slab[dc.me] + dc.seq_off) + blockIdx.x;
    unsigned long long t = *seq + 1;
    *seq = t;
    pt.ticket = t;
    pt.failed = 0;
  }
  __syncthreads();
  const unsigned long long t = pt.ticket;
  if (tid < dc.nranks) {
    const unsigned long long tag = (t & 0xffffffull) << kTagShift;
    PadSlot* remote = pad_slot(dc, tid, blockIdx.x, dc.me);
    st_relaxed_sys(&remote->a, tag | send_off);
    st_relaxed_sys(&remote->b, tag | recv_off);
    st_relaxed_sys(&remote->aux, tag | (aux_for_peer(tid) & kPayloadMask));
    PadSlot* local = pad_slot(dc, dc.me, blockIdx.x, tid);
    auto fresh = [tag](unsigned long long w) { return (w & ~kPayloadMask) == tag; };
    const bool ok = spin_on(dc, &local->a, fresh) && spin_on(dc, &local->b, fresh) && spin_on(dc, &local->aux, fresh);
    if (ok) {
      pt.send[tid] = dc.slab[tid] + (ld_relaxed_sys(&local->a) & kPayloadMask);
      pt.recv[tid] = dc.slab[tid] + (ld_relaxed_sys(&local->b) & kPayloadMask);
      pt.aux[tid] = ld_relaxed_sys(&local->aux) & kPayloadMask;
    } else {   // peer never showed up: keep every address valid (results are garbage, the host reports the error)
      pt.failed = 1;
      pt.send[tid] = dc.slab[dc.me] + send_off;
      pt.recv[tid] = dc.slab[dc.me] + recv_off;
      pt.aux[tid] = 0;
    }
  }
  __syncthreads();
  return t;
}

struct NoAux {
  __device__ __forceinline__ unsigned long long operator()(int) const { return 0ull; }
};

// Handshake k (1..3) between the same channel of every member.  `published_writes`: this CTA stored data a peer
// will read after the handshake (into peer memory, or into local memory a peer pulls from), so those writes must be
// performed system-wide before the flag can be seen: bar.sync (CTA-scope happens-before) + fence.sys by the
// signalling thread (cumulative) + relaxed flag store.
__device__ __forceinline__ void comm_sync(const DevComm& dc, PeerTable& pt, unsigned long long t, int k,
                                          bool published_writes) {
  __syncthreads();
  const int tid = threadIdx.x;
  if (tid < dc.nranks) {
    if (published_writes) __threadfence_system();
    const unsigned long long want = 4 * t + k;
    st_relaxed_sys(&pad_slot(dc, tid, blockIdx.x, dc.me)->flag, want);
    if (!spin_on(dc, &pad_slot(dc, dc.me, blockIdx.x, tid)->flag, [want](unsigned long long w) { return w >= want; }))
      pt.failed = 1;
    // order the data reads that follow after the flag observation
    __threadfence_system();
  }
  __syncthreads();
}

// ---- element math ----------------------------------------------------------------------------------------------
struct OpSum {
  template <typename A> __device__ __forceinline__ static A apply(A a, A b) { return a + b; }
};
struct OpMin {
  template <typename A> __device__ __forceinline__ static A apply(A a, A b) { return a < b ? a : b; }
};
struct OpMax {
  template <typename A> __device__ __forceinline__ static A apply(A a, A b) { return a > b ? a : b; }
};

// VecTraits<T>: how a 16-byte vector of T is unpacked to accumulators and packed back.
template <typename T> struct VecTraits;
template <> struct VecTraits<float> {
  using Acc = float;
  static constexpr int N = 4;
  __device__ __forceinline__ static void unpack(const uint4& v, Acc* a) {
    a[0] = __uint_as_float(v.x); a[1] = __uint_as_float(v.y); a[2] = __uint_as_float(v.z); a[3] = __uint_as_float(v.w);
  }
  __device__ __forceinline__ static uint4 pack(const Acc* a) {
    return make_uint4(__float_as_uint(a[0]), __float_as_uint(a[1]), __float_as_uint(a[2]), __float_as_uint(a[3]));
  }
  __device__ __forceinline__ static Acc load1(const void* p) { return *(const float*)p; }
  __device__ __forceinline__ static void store1(void* p, Acc a) { *(float*)p = a; }
  __device__ __forceinline__ static Acc scale(Acc a, float s) { return a * s; }
};
template <> struct VecTraits<double> {
  using Acc = double;
  static constexpr int N = 2;
  __device__ __forceinline__ static void unpack(const uint4& v, Acc* a) {
    a[0] = __hiloint2double((int)v.y, (int)v.x); a[1] = __hiloint2double((int)v.w, (int)v.z);
  }
  __device__ __forceinline__ static uint4 pack(const Acc* a) {
    return make_uint4((unsigned)__double2loint(a[0]), (unsigned)__double2hiint(a[0]), (unsigned)__double2loint(a[1]),
                      (unsigned)__double2hiint(a[1]));
  }
  __device__ __forceinline__ static Acc load1(const void* p) { return *(const double*)p; }
  __device__ __forceinline__ static void store1(void* p, Acc a) { *(double*)p = a; }
  __device__ __forceinline__ static Acc scale(Acc a, float s) { return a * (double)s; }
};
template <> struct VecTraits<int> {
  using Acc = int;
  static constexpr int N = 4;
  __device__ __forceinline__ static void unpack(const uint4& v, Acc* a) { a[0] = (int)v.x; a[1] = (int)v.y; a[2] = (int)v.z; a[3] = (int)v.w; }
  __device__ __forceinline__ static uint4 pack(const Acc* a) { return make_uint4((unsigned)a[0], (unsigned)a[1], (unsigned)a[2], (unsigned)a[3]); }
  __device__ __forceinline__ static Acc load1(const void* p) { return *(const int*)p; }
  __device__ __forceinline__ static void store1(void* p, Acc a) { *(int*)p = a; }
  __device__ __forceinline__ static Acc scale(Acc a, float) { return a; }
};
template <> struct VecTraits<unsigned char> {
  using Acc = unsigned;   // per-byte lanes kept in 32-bit accumulators, wrap-around on pack
  static constexpr int N = 16;
  __device__ __forceinline__ static void unpack(const uint4& v, Acc* a) {
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (w[i >> 2] >> ((i & 3) * 8)) & 0xffu;
  }
  __device__ __forceinline__ static uint4 pack(const Acc* a) {
    unsigned w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i >> 2] |= (a[i] & 0xffu) << ((i & 3) * 8);
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
  __device__ __forceinline__ static Acc load1(const void* p) { return *(const unsigned char*)p; }
  __device__ __forceinline__ static void store1(void* p, Acc a) { *(unsigned char*)p = (unsigned char)a; }
  __device__ __forceinline__ static Acc scale(Acc a, float) { return a; }
};
template <> struct VecTraits<__nv_bfloat16> {
  using Acc = float;
  static constexpr int N = 8;
  __device__ __forceinline__ static void unpack(const uint4& v, Acc* a) {
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[2 * i] = __uint_as_float(w[i] << 16);
      a[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ static uint4 pack(const Acc* a) {
    unsigned w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(a[2 * i], a[2 * i + 1]);
      w[i] = *reinterpret_cast<unsigned*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
  __device__ __forceinline__ static Acc load1(const void* p) { return __bfloat162float(*(const __nv_bfloat16*)p); }
  __device__ __forceinline__ static void store1(void* p, Acc a) { *(__nv_bfloat16*)p = __float2bfloat16_rn(a); }
  __device__ __forceinline__ static Acc scale(Acc a, float s) { return a * s; }
};
template <> struct VecTraits<__half> {
  using Acc = float;
  static constexpr int N = 8;
  __device__ __forceinline__ static void unpack(const uint4& v, Acc* a) {
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      float2 f = __half22float2(h);
      a[2 * i] = f.x;
      a[2 * i + 1] = f.y;
    }
  }
  __device__ __forceinline__ static uint4 pack(const Acc* a) {
    unsigned w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = __floats2half2_rn(a[2 * i], a[2 * i + 1]);
      w[i] = *reinterpret_cast<unsigned*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
  __device__ __forceinline__ static Acc load1(const void* p) { return __half2float(*(const __half*)p); }
  __device__ __forceinline__ static void store1(void* p, Acc a) { *(__half*)p = __float2half_rn(a); }
  __device__ __forceinline__ static Acc scale(Acc a, float s) { return a * s; }
};

}  // namespace mlslb
