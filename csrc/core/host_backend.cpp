// Host shared-memory backend: every collective of the library executed by CPU threads over POSIX shared memory
// (or plain process memory for in-process virtual ranks).
//
// Role: (1) the "CPU plumbing" configuration of BASELINE.json (mlsl_sample on 2 CPU ranks), (2) makes the whole
// graph/API layer testable without GPUs, (3) reference semantics oracle for the CUDA kernels.  It deliberately
// uses the SAME protocol shape as the device kernels - publish buffer offsets in a per-(group,lane) signal row,
// barrier, pull from the peers' buffers, barrier - where the reference issues MPI_I* calls on endpoint
// communicators (reference src/comm_ep.cpp:768-1378) and stages foreign buffers through the shared heap
// (ReplaceIn/ReplaceOut, src/comm_ep.cpp:363-566): buffers obtained from Environment::Alloc are zero-copy,
// anything else is transparently staged.
#include <sched.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>

#include "heap.hpp"
#include "log.hpp"
#include "numeric.hpp"
#include <dlfcn.h>

#include "quant.hpp"
#include "runtime.hpp"

namespace mlslb {

namespace {

struct alignas(64) HostPub {
  std::atomic<uint64_t> phase;            // 4*ticket + step
  uint64_t send_off, recv_off;            // offsets in the owner's region (absolute pointers in-process)
  uint64_t aux[kMaxHostRanks];            // per-destination send offsets (elements) for the *v collectives
};

constexpr int kPubRows = kMaxGroupRows * 2;

struct HostReqState {
  std::vector<float> residual;            // error-feedback residual of the quantised path (per request)
};

template <typename T>
struct Acc { using type = T; };

inline float load_as_float(const uint16_t* p, bool bf) { return bf ? bf16_to_f32(*p) : f16_to_f32(*p); }

template <typename T>
inline T apply(RedOp op, T a, T b) {
  switch (op) {
    case RedOp::SUM: return (T)(a + b);
    case RedOp::MIN: return a < b ? a : b;
    case RedOp::MAX: return a > b ? a : b;
  }
  return a;
}

// dst[i] = scale * reduce_p src[p][i]   (fixed p order => bitwise identical on every rank)
// Blocked so that every pass is a plain two-array loop the compiler vectorises: a block of the first source is copied
// into a small accumulator, the other sources are folded in one after the other, then the block is written out - dst
// may alias one of the sources (in-place all-reduce), every block is read completely before it is written.
struct SumOp {
  template <typename T> static inline T f(T a, T b) { return (T)(a + b); }
};
struct MinOp {
  template <typename T> static inline T f(T a, T b) { return a < b ? a : b; }
};
struct MaxOp {
  template <typename T> static inline T f(T a, T b) { return a > b ? a : b; }
};

template <typename T, typename Op>
static inline void reduce_blocked(T* dst, const void* const* srcs, size_t nsrc, size_t n, float scale) {
  constexpr size_t B = 8192 / sizeof(T);
  T acc[B];
  const bool do_scale = scale != 1.0f;
  const T sc = (T)scale;
  for (size_t b0 = 0; b0 < n; b0 += B) {
    const size_t m = n - b0 < B ? n - b0 : B;
    const T* __restrict__ s0 = (const T*)srcs[0] + b0;
    for (size_t i = 0; i < m; ++i) acc[i] = s0[i];
    for (size_t p = 1; p < nsrc; ++p) {
      const T* __restrict__ sp = (const T*)srcs[p] + b0;
      for (size_t i = 0; i < m; ++i) acc[i] = Op::f(acc[i], sp[i]);
    }
    if (do_scale)
      for (size_t i = 0; i < m; ++i) acc[i] = (T)(acc[i] * sc);
    T* d = dst + b0;
    for (size_t i = 0; i < m; ++i) d[i] = acc[i];
  }
}

template <typename T>
static inline void reduce_by_op(T* dst, const void* const* srcs, size_t nsrc, size_t n, RedOp op, float scale) {
  switch (op) {
    case RedOp::SUM: reduce_blocked<T, SumOp>(dst, srcs, nsrc, n, scale); break;
    case RedOp::MIN: reduce_blocked<T, MinOp>(dst, srcs, nsrc, n, scale); break;
    case RedOp::MAX: reduce_blocked<T, MaxOp>(dst, srcs, nsrc, n, scale); break;
  }
}

// one entry point per element type, cloned for the vector ISAs of the machine it runs on (resolved at load time)
// (not under ThreadSanitizer: ifunc resolvers run before its runtime is up)
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__) && !defined(__SANITIZE_THREAD__)
#define MLSLB_SIMD_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define MLSLB_SIMD_CLONES
#endif
MLSLB_SIMD_CLONES void reduce_f32(float* d, const void* const* s, size_t ns, size_t n, RedOp op, float sc) { reduce_by_op<float>(d, s, ns, n, op, sc); }
MLSLB_SIMD_CLONES void reduce_f64(double* d, const void* const* s, size_t ns, size_t n, RedOp op, float sc) { reduce_by_op<double>(d, s, ns, n, op, sc); }
MLSLB_SIMD_CLONES void reduce_i32(int32_t* d, const void* const* s, size_t ns, size_t n, RedOp op) { reduce_by_op<int32_t>(d, s, ns, n, op, 1.0f); }
MLSLB_SIMD_CLONES void reduce_u8(uint8_t* d, const void* const* s, size_t ns, size_t n, RedOp op) { reduce_by_op<uint8_t>(d, s, ns, n, op, 1.0f); }

template <typename Op, bool BF>
static inline void reduce_half_blocked(uint16_t* dst, const void* const* srcs, size_t nsrc, size_t n, float scale) {
  constexpr size_t B = 2048;
  float acc[B];
  for (size_t b0 = 0; b0 < n; b0 += B) {
    const size_t m = n - b0 < B ? n - b0 : B;
    const uint16_t* s0 = (const uint16_t*)srcs[0] + b0;
    for (size_t i = 0; i < m; ++i) acc[i] = BF ? bf16_to_f32(s0[i]) : f16_to_f32(s0[i]);
    for (size_t p = 1; p < nsrc; ++p) {
      const uint16_t* sp = (const uint16_t*)srcs[p] + b0;
      for (size_t i = 0; i < m; ++i) acc[i] = Op::f(acc[i], BF ? bf16_to_f32(sp[i]) : f16_to_f32(sp[i]));
    }
    uint16_t* d = dst + b0;
    for (size_t i = 0; i < m; ++i) d[i] = BF ? f32_to_bf16(acc[i] * scale) : f32_to_f16(acc[i] * scale);
  }
}

MLSLB_SIMD_CLONES void reduce_half(uint16_t* dst, const void* const* srcs, size_t nsrc, size_t n, RedOp op, float scale, bool bf) {
  switch (op) {
    case RedOp::SUM:
      bf ? reduce_half_blocked<SumOp, true>(dst, srcs, nsrc, n, scale) : reduce_half_blocked<SumOp, false>(dst, srcs, nsrc, n, scale);
      break;
    case RedOp::MIN:
      bf ? reduce_half_blocked<MinOp, true>(dst, srcs, nsrc, n, scale) : reduce_half_blocked<MinOp, false>(dst, srcs, nsrc, n, scale);
      break;
    case RedOp::MAX:
      bf ? reduce_half_blocked<MaxOp, true>(dst, srcs, nsrc, n, scale) : reduce_half_blocked<MaxOp, false>(dst, srcs, nsrc, n, scale);
      break;
  }
}

void reduce_any(DType dt, void* dst, const std::vector<const void*>& srcs, size_t n, RedOp op, float scale) {
  switch (dt) {
    case DType::F32: reduce_f32((float*)dst, srcs.data(), srcs.size(), n, op, scale); break;
    case DType::F64: reduce_f64((double*)dst, srcs.data(), srcs.size(), n, op, scale); break;
    case DType::U8:
    case DType::F8E4M3: reduce_u8((uint8_t*)dst, srcs.data(), srcs.size(), n, op); break;
    case DType::I32: reduce_i32((int32_t*)dst, srcs.data(), srcs.size(), n, op); break;
    case DType::BF16: reduce_half((uint16_t*)dst, srcs.data(), srcs.size(), n, op, scale, true); break;
    case DType::F16: reduce_half((uint16_t*)dst, srcs.data(), srcs.size(), n, op, scale, false); break;
  }
}

class HostBackend final : public Backend {
 public:
  explicit HostBackend(RankContext* ctx) : ctx_(ctx) {
    Bootstrap* b = ctx->boot.get();
    inproc_ = b->inproc();
    world_ = b->size();
    rank_ = b->rank();
    dir_off_ = round_up(sizeof(HostPub) * kPubRows, 4096);
    pub_bytes_ = dir_off_ + round_up(sizeof(RegionDir), 4096);   // [signal rows | region directory | heap ...]
    if (inproc_) {
      // all virtual ranks share the address space: only the signal rows need a shared region
      region_bytes_ = pub_bytes_;
    } else {
      double gb = ctx->env.heap_size_gb;
      region_bytes_ = pub_bytes_ + (size_t)(gb * 1024.0 * 1024.0 * 1024.0);
    }
    base_.assign(world_, nullptr);
    base_[rank_] = (char*)b->create_region("hheap", region_bytes_);
    b->barrier();
    for (int p = 0; p < world_; ++p)
      if (p != rank_) base_[p] = (char*)b->attach_region(p, "hheap", region_bytes_);
    b->barrier();
    b->seal_regions();
    regions_.assign(world_, {});
    for (int p = 0; p < world_; ++p) regions_[p].push_back(Region{base_[p], region_bytes_});
    if (!inproc_) {
      heaps_.emplace_back(new SlabAllocator());
      heaps_[0]->reset(pub_bytes_, region_bytes_ - pub_bytes_);
      dir(rank_)->bytes[0] = region_bytes_;
      dir(rank_)->count.store(1, std::memory_order_release);
    }
    MLSLB_LOG(LOG_DEBUG, "host backend: world %d inproc %d region %zu bytes", world_, (int)inproc_, region_bytes_);
  }

  ~HostBackend() override {}

  const char* name() const override { return "host"; }

  void* alloc(size_t bytes, size_t align) override {
    if (align == 0) align = 64;
    void* p = nullptr;
    if (inproc_) {
      p = aligned_host_alloc(bytes, align, ctx_->env.thp_threshold_mb << 20);
      MLSLB_ASSERT(p != nullptr, "host alloc of %zu bytes failed", bytes);
      std::lock_guard<std::mutex> g(mu_);
      inproc_live_[p] = bytes;
    } else {
      p = heap_alloc(bytes, align);
    }
    ctx_->ptrcheck.add(p, bytes);
    return p;
  }

  void free(void* p) override {
    if (!p) return;
    ctx_->ptrcheck.remove(p);
    if (inproc_) {
      std::lock_guard<std::mutex> g(mu_);
      auto it = inproc_live_.find(p);
      MLSLB_ASSERT(it != inproc_live_.end(), "Free of a pointer that did not come from Alloc");
      inproc_live_.erase(it);
      ::free(p);
    } else {
      std::lock_guard<std::mutex> g(mu_);
      int k = own_region_of(p, 1);
      MLSLB_ASSERT(k >= 0 && heaps_[k]->free((size_t)((char*)p - regions_[rank_][k].base)),
                   "Free of a pointer that did not come from Alloc");
    }
  }

  bool owns(const void* p, size_t len) const override {
    if (inproc_) return true;   // one address space: every buffer is directly visible to the peers
    std::lock_guard<std::mutex> g(mu_);
    return own_region_of(p, len) >= 0;
  }

  uint64_t heap_offset(const void* p) const override { return to_off(p); }
  void* peer_heap_ptr(int global_rank, uint64_t offset) override { return peer_ptr(global_rank, offset); }

  void prepare(CommRequest& r) override {
    if (!r.backend_state) r.backend_state = new HostReqState();
  }
  void release(CommRequest& r) override {
    delete (HostReqState*)r.backend_state;
    r.backend_state = nullptr;
  }

  void launch(CommRequest& r) override {
    execute(r);
    r.state.store(CommRequest::DONE, std::memory_order_release);
  }
  bool test(CommRequest& r) override { return r.state.load(std::memory_order_acquire) >= CommRequest::DONE; }
  bool peek_done(CommRequest& r) override { return r.state.load(std::memory_order_acquire) != CommRequest::LAUNCHED; }
  void wait(CommRequest& r) override {
    uint64_t spins = 0, t0 = 0;
    while (r.state.load(std::memory_order_acquire) < CommRequest::DONE) {
      if ((++spins & 0xff) == 0) sched_yield();
      if ((spins & 0xffff) == 0) {   // a progress thread that died or a peer that failed must not leave the waiter spinning
        if (ctx_->boot && ctx_->boot->poisoned())
          MLSLB_ASSERT(false, "job poisoned by rank %d", (int)ctx_->boot->poisoned() - 1);
        if (!t0) t0 = now_ns();
        const int wd = ctx_->env.watchdog_sec;
        if (wd > 0 && now_ns() - t0 > (uint64_t)wd * 1000000000ull) {
          if (ctx_->boot) ctx_->boot->poison(ctx_->rank);
          MLSLB_ASSERT(false, "watchdog: %s never completed on the progress thread", opkind_name(r.desc.kind));
        }
      }
    }
  }

  void finalize() override {
    Bootstrap* b = ctx_->boot.get();
    try {
      b->barrier();
    } catch (const std::exception&) {   // poisoned job: still unmap
    }
    for (int p = 0; p < world_; ++p)
      for (size_t k = 0; k < regions_[p].size(); ++k)
        if (regions_[p][k].base) b->release_region(regions_[p][k].base, regions_[p][k].bytes, p == rank_, region_name((int)k));
    regions_.clear();
    base_.clear();
  }

  std::string describe() const override {
    return std::string("host shared-memory backend (") + (inproc_ ? "in-process ranks" : "POSIX shm") + ")";
  }

 private:
  RankContext* ctx_;
  bool inproc_ = false;
  int world_ = 1, rank_ = 0;
  size_t pub_bytes_ = 0, region_bytes_ = 0;
  std::vector<char*> base_;
  // The heap grows on demand like the reference's (eplib/memory.c:396-410: when the mspace is full another shared
  // region, at least twice as large, is registered on the client and on every server; at most 99 of them).  Region 0
  // carries a directory of the region sizes; peers attach a new region the first time an offset points into it.
  // Offsets that travel between ranks are (region index << 48) | byte offset.
  static constexpr int kMaxRegions = 100;
  static constexpr int kRegionShift = 48;
  struct RegionDir {
    std::atomic<uint32_t> count;
    uint32_t pad;
    uint64_t bytes[kMaxRegions];
  };
  struct Region {
    char* base;
    size_t bytes;
  };
  size_t dir_off_ = 0;
  mutable std::vector<std::vector<Region>> regions_;   // [rank][region]: what this process has mapped so far
  std::vector<std::unique_ptr<SlabAllocator>> heaps_;  // my regions
  mutable std::mutex mu_;

  RegionDir* dir(int global_rank) const { return (RegionDir*)(base_[global_rank] + dir_off_); }
  static std::string region_name(int k) { return k == 0 ? std::string("hheap") : "hheap" + std::to_string(k); }
  // index of my region that contains [p, p+len), -1 if none (mu_ held)
  int own_region_of(const void* p, size_t len) const {
    const char* c = (const char*)p;
    const auto& mine = regions_[rank_];
    for (size_t k = 0; k < mine.size(); ++k) {
      const char* lo = mine[k].base + (k == 0 ? pub_bytes_ : 0);
      if (c >= lo && c + len <= mine[k].base + mine[k].bytes) return (int)k;
    }
    return -1;
  }
  void* heap_alloc(size_t bytes, size_t align) {
    std::lock_guard<std::mutex> g(mu_);
    for (size_t k = 0; k < heaps_.size(); ++k) {
      size_t off = heaps_[k]->alloc(bytes, align);
      if (off != SIZE_MAX) return regions_[rank_][k].base + off;
    }
    const int k = (int)heaps_.size();
    size_t in_use = 0, cap = 0;
    for (auto& h : heaps_) in_use += h->bytes_in_use(), cap += h->capacity();
    MLSLB_ASSERT(k < kMaxRegions,
                 "symmetric host heap exhausted (%zu bytes requested, %zu of %zu in use in %d regions): raise "
                 "MLSL_HEAP_SIZE_GB", bytes, in_use, cap, k);
    size_t want = std::max(2 * regions_[rank_].back().bytes, round_up(bytes + align + 4096, (size_t)1 << 20));
    char* base = (char*)ctx_->boot->create_region(region_name(k), want);
    regions_[rank_].push_back(Region{base, want});
    heaps_.emplace_back(new SlabAllocator());
    heaps_[k]->reset(0, want);
    dir(rank_)->bytes[k] = want;
    dir(rank_)->count.store((uint32_t)k + 1, std::memory_order_release);
    MLSLB_LOG(LOG_INFO, "host heap grown: region %d of %.1f MiB (%zu of %zu bytes were in use)", k, want / 1048576.0,
              in_use, cap);
    size_t off = heaps_[k]->alloc(bytes, align);
    MLSLB_ASSERT(off != SIZE_MAX, "allocation of %zu bytes failed in a fresh region of %zu bytes", bytes, want);
    return base + off;
  }
  char* peer_region(int global_rank, int k) const {
    std::lock_guard<std::mutex> g(mu_);
    auto& v = regions_[global_rank];
    while ((int)v.size() <= k) {
      const int kk = (int)v.size();
      RegionDir* d = dir(global_rank);
      MLSLB_ASSERT(d->count.load(std::memory_order_acquire) > (uint32_t)kk, "rank %d published an offset into region %d it never created",
                   global_rank, kk);
      size_t bytes = d->bytes[kk];
      v.push_back(Region{(char*)ctx_->boot->attach_region(global_rank, region_name(kk), bytes), bytes});
    }
    return v[k].base;
  }
  std::map<void*, size_t> inproc_live_;

  HostPub* pub(int global_rank, int prow) const { return (HostPub*)base_[global_rank] + prow; }
  uint64_t to_off(const void* p) const {
    if (inproc_) return (uint64_t)(uintptr_t)p;
    if (!p) return 0;   // barrier, non-root sides of rooted collectives
    const char* c = (const char*)p;
    if (c >= base_[rank_] && c < base_[rank_] + region_bytes_) return (uint64_t)(c - base_[rank_]);
    std::lock_guard<std::mutex> g(mu_);
    int k = own_region_of(p, 1);
    MLSLB_ASSERT(k > 0, "buffer %p is not in the symmetric heap", p);
    return ((uint64_t)k << kRegionShift) | (uint64_t)(c - regions_[rank_][k].base);
  }
  char* peer_ptr(int global_rank, uint64_t off) const {
    if (inproc_) return (char*)(uintptr_t)off;
    const int k = (int)(off >> kRegionShift);
    if (k == 0) return base_[global_rank] + off;
    return peer_region(global_rank, k) + (off & (((uint64_t)1 << kRegionShift) - 1));
  }

  struct Stage {
    void* user = nullptr;
    void* slab = nullptr;
    size_t bytes = 0;
  };

  void wait_all(const ProcessGroup& g, int prow, uint64_t target) {
    uint64_t spins = 0, t0 = 0;
    for (int m : g.members) {
      HostPub* pp = pub(m, prow);
      while (pp->phase.load(std::memory_order_acquire) < target) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if ((++spins & 0x3ff) == 0) {
          sched_yield();
          if (ctx_->boot->poisoned())
            MLSLB_ASSERT(false, "job poisoned by rank %d during a host collective", (int)ctx_->boot->poisoned() - 1);
          if (!t0) t0 = now_ns();
          int wd = ctx_->env.watchdog_sec;
          if (wd > 0 && now_ns() - t0 > (uint64_t)wd * 1000000000ull) {
            ctx_->boot->poison(rank_);
            MLSLB_ASSERT(false, "watchdog: rank %d never reached phase %llu on signal row %d", m,
                         (unsigned long long)target, prow);
          }
        }
      }
    }
  }

  void execute(CommRequest& r);
  void exec_quantized_allreduce(CommRequest& r, const ProcessGroup& g, int prow, char* S, char* R);
  void exec_plugin_allreduce(CommRequest& r, const ProcessGroup& g, int prow, char* S, char* R);
  void load_quant_plugin();
  // user-supplied compression library (Environment::SetQuantizationParams with a lib_path)
  typedef int (*PluginQuant)(void* src, void* dst, size_t count, void* diff, int src_dtype, size_t comp_ratio, int method);
  typedef int (*PluginDequant)(void* src, void* dst, size_t count);
  typedef int (*PluginReduce)(const void* in, void* inout, size_t block_count);
  void* plugin_lib_ = nullptr;
  PluginQuant plugin_quant_ = nullptr;
  PluginDequant plugin_dequant_ = nullptr;
  PluginReduce plugin_reduce_ = nullptr;
};

void HostBackend::execute(CommRequest& r) {
  const CommDesc& d = r.desc;
  ProcessGroup* gp = d.group;
  const size_t dt = dtype_size(d.dtype);
  const size_t n = d.count;
  // ---- single-rank groups: local semantics only --------------------------------------------------------
  if ((!gp || gp->size() <= 1) && d.kind != OpKind::FUSED_UPDATE) {
    switch (d.kind) {
      case OpKind::ALLREDUCE:
      case OpKind::REDUCE:
      case OpKind::REDUCE_SCATTER:
      case OpKind::ALLGATHER:
      case OpKind::GATHER:
      case OpKind::SCATTER:
      case OpKind::ALLTOALL:
        if (r.recv && r.recv != r.send && n) memmove(r.recv, r.send, n * dt);
        if (d.scale != 1.0f && r.recv && (d.kind == OpKind::ALLREDUCE || d.kind == OpKind::REDUCE_SCATTER)) {
          std::vector<const void*> one{r.recv};
          reduce_any(d.dtype, r.recv, one, n, RedOp::SUM, d.scale);
        }
        break;
      case OpKind::ALLGATHERV:
        if (r.recv && r.recv != r.send && n) memmove(r.recv, r.send, n * dt);
        break;
      case OpKind::ALLTOALLV:
      case OpKind::SENDRECV_LIST:
        if (!d.send_counts.empty() && d.send_counts[0])
          memmove((char*)r.recv + d.recv_offsets[0] * dt, (char*)r.send + d.send_offsets[0] * dt, d.send_counts[0] * dt);
        break;
      default: break;
    }
    return;
  }
  const ProcessGroup& g = *gp;
  const int P = g.size();
  const int me = g.idx;
  // single-rank groups have no signal row of their own: they use the reserved last one (only this rank touches it)
  const int prow = (g.row >= 0 ? g.row : kMaxGroupRows - 1) * 2 + r.lane;
  const uint64_t t = r.group_seq;
  HostPub* mine = pub(rank_, prow);

  // ---- stage foreign buffers into the symmetric heap -----------------------------------------------------
  std::vector<Stage> stages;
  auto stage = [&](void* user, size_t bytes, bool copy_in) -> char* {
    if (!user || bytes == 0 || owns(user, bytes)) return (char*)user;
    Stage s;
    s.user = user;
    s.bytes = bytes;
    s.slab = alloc(bytes, 64);
    if (copy_in) memcpy(s.slab, user, bytes);
    stages.push_back(s);
    return (char*)s.slab;
  };
  size_t sbytes = r.send_bytes(), rbytes = r.recv_bytes();
  if (d.kind == OpKind::REDUCE && me != (int)d.root) rbytes = 0;
  if (d.kind == OpKind::GATHER && me != (int)d.root) rbytes = 0;
  if (d.kind == OpKind::SCATTER && me != (int)d.root) sbytes = 0;
  char* R = nullptr;
  char* S = nullptr;
  bool recv_staged = false;
  if (r.recv && rbytes) {
    size_t before = stages.size();
    R = stage(r.recv, rbytes, true);
    recv_staged = stages.size() != before;
  }
  if (r.send && sbytes) {
    char* us = (char*)r.send;
    char* ur = (char*)r.recv;
    if (ur && rbytes && us >= ur && us + sbytes <= ur + rbytes) {
      S = R + (us - ur);             // send aliases the receive buffer (in-place variants)
    } else {
      S = stage(r.send, sbytes, true);
    }
  }

  auto arrive = [&](int step) { mine->phase.store(4 * t + step, std::memory_order_release); };
  auto sync = [&](int step) {
    arrive(step);
    wait_all(g, prow, 4 * t + step);
  };
  mine->send_off = to_off(S);
  mine->recv_off = to_off(R);
  if (d.kind == OpKind::ALLTOALLV || d.kind == OpKind::SENDRECV_LIST)
    for (int p = 0; p < P; ++p) mine->aux[p] = d.send_offsets[p];
  sync(0);

  auto peerS = [&](int p) { return (const char*)peer_ptr(g.members[p], pub(g.members[p], prow)->send_off); };
  auto peerR = [&](int p) { return (const char*)peer_ptr(g.members[p], pub(g.members[p], prow)->recv_off); };

  switch (d.kind) {
    case OpKind::BARRIER: break;
    case OpKind::ALLREDUCE: {
      if (d.compress && d.dtype == DType::F32 && d.rop == RedOp::SUM) {
        if (ctx_->quant.set && !ctx_->quant.lib_path.empty()) exec_plugin_allreduce(r, g, prow, S, R);
        else exec_quantized_allreduce(r, g, prow, S, R);
        break;
      }
      size_t per = ceil_div(n, (size_t)P);
      size_t lo = std::min(n, (size_t)me * per), hi = std::min(n, lo + per);
      std::vector<const void*> srcs(P);
      for (int p = 0; p < P; ++p) srcs[p] = peerS(p) + lo * dt;
      if (hi > lo) reduce_any(d.dtype, R + lo * dt, srcs, hi - lo, d.rop, d.scale);
      sync(1);
      for (int p = 0; p < P; ++p) {
        if (p == me) continue;
        size_t plo = std::min(n, (size_t)p * per), phi = std::min(n, plo + per);
        if (phi > plo) memcpy(R + plo * dt, peerR(p) + plo * dt, (phi - plo) * dt);
      }
      sync(2);
      break;
    }
    case OpKind::REDUCE_SCATTER: {
      std::vector<const void*> srcs(P);
      for (int p = 0; p < P; ++p) srcs[p] = peerS(p) + (size_t)me * n * dt;
      bool alias = R >= S && R < S + sbytes;
      if (alias) {
        std::vector<char> tmp(n * dt);
        reduce_any(d.dtype, tmp.data(), srcs, n, d.rop, d.scale);
        sync(1);
        memcpy(R, tmp.data(), n * dt);
      } else {
        reduce_any(d.dtype, R, srcs, n, d.rop, d.scale);
        sync(1);
      }
      break;
    }
    case OpKind::ALLGATHER: {
      for (int p = 0; p < P; ++p) {
        char* dst = R + (size_t)p * n * dt;
        const char* src = peerS(p);
        if (dst != src) memcpy(dst, src, n * dt);
      }
      sync(1);
      break;
    }
    case OpKind::ALLGATHERV: {
      size_t off = 0;
      for (int p = 0; p < P; ++p) {
        char* dst = R + off * dt;
        const char* src = peerS(p);
        if (dst != src && d.recv_counts[p]) memcpy(dst, src, d.recv_counts[p] * dt);
        off += d.recv_counts[p];
      }
      sync(1);
      break;
    }
    case OpKind::BCAST: {
      // Bcast uses one buffer: it was published as the recv pointer
      if (me != (int)d.root) memcpy(R, peerR((int)d.root), n * dt);
      sync(1);
      break;
    }
    case OpKind::REDUCE: {
      if (me == (int)d.root) {
        std::vector<const void*> srcs(P);
        for (int p = 0; p < P; ++p) srcs[p] = peerS(p);
        reduce_any(d.dtype, R, srcs, n, d.rop, d.scale);
      }
      sync(1);
      break;
    }
    case OpKind::ALLTOALL: {
      bool alias = R == S;
      std::vector<char> tmp;
      char* dstbase = R;
      if (alias) {
        tmp.resize((size_t)P * n * dt);
        dstbase = tmp.data();
      }
      for (int p = 0; p < P; ++p) memcpy(dstbase + (size_t)p * n * dt, peerS(p) + (size_t)me * n * dt, n * dt);
      sync(1);
      if (alias) memcpy(R, tmp.data(), tmp.size());
      break;
    }
    case OpKind::ALLTOALLV:
    case OpKind::SENDRECV_LIST: {
      for (int p = 0; p < P; ++p) {
        size_t cnt = d.recv_counts[p];
        if (!cnt) continue;
        size_t soff = pub(g.members[p], prow)->aux[me];
        memcpy(R + d.recv_offsets[p] * dt, peerS(p) + soff * dt, cnt * dt);
      }
      sync(1);
      break;
    }
    case OpKind::GATHER: {
      if (me == (int)d.root)
        for (int p = 0; p < P; ++p) memcpy(R + (size_t)p * n * dt, peerS(p), n * dt);
      sync(1);
      break;
    }
    case OpKind::SCATTER: {
      const char* src = peerS((int)d.root) + (size_t)me * n * dt;
      if (R != src) memcpy(R, src, n * dt);
      sync(1);
      break;
    }
    case OpKind::FUSED_UPDATE: {
      // reduce-scatter -> optimizer on the owned shard -> all-gather of the updated parameters.
      // S: full gradient (P*n elements of d.dtype), d.fused.param: full parameters (P*n of out dtype).
      MLSLB_ASSERT(d.dtype == DType::F32 || d.dtype == DType::BF16, "fused update: gradient dtype must be f32/bf16");
      DType pdt = d.has_out_dtype ? d.out_dtype : d.dtype;
      MLSLB_ASSERT(pdt == DType::F32 || pdt == DType::BF16, "fused update: parameter dtype must be f32/bf16");
      const CommDesc::FusedUpdate& f = d.fused;
      std::vector<float> gsum(n);
      for (size_t i = 0; i < n; ++i) {
        float a = 0.f;
        for (int p = 0; p < P; ++p) {
          const char* sp = peerS(p) + ((size_t)me * n + i) * dt;
          a += d.dtype == DType::F32 ? *(const float*)sp : bf16_to_f32(*(const uint16_t*)sp);
        }
        gsum[i] = a * d.scale;
      }
      char* param = (char*)f.param;
      size_t pdts = dtype_size(pdt);
      host_optimizer_step(f, pdt, param + (size_t)me * n * pdts, gsum.data(), n);
      // publish parameter buffer for the gather step
      mine->recv_off = to_off(param);
      sync(1);
      for (int p = 0; p < P; ++p) {
        if (p == me) continue;
        const char* src = peer_ptr(g.members[p], pub(g.members[p], prow)->recv_off) + (size_t)p * n * pdts;
        memcpy(param + (size_t)p * n * pdts, src, n * pdts);
      }
      sync(2);
      break;
    }
    case OpKind::GEMM_RS:
      MLSLB_ASSERT(false, "GEMM+reduce-scatter is a device-only fused op");
      break;
    case OpKind::AG_GEMM:
      MLSLB_ASSERT(false, "all-gather+GEMM is a device-only fused op");
      break;
  }

  // ---- copy staged results back, drop the staging buffers -----------------------------------------------
  for (auto& s : stages) {
    if (recv_staged && s.user == r.recv) memcpy(s.user, s.slab, s.bytes);
    free(s.slab);
  }
}

// Quantised all-reduce with error feedback, CPU edition of the fused device kernel (same block format and the
// same three steps, so both backends agree to rounding): see quant.hpp.
void HostBackend::exec_quantized_allreduce(CommRequest& r, const ProcessGroup& g, int prow, char* S, char* R) {
  const CommDesc& d = r.desc;
  const size_t n = d.count;
  const int P = g.size(), me = g.idx;
  const uint64_t t = r.group_seq;
  HostPub* mine = pub(rank_, prow);
  HostReqState* st = (HostReqState*)r.backend_state;
  if (st->residual.size() != n) st->residual.assign(n, 0.f);
  const size_t nblk = ceil_div(n, (size_t)kQuantBlock);
  const size_t blk_per = ceil_div(nblk, (size_t)P);
  // staging: [q bytes | scales] x2 (own quantised input, reduced+requantised slice exchange)
  size_t qbytes = round_up(nblk * kQuantBlock, 64), sbytes = round_up(nblk * sizeof(float), 64);
  char* stage = (char*)alloc(2 * (qbytes + sbytes), 64);
  uint8_t* q1 = (uint8_t*)stage;
  float* s1 = (float*)(stage + qbytes);
  uint8_t* q2 = (uint8_t*)(stage + qbytes + sbytes);
  float* s2 = (float*)(stage + 2 * qbytes + sbytes);
  const float* x = (const float*)S;
  float* y = (float*)R;
  const bool mx = ctx_->env.tune.quant_mx != 0;          // four ue8m0 exponent bytes in place of the block's fp32 scale
  auto block_scale = [&](const float* sc, size_t b, size_t i) {   // scale of element i of block b
    return mx ? mx_scale_of(((const uint8_t*)(sc + b))[(i % kQuantBlock) / kMxBlock]) : sc[b];
  };
  // step 1: x + residual -> fp8 blocks, residual update
  for (size_t b = 0; b < nblk; ++b) {
    size_t lo = b * kQuantBlock, hi = std::min(n, lo + kQuantBlock);
    float v[kQuantBlock];
    for (size_t i = lo; i < hi; ++i) v[i - lo] = x[i] + st->residual[i];
    for (size_t i = hi - lo; i < (size_t)kQuantBlock; ++i) v[i] = 0.f;
    if (mx) quant_block_mx(v, q1 + lo, (uint8_t*)(s1 + b));
    else s1[b] = quant_block(v, q1 + lo);
    for (size_t i = lo; i < hi; ++i) st->residual[i] = v[i - lo] - e4m3_to_f32(q1[i]) * block_scale(s1, b, i - lo);
  }
  mine->send_off = to_off(stage);
  auto sync = [&](int step) {
    mine->phase.store(4 * t + step, std::memory_order_release);
    wait_all(g, prow, 4 * t + step);
  };
  sync(1);
  // step 2: my block range: dequant-accumulate over peers in fp32, requantise
  size_t blo = std::min(nblk, (size_t)me * blk_per), bhi = std::min(nblk, blo + blk_per);
  for (size_t b = blo; b < bhi; ++b) {
    float acc[kQuantBlock];
    for (int i = 0; i < kQuantBlock; ++i) acc[i] = 0.f;
    for (int p = 0; p < P; ++p) {
      const char* ps = peer_ptr(g.members[p], pub(g.members[p], prow)->send_off);
      const uint8_t* pq = (const uint8_t*)ps + b * kQuantBlock;
      const float* psc1 = (const float*)(ps + qbytes);
      for (int i = 0; i < kQuantBlock; ++i) acc[i] += e4m3_to_f32(pq[i]) * block_scale(psc1, b, (size_t)i);
    }
    if (mx) quant_block_mx(acc, q2 + b * kQuantBlock, (uint8_t*)(s2 + b));
    else s2[b] = quant_block(acc, q2 + b * kQuantBlock);
  }
  sync(2);
  // step 3: gather every slice, dequantise with the fused output scale
  for (int p = 0; p < P; ++p) {
    const char* ps = peer_ptr(g.members[p], pub(g.members[p], prow)->send_off);
    const uint8_t* pq = (const uint8_t*)ps + qbytes + sbytes;
    const float* psc = (const float*)(ps + 2 * qbytes + sbytes);
    size_t plo = std::min(nblk, (size_t)p * blk_per), phi = std::min(nblk, plo + blk_per);
    for (size_t b = plo; b < phi; ++b) {
      size_t lo = b * kQuantBlock, hi = std::min(n, lo + kQuantBlock);
      for (size_t i = lo; i < hi; ++i) y[i] = e4m3_to_f32(pq[i]) * block_scale(psc, b, i - lo) * d.scale;
    }
  }
  sync(3);
  free(stage);
}

// Quantised all-reduce through a USER compression library, the reference's only flavour (reference quant/quant.c:
// dlopen + three dlsym'ed functions, quantise in place with a per-buffer error-feedback residual, all-reduce of
// ceil(count / elem_in_block) opaque blocks of block_size bytes with the library's own block sum, dequantise in place;
// hook-up in eplib/cqueue.c:1977-1994, 2283-2284).  Here the block range is split over the ranks: every rank sums its
// slice of blocks over all peers in a fixed order with the plugin's reduce function, then the slices are gathered, so
// the result is bitwise identical everywhere.
void HostBackend::load_quant_plugin() {
  if (plugin_lib_) return;
  const QuantConfig& q = ctx_->quant;
  plugin_lib_ = dlopen(q.lib_path.c_str(), RTLD_NOW | RTLD_LOCAL);
  MLSLB_ASSERT(plugin_lib_ != nullptr, "quantization library can't be loaded: %s", dlerror());
  plugin_quant_ = (PluginQuant)dlsym(plugin_lib_, q.quant_name.c_str());
  MLSLB_ASSERT(plugin_quant_ != nullptr, "quantization function can't be loaded: %s", q.quant_name.c_str());
  plugin_dequant_ = (PluginDequant)dlsym(plugin_lib_, q.dequant_name.c_str());
  MLSLB_ASSERT(plugin_dequant_ != nullptr, "dequantization function can't be loaded: %s", q.dequant_name.c_str());
  plugin_reduce_ = (PluginReduce)dlsym(plugin_lib_, q.reduce_name.c_str());
  MLSLB_ASSERT(plugin_reduce_ != nullptr, "reduce function can't be loaded: %s", q.reduce_name.c_str());
  MLSLB_ASSERT(q.block_size > 0 && q.elem_in_block > 0, "quantization block_size / elem_in_block must be positive");
  MLSLB_LOG(LOG_INFO, "quantization plugin %s: block %zu bytes / %zu elements", q.lib_path.c_str(), q.block_size,
            q.elem_in_block);
}

void HostBackend::exec_plugin_allreduce(CommRequest& r, const ProcessGroup& g, int prow, char* S, char* R) {
  load_quant_plugin();
  const CommDesc& d = r.desc;
  const QuantConfig& q = ctx_->quant;
  const size_t n = d.count;
  const int P = g.size(), me = g.idx;
  const uint64_t t = r.group_seq;
  HostPub* mine = pub(rank_, prow);
  HostReqState* st = (HostReqState*)r.backend_state;
  if (st->residual.size() != n) st->residual.assign(n, 0.f);
  const size_t nblk = ceil_div(n, q.elem_in_block);
  const size_t blk_per = ceil_div(nblk, (size_t)P);
  // the plugin works in place on a buffer of `count` floats (compressed data occupies its head) and may touch whole
  // blocks, so both areas are sized for max(count floats, nblk blocks)
  const size_t area = round_up(std::max(n * sizeof(float), nblk * q.block_size) + q.block_size, 64);
  char* stage = (char*)alloc(2 * area, 64);
  char* mineq = stage;          // my quantised input
  char* red = stage + area;     // my slice of reduced blocks (at its global block position)
  memcpy(mineq, S, n * sizeof(float));
  // DL_COMP_FLOAT32 = 2, compression ratio 4, DL_COMP_DFP = 1: the constants the reference passes (quant/quant.c:201)
  int rc = plugin_quant_(mineq, mineq, n, st->residual.data(), 2, 4, 1);
  MLSLB_ASSERT(rc == 0, "quantization failed: error code %d", rc);
  mine->send_off = to_off(stage);
  auto sync = [&](int step) {
    mine->phase.store(4 * t + step, std::memory_order_release);
    wait_all(g, prow, 4 * t + step);
  };
  sync(1);
  const size_t blo = std::min(nblk, (size_t)me * blk_per), bhi = std::min(nblk, blo + blk_per);
  if (bhi > blo) {
    const char* p0 = peer_ptr(g.members[0], pub(g.members[0], prow)->send_off);
    memcpy(red + blo * q.block_size, p0 + blo * q.block_size, (bhi - blo) * q.block_size);
    for (int p = 1; p < P; ++p) {
      const char* ps = peer_ptr(g.members[p], pub(g.members[p], prow)->send_off);
      rc = plugin_reduce_(ps + blo * q.block_size, red + blo * q.block_size, bhi - blo);
      MLSLB_ASSERT(rc == 0, "quantized reduction failed: error code %d", rc);
    }
  }
  sync(2);
  // gather the reduced slices into my first area (everyone is past reading it), dequantise in place
  for (int p = 0; p < P; ++p) {
    const char* ps = peer_ptr(g.members[p], pub(g.members[p], prow)->send_off) + area;
    size_t plo = std::min(nblk, (size_t)p * blk_per), phi = std::min(nblk, plo + blk_per);
    if (phi > plo) memcpy(mineq + plo * q.block_size, ps + plo * q.block_size, (phi - plo) * q.block_size);
  }
  sync(3);
  rc = plugin_dequant_(mineq, mineq, n);
  MLSLB_ASSERT(rc == 0, "dequantization failed: error code %d", rc);
  const float* v = (const float*)mineq;
  float* y = (float*)R;
  for (size_t i = 0; i < n; ++i) y[i] = v[i] * d.scale;
  free(stage);
}

}  // namespace

// ---- shared with the net backend ---------------------------------------------------------------------------------------
void host_reduce(DType dt, void* dst, const std::vector<const void*>& srcs, size_t n, RedOp op, float scale) {
  reduce_any(dt, dst, srcs, n, op, scale);
}

// SGD(momentum) / AdamW on one shard: `param_owned` points at the first owned element of the (fp32 or bf16) parameters,
// gsum holds the reduced, already scaled gradient of the shard (same arithmetic as the fused device kernel).
void host_optimizer_step(const CommDesc::FusedUpdate& f, DType pdt, char* param_owned, const float* gsum, size_t n) {
  const size_t pdts = dtype_size(pdt);
  float* master = (float*)f.master;
  float* m1 = (float*)f.state1;
  float* m2 = (float*)f.state2;
  float bc1 = 1.f, bc2 = 1.f;
  if (f.optimizer == 1) {
    bc1 = 1.f - powf(f.beta1, (float)f.step);
    bc2 = 1.f - powf(f.beta2, (float)f.step);
  }
  for (size_t i = 0; i < n; ++i) {
    char* pp = param_owned + i * pdts;
    float w = master ? master[i] : (pdt == DType::F32 ? *(float*)pp : bf16_to_f32(*(uint16_t*)pp));
    float gr = gsum[i];
    if (f.optimizer == 0) {
      gr += f.weight_decay * w;
      if (m1) {
        m1[i] = f.momentum * m1[i] + gr;
        gr = m1[i];
      }
      w -= f.lr * gr;
    } else {
      m1[i] = f.beta1 * m1[i] + (1.f - f.beta1) * gr;
      m2[i] = f.beta2 * m2[i] + (1.f - f.beta2) * gr * gr;
      float mh = m1[i] / bc1, vh = m2[i] / bc2;
      w -= f.lr * (mh / (sqrtf(vh) + f.eps) + f.weight_decay * w);
    }
    if (master) master[i] = w;
    if (pdt == DType::F32) *(float*)pp = w;
    else *(uint16_t*)pp = f32_to_bf16(w);
  }
}

std::unique_ptr<Backend> make_host_backend(RankContext* ctx) { return std::unique_ptr<Backend>(new HostBackend(ctx)); }

}  // namespace mlslb
