// CUDA backend: symmetric device heap + peer mappings + stream/event plumbing around the collective kernels.
//
// What the reference builds out of POSIX shared memory, dlmalloc and ep_server processes (reference eplib/memory.c,
// eplib/client.c, eplib/server.c, src/comm_ep.cpp:363-566) becomes, on an NVSwitch node:
//   * one device slab per rank, mapped by every peer (CUDA IPC between processes, plain pointers between in-process
//     ranks) - Environment::Alloc sub-allocates from it, so user buffers are directly addressable by peer kernels;
//   * buffers that are NOT in the slab (arbitrary cudaMalloc / host memory) are staged through slab scratch on the
//     same stream - the GPU analogue of the reference's ReplaceIn / ReplaceOut shadow-buffer copies;
//   * every (group, lane) has its own CUDA stream (lane 1 = high priority): Start() makes it wait for the user's
//     stream, Wait() either blocks the host on the completion event or just orders the user's stream after it.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <fcntl.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstring>
#include <map>
#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

#include "core/log.hpp"
#include "core/quant.hpp"
#include "core/runtime.hpp"
#include "core/sysinfo.hpp"
#include "cuda/kernels.hpp"
#include "cuda/vmm.hpp"

namespace mlslb {

namespace {

constexpr size_t kPadRowBytes = (size_t)kMaxChannels * kMaxDevRanks * sizeof(PadSlot);   // per (row, lane)
constexpr size_t kSeqRowBytes = (size_t)kMaxChannels * sizeof(unsigned long long);
constexpr int kPadRows = kMaxGroupRows * 2;
constexpr size_t kPadBytes = kPadRows * kPadRowBytes;
constexpr size_t kSeqBase = kPadBytes;
constexpr size_t kLLBase = (kPadBytes + kPadRows * kSeqRowBytes + 4095) & ~(size_t)4095;   // low-latency arenas, one per row
constexpr size_t kMidSeqBase = kLLBase + kLLRows * kLLRowBytes;                      // launch counters of the mid kernel, per row
constexpr size_t kMidSeqRowBytes = (size_t)kMidCtas * sizeof(unsigned long long);
constexpr size_t kMidBase = (kMidSeqBase + kMidRows * kMidSeqRowBytes + 4095) & ~(size_t)4095;   // mid-size arenas
constexpr size_t kHeaderBytes = (kMidBase + kMidRows * kMidRowBytes + ((size_t)1 << 20) - 1) >> 20 << 20;

struct StageBuf {
  void* user;
  void* slab;
  size_t bytes;
  bool copy_out;
};

struct CudaReqState {
  cudaEvent_t ready = nullptr, done = nullptr;
  std::vector<StageBuf> stages;
  float* residual = nullptr;      // error-feedback residual (quantised all-reduce)
  void* qstage = nullptr;         // fp8 staging area inside the slab
  size_t residual_elems = 0;
  cudaStream_t stream = nullptr;
  bool inflight = false;
  bool recorded = false;          // `done` was recorded for the current launch
  cudaEvent_t t0 = nullptr, t1 = nullptr;   // timing pair around the kernel (statistics / trace), from the timing pool
  bool timed = false;             // t0 / t1 were recorded for a run whose duration has not been read yet
};

struct PeerInfo {
  cudaIpcMemHandle_t handle;
  unsigned long long ptr;
  int pid;
  int device;
  int pad;
};

class CudaBackend final : public Backend {
 public:
  explicit CudaBackend(RankContext* ctx) : ctx_(ctx) { init(); }
  ~CudaBackend() override {}

  const char* name() const override { return "cuda"; }
  bool is_device() const override { return true; }

  void* alloc(size_t bytes, size_t align) override {
    size_t off = heap_.alloc(bytes, std::max<size_t>(align, 256));
    if (off == SIZE_MAX) {   // scratch of finished stream-ordered requests may still be parked
      sweep_parked();
      off = heap_.alloc(bytes, std::max<size_t>(align, 256));
    }
    if (off == SIZE_MAX && grow_heap(bytes + std::max<size_t>(align, 256))) off = heap_.alloc(bytes, std::max<size_t>(align, 256));
    MLSLB_ASSERT(off != SIZE_MAX,
                 "symmetric device heap exhausted (%zu bytes requested, %zu of %zu in use%s): raise MLSL_HEAP_SIZE_GB%s",
                 bytes, heap_.bytes_in_use(), heap_.capacity(),
                 vmm_.ok ? ", growth failed or the reserved range is full" : "; this slab cannot grow (CUDA IPC / in-process ranks)",
                 vmm_.ok ? " / MLSL_HEAP_MAX_GB" : "");
    void* p = slab_ + off;
    ctx_->ptrcheck.add(p, bytes);
    return p;
  }
  void free(void* p) override {
    if (!p) return;
    ctx_->ptrcheck.remove(p);
    MLSLB_ASSERT(heap_.free((size_t)((char*)p - slab_)), "Free of a pointer that did not come from Alloc");
  }
  bool owns(const void* p, size_t len) const override {
    const char* c = (const char*)p;
    return c >= slab_ + kHeaderBytes && c + len <= slab_ + slab_bytes_;
  }

  void group_created(ProcessGroup& g) override {
    if (g.size() <= 1 || g.row < 0) return;
    MLSLB_ASSERT(g.size() <= kMaxDevRanks, "device groups are limited to %d ranks (got %d)", kMaxDevRanks, g.size());
    set_device();
    // fresh row: zero my pads and ticket counters, then make sure every member has done so before anyone signals
    // (on the row's own stream: every extra stream is one more hardware queue that loop-back ranks sharing a GPU
    // compete for, and two spinning kernels falsely serialised on one queue dead-lock each other)
    cudaStream_t zs = inline_stream_ && user_stream_set_ ? user_stream_ : stream_for(g.row, 0);
    MLSLB_CUDA(cudaMemsetAsync(slab_ + (size_t)g.row * 2 * kPadRowBytes, 0, 2 * kPadRowBytes, zs));
    MLSLB_CUDA(cudaMemsetAsync(slab_ + kSeqBase + (size_t)g.row * 2 * kSeqRowBytes, 0, 2 * kSeqRowBytes, zs));
    if (g.row < kLLRows) MLSLB_CUDA(cudaMemsetAsync(slab_ + kLLBase + (size_t)g.row * kLLRowBytes, 0, kLLRowBytes, zs));
    if (g.row < kMidRows) {
      MLSLB_CUDA(cudaMemsetAsync(slab_ + kMidBase + (size_t)g.row * kMidRowBytes, 0, kMidRowBytes, zs));
      MLSLB_CUDA(cudaMemsetAsync(slab_ + kMidSeqBase + (size_t)g.row * kMidSeqRowBytes, 0, kMidSeqRowBytes, zs));
    }
    MLSLB_CUDA(cudaStreamSynchronize(zs));
    if (BootCtl* c = ctx_->boot->ctl())
      for (int l = 0; l < 2; ++l) c->launch_seq[g.row * 2 + l][ctx_->rank].store(0, std::memory_order_release);
    if (!g.is_world) ctx_->group_barrier(&g);
  }

  void prepare(CommRequest& r) override {
    if (r.backend_state) return;
    r.backend_state = new CudaReqState();   // events are created on first use (none at all in inline-stream mode)
  }
  // Events come from a pool filled at start-up and are never destroyed before finalize: cuEventCreate / cuEventDestroy
  // need the writer side of a driver lock that a host thread sitting in a pageable device-to-host copy behind a
  // spinning kernel holds for the whole copy (csrc/tools/probe_blocking.cu) - with loop-back ranks that is a dead-lock.
  cudaEvent_t take_event() {
    {
      std::lock_guard<std::mutex> g(park_mu_);
      if (!event_pool_.empty()) {
        cudaEvent_t ev = event_pool_.back();
        event_pool_.pop_back();
        return ev;
      }
    }
    cudaEvent_t ev = nullptr;
    MLSLB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    return ev;
  }
  void give_event(cudaEvent_t ev) {
    if (!ev) return;
    std::lock_guard<std::mutex> g(park_mu_);
    event_pool_.push_back(ev);
  }
  // device timestamps: one timing-enabled event pair per request brackets the kernel on its stream
  bool want_timing() const { return ctx_->env.tune.dev_timestamps && (ctx_->env.stats || !ctx_->trace_prefix.empty()); }
  cudaEvent_t take_timing_event() {
    {
      std::lock_guard<std::mutex> g(park_mu_);
      if (!tevent_pool_.empty()) {
        cudaEvent_t ev = tevent_pool_.back();
        tevent_pool_.pop_back();
        return ev;
      }
    }
    cudaEvent_t ev = nullptr;
    MLSLB_CUDA(cudaEventCreate(&ev));
    return ev;
  }
  void harvest(CommRequest& r, CudaReqState* st, bool complete_for_sure) {
    if (!st->timed) return;
    if (!complete_for_sure) {
      cudaError_t q = cudaEventQuery(st->t1);
      if (q == cudaErrorNotReady) {
        cudaGetLastError();
        return;
      }
    }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, st->t0, st->t1) == cudaSuccess) {
      r.device_ns_last = (uint64_t)((double)ms * 1e6);
      r.device_ns_pending += r.device_ns_last;
    } else {
      cudaGetLastError();
    }
    st->timed = false;
  }
  void harvest_device_time(CommRequest& r) override {
    if (auto* st = (CudaReqState*)r.backend_state) harvest(r, st, false);
  }
  bool supports_strided_alltoall() const override { return true; }
  bool stream_ordered_wait() const override { return stream_wait_; }
  bool peek_done(CommRequest& r) override {
    auto* st = (CudaReqState*)r.backend_state;
    if (!st || !st->recorded) return true;
    set_device();
    cudaError_t e = cudaEventQuery(st->done);
    if (e == cudaErrorNotReady) {
      cudaGetLastError();
      return false;
    }
    return true;
  }
  int default_servers() const override { return (inproc_ || inline_stream_ || ranks_per_device_ > 1) ? 0 : 1; }
  void ensure_events(CudaReqState* st) {
    if (st->done) return;
    st->ready = take_event();
    st->done = take_event();
  }
  // inline-stream + stream-ordered wait: the collective is just a kernel on the user's stream, nothing to track
  bool eventless() const { return inline_stream_ && stream_wait_; }
  void release(CommRequest& r) override {
    auto* st = (CudaReqState*)r.backend_state;
    if (!st) return;
    set_device();
    if (st->inflight && st->recorded) cudaEventSynchronize(st->done);
    else if (st->inflight && st->stream) cudaStreamSynchronize(st->stream);
    // Scratch that a stream-ordered (not host-waited) kernel may still be using must not be recycled early.  It is
    // parked behind an event instead of synchronising the stream: one-shot requests (Distribution::*, the fused GEMM
    // ops) are released inside Environment::Wait, which must not block the CPU in stream-ordered mode.
    if (st->residual) st->stages.push_back(StageBuf{nullptr, st->residual, 0, false});
    if (st->qstage) st->stages.push_back(StageBuf{nullptr, st->qstage, 0, false});
    st->residual = nullptr;
    st->qstage = nullptr;
    if (!st->stages.empty()) {
      if (st->stream) park_stages(st);
      else drop_stages(st, false);
    }
    give_event(st->ready);
    give_event(st->done);
    if (st->t0) {
      std::lock_guard<std::mutex> g(park_mu_);
      tevent_pool_.push_back(st->t0);
      tevent_pool_.push_back(st->t1);
    }
    delete st;
    r.backend_state = nullptr;
  }

  void on_start(CommRequest& r) override {
    // runs on the API thread: pin the point of the user's stream the collective has to wait for
    if (inline_stream_) return;
    auto* st = (CudaReqState*)r.backend_state;
    set_device();
    ensure_events(st);
    MLSLB_CUDA(cudaEventRecord(st->ready, ustream()));
  }

  void launch(CommRequest& r) override;
  // ---- device-heap expansion (reference eplib/memory.c:396-410, eplib/cqueue.c:1451-1510) ---------------------------------
  // The slab of a VMM job sits at the start of a much larger reserved address range.  When Alloc runs out of room this rank
  // backs the next part of ITS range with a new chunk (at least as big as the slab so far: sizes double), publishes the
  // chunk's descriptor in the shared control block and waits until every peer's watcher thread has mapped it at the same
  // offset of the range it reserved for this rank - only then can an offset inside the chunk reach a peer's kernel.
  std::thread grow_watcher_;
  std::atomic<bool> grow_stop_{false};
  std::vector<uint64_t> grow_seen_;
  std::mutex grow_mu_;
  bool grow_heap(size_t need) {
    if (!vmm_.ok || !ctx_->boot->ctl()) return false;
    std::lock_guard<std::mutex> g(grow_mu_);
    BootCtl* c = ctx_->boot->ctl();
    GrowRec& rec = c->grow[ctx_->rank];
    const uint64_t n = rec.gen.load(std::memory_order_relaxed);
    if (n >= (uint64_t)kMaxGrowChunks) return false;
    size_t add = std::max(need, vmm_.local_mapped - 0), off = 0, got = 0;
    add = std::min(add, vmm_.reserved - vmm_.local_mapped);
    if (add < need) return false;
    set_device();
    const int fd = vmm_slab_grow(vmm_, add, &off, &got);
    if (fd < 0) return false;
    rec.pid = (int64_t)getpid();
    rec.chunk[n].fd = fd;
    rec.chunk[n].off = off;
    rec.chunk[n].bytes = got;
    rec.gen.store(n + 1, std::memory_order_release);
    // hand the descriptor to every peer's watcher thread (SCM_RIGHTS over the bootstrap's sockets)
    const uint64_t msg[3] = {(uint64_t)ctx_->rank, (uint64_t)off, (uint64_t)got};
    for (int p = 0; p < ctx_->world; ++p)
      if (p != ctx_->rank && !ctx_->boot->send_fd_to(p, fd, msg)) {
        MLSLB_LOG(LOG_ERROR, "heap growth: cannot send the chunk descriptor to rank %d", p);
        return false;
      }
    // wait for every peer (their watcher threads run independently of what their API threads are doing)
    const uint64_t t0 = now_ns();
    for (int p = 0; p < ctx_->world; ++p) {
      if (p == ctx_->rank) continue;
      while (c->grow_ack[ctx_->rank][p].load(std::memory_order_acquire) < n + 1) {
        usleep(50);
        if (ctx_->boot->poisoned() || now_ns() - t0 > 30000000000ull) {
          MLSLB_LOG(LOG_ERROR, "heap growth: rank %d never mapped the new chunk", p);
          return false;
        }
      }
    }
    close(fd);                        // every peer holds its own duplicate now
    heap_.extend(got);
    slab_bytes_ = vmm_.local_mapped;
    MLSLB_LOG(LOG_INFO, "device heap of rank %d grew by %.2f GiB to %.2f GiB", ctx_->rank, got / 1073741824.0, slab_bytes_ / 1073741824.0);
    return true;
  }
  void grow_watch() {
    BootCtl* c = ctx_->boot->ctl();
    cudaSetDevice(device_);
    while (!grow_stop_.load(std::memory_order_acquire)) {
      bool any = false;
      int fd = -1;
      uint64_t msg[3];
      while (ctx_->boot->try_recv_fd(&fd, msg)) {
        const int p = (int)msg[0];
        const bool ok = p >= 0 && p < ctx_->world && vmm_slab_map_peer_chunk(vmm_, p, fd, (size_t)msg[1], (size_t)msg[2]);
        close(fd);
        if (!ok) {
          MLSLB_LOG(LOG_ERROR, "heap growth: cannot map the chunk rank %d added at offset %llu", p, (unsigned long long)msg[1]);
          ctx_->boot->poison(ctx_->rank);
          return;
        }
        grow_seen_[p]++;
        c->grow_ack[p][ctx_->rank].store(grow_seen_[p], std::memory_order_release);
        any = true;
      }
      if (!any) usleep(200);
    }
  }
  // Loop-back ranks (several ranks on ONE GPU: the single-GPU test mode): a kernel that spins on the device for a peer
  // that has not launched yet starves every host call that needs the device idle for a moment - a first kernel with a
  // bigger stack, cudaFree, cublasCreate, any allocation while a thread sits in a pageable copy (probe_blocking.cu) -
  // and if the late peer is inside such a call the job dead-locks.  So the members first meet on the HOST: each counts
  // its launch on the (row, lane) and waits until every member has counted the same one; then all launch within
  // microseconds of each other.  A bounded wait: ranks whose collectives are started from ONE shared thread (PyTorch's
  // autograd thread runs the hooks of all in-process ranks) can never meet, they fall through and spin as before.
  void loopback_rendezvous(const ProcessGroup& g, int lane) {
    const long ms = ctx_->env.tune.loopback_rendezvous_ms;
    BootCtl* c = ctx_->boot->ctl();
    if (ranks_per_device_ <= 1 || ms <= 0 || !c || g.row < 0 || g.size() <= 1) return;
    std::atomic<uint64_t>* row = c->launch_seq[g.row * 2 + lane];
    const uint64_t v = row[ctx_->rank].fetch_add(1, std::memory_order_acq_rel) + 1;
    const uint64_t deadline = now_ns() + (uint64_t)ms * 1000000ull;
    for (int m : g.members) {
      unsigned spins = 0;
      while (row[m].load(std::memory_order_acquire) < v) {
        if ((++spins & 63) == 0) {
          if (now_ns() > deadline || ctx_->boot->poisoned()) return;
          sched_yield();
        }
      }
    }
  }
  bool test(CommRequest& r) override {
    auto* st = (CudaReqState*)r.backend_state;
    set_device();
    cudaError_t e = st->recorded ? cudaEventQuery(st->done) : cudaStreamQuery(st->stream);
    if (e == cudaErrorNotReady) {
      cudaGetLastError();   // not an error, but it would be picked up by the next launch check
      return false;
    }
    MLSLB_CUDA(e);
    finish(r, st);
    return true;
  }
  void wait(CommRequest& r) override {
    auto* st = (CudaReqState*)r.backend_state;
    set_device();
    if (stream_wait_) {
      // stream-ordered completion: nothing blocks the host.  Staged (foreign) buffers were copied back on the
      // collective's stream right behind the kernel; only their slab scratch has to outlive the kernel, so it is
      // parked with an event and recycled by a later call.
      if (!inline_stream_) MLSLB_CUDA(cudaStreamWaitEvent(ustream(), st->done, 0));
      if (!st->stages.empty()) park_stages(st);
      sweep_parked();
      st->inflight = false;
      check_error(opkind_name(r.desc.kind));   // a device watchdog hit of an EARLIER operation surfaces here
      return;
    }
    if (st->recorded) MLSLB_CUDA(cudaEventSynchronize(st->done));
    else MLSLB_CUDA(cudaStreamSynchronize(st->stream));
    finish(r, st);
  }

  // A device-to-host copy into PAGEABLE memory blocks the calling thread until the stream has drained and holds a driver
  // lock all that time (no event / stream / allocation call of any other thread gets through, probe_blocking.cu).
  // Drain the stream first - the call blocks either way - so the lock is only held for the copy itself.
  bool is_pageable_host(const void* p) const {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
      cudaGetLastError();
      return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
  }
  void copy_out(const StageBuf& sb, cudaStream_t s) {
    if (is_pageable_host(sb.user)) MLSLB_CUDA(cudaStreamSynchronize(s));
    MLSLB_CUDA(cudaMemcpyAsync(sb.user, sb.slab, sb.bytes, cudaMemcpyDefault, s));
  }
  bool is_device_pointer(const void* p) const override {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
  }
  void copy_from_host(void* dst, const void* src, size_t bytes) override {
    cudaSetDevice(device_);
    MLSLB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, aux_stream()));
    MLSLB_CUDA(cudaStreamSynchronize(aux_stream()));
  }
  uint64_t heap_offset(const void* p) const override { return (uint64_t)((const char*)p - slab_); }
  void* peer_heap_ptr(int global_rank, uint64_t offset) override { return peer_slab_[global_rank] + offset; }
  void rma_copy(void* dst, const void* src, size_t bytes) override {
    set_device();
    MLSLB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ustream()));   // peers are mapped: a plain copy
  }
  void set_user_stream(void* s) override {
    user_stream_ = (cudaStream_t)s;
    user_stream_set_ = true;   // null is a valid choice: the legacy default stream
  }
  void* user_stream() override { return (void*)user_stream_; }
  void set_wait_mode(bool stream_ordered) override { stream_wait_ = stream_ordered; }

  void pack_blocks(const BlockDesc* blocks, size_t nblocks, size_t local_fm_count, DType dt, const void* src,
                   void* dst, bool unpack) override {
    set_device();
    size_t mx = 0;
    for (size_t i = 0; i < nblocks; ++i) mx = std::max(mx, blocks[i].mb_cnt * blocks[i].fm_cnt * blocks[i].fm_size);
    for (size_t i = 0; i < nblocks; i += kMaxDevRanks) {
      PackPlan plan;
      plan.n = (int)std::min<size_t>(kMaxDevRanks, nblocks - i);
      for (int k = 0; k < plan.n; ++k) plan.b[k] = blocks[i + k];
      MLSLB_CUDA(launch_pack_blocks(plan, local_fm_count, (int)dtype_size(dt), src, dst, unpack, mx, ustream()));
    }
  }

  void finalize() override {
    set_device();
    // a poisoned job cannot synchronise any more: release the resources anyway (kernels have watchdog deadlines)
    auto quiet_barrier = [&] {
      try {
        ctx_->boot->barrier();
      } catch (const std::exception&) {
      }
    };
    cudaDeviceSynchronize();
    if (grow_watcher_.joinable()) {
      quiet_barrier();                 // nobody is growing any more
      grow_stop_.store(true, std::memory_order_release);
      grow_watcher_.join();
    }
    sweep_parked(true);
    for (int b = 0; b < kPipeBufs; ++b) {
      give_event(pipe_h2d_[b]);
      give_event(pipe_ar_[b]);
      give_event(pipe_d2h_[b]);
      pipe_h2d_[b] = pipe_ar_[b] = pipe_d2h_[b] = nullptr;
    }
    give_event(pipe_start_);
    pipe_start_ = nullptr;
    for (cudaEvent_t ev : event_pool_) cudaEventDestroy(ev);
    event_pool_.clear();
    for (cudaEvent_t ev : tevent_pool_) cudaEventDestroy(ev);
    tevent_pool_.clear();
    quiet_barrier();
    for (size_t p = 0; p < peer_slab_.size(); ++p)
      if (peer_opened_[p]) cudaIpcCloseMemHandle(peer_slab_[p]);
    quiet_barrier();
    for (auto& s : streams_)
      if (s) cudaStreamDestroy(s);
    if (aux_stream_) cudaStreamDestroy(aux_stream_);
    if (own_user_stream_) cudaStreamDestroy(own_user_stream_);
    if (vmm_.ok) vmm_slab_destroy(vmm_);
    else if (slab_ && !(inproc_ && ctx_->boot->poisoned())) cudaFree(slab_);   // failed loop-back job: a peer thread's kernel
                                                                              // may still touch it - leak rather than fault
    if (err_host_) cudaFreeHost((void*)err_host_);
    slab_ = nullptr;
  }

  std::string describe() const override {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuda peer-memory backend (device %d, %d SMs, slab %.1f GiB, %s%s, ranks/device %d)", device_,
             sm_count_, slab_bytes_ / 1073741824.0, inproc_ ? "in-process ranks" : (vmm_.ok ? "VMM fd-shared" : "CUDA IPC"),
             mc_ ? " + NVLS multicast" : "", ranks_per_device_);
    return buf;
  }

 private:
  RankContext* ctx_;
  int device_ = 0, sm_count_ = 148, ranks_per_device_ = 1;
  bool inproc_ = false;
  char* slab_ = nullptr;
  size_t slab_bytes_ = 0;
  SlabAllocator heap_;
  std::vector<char*> peer_slab_;
  std::vector<bool> peer_opened_;
  std::vector<cudaStream_t> streams_;
  cudaStream_t aux_stream_ = nullptr, user_stream_ = nullptr, own_user_stream_ = nullptr;
  int stream_prio_hi_ = 0;
  bool user_stream_set_ = false;
  // both created on first use only (see group_created for why streams are rationed)
  cudaStream_t aux_stream() {
    if (!aux_stream_) MLSLB_CUDA(cudaStreamCreateWithPriority(&aux_stream_, cudaStreamNonBlocking, stream_prio_hi_));
    return aux_stream_;
  }
  cudaStream_t ustream() {
    if (user_stream_set_) return user_stream_;
    if (!own_user_stream_) MLSLB_CUDA(cudaStreamCreateWithFlags(&own_user_stream_, cudaStreamNonBlocking));
    return own_user_stream_;
  }
  bool stream_wait_ = false, inline_stream_ = false;
  // host-buffer pipeline (see launch_host_pipelined)
  static constexpr int kPipeBufs = 8;          // upper bound; MLSL_PIPE_BUFS of them are used
  int pipe_bufs_ = 4;
  size_t pipe_chunk_ = (size_t)16 << 20;
  char* pipe_buf_[kPipeBufs] = {};
  cudaEvent_t pipe_h2d_[kPipeBufs] = {}, pipe_ar_[kPipeBufs] = {}, pipe_d2h_[kPipeBufs] = {}, pipe_start_ = nullptr;
  cudaStream_t h2d_stream_ = nullptr, d2h_stream_ = nullptr;
  bool pipe_used_[kPipeBufs] = {};
  void bind_to_gpu_numa_node();
  bool launch_host_pipelined(CommRequest& r, const DevComm& dc, cudaStream_t s, bool solo = false);
  VmmSlab vmm_;
  char* mc_ = nullptr;
  volatile int* err_host_ = nullptr;
  int* err_dev_ = nullptr;
  std::mutex mu_;

  void set_device() { cudaSetDevice(device_); }
  void init();
  cudaStream_t stream_for(int row, int lane);
  int pick_channels(size_t bytes) const;
  int ar_channels(size_t bytes) const {   // large all-reduce: MLSL_AR_CHANNELS caps / raises the grid of that kernel only
    const long c = ctx_->env.tune.ar_channels;
    if (c <= 0) return pick_channels(bytes);
    size_t want = ceil_div(std::max<size_t>(bytes, 1), (size_t)kCommThreads * 16);
    int cap = std::min(kMaxChannels, std::max(1, sm_count_ / std::max(1, ranks_per_device_)));
    return (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)std::min<long>(c, cap)));
  }
  // fp8 all-reduce: one warp per 128-element block and pass; HBM-bound local phases -> up to one 1024-thread CTA per SM
  int quant_channels(size_t elems) const {
    size_t blocks = ceil_div(std::max<size_t>(elems, 1), (size_t)128);
    size_t want = ceil_div(blocks, (size_t)32 * 2);
    int cap = std::min(kMaxChannels, std::max(1, sm_count_ / std::max(1, ranks_per_device_)));
    return (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)cap));
  }
  DevComm make_comm(const ProcessGroup& g, int lane) const;
  void drop_stages(CudaReqState* st, bool copied);
  // scratch blocks of stream-ordered requests: freed once the event recorded behind their last use has fired
  struct Parked {
    cudaEvent_t ev;
    void* slab;
  };
  std::vector<Parked> parked_;
  std::vector<cudaEvent_t> event_pool_;
  std::vector<cudaEvent_t> tevent_pool_;    // timing-enabled events (device timestamps)
  std::mutex park_mu_;
  void park_stages(CudaReqState* st) {
    std::lock_guard<std::mutex> g(park_mu_);
    cudaEvent_t ev = nullptr;
    if (!event_pool_.empty()) {
      ev = event_pool_.back();
      event_pool_.pop_back();
    } else if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
      ev = nullptr;
    }
    // never throws (also runs from destructors): if the event cannot be recorded - e.g. the stream is already gone at
    // teardown, after a device synchronise - the blocks are simply released
    if (!ev || cudaEventRecord(ev, st->stream) != cudaSuccess) {
      cudaGetLastError();
      if (ev) event_pool_.push_back(ev);
      for (auto& sb : st->stages) heap_.free((size_t)((char*)sb.slab - slab_)), ctx_->ptrcheck.remove(sb.slab);
      st->stages.clear();
      return;
    }
    // one event guards all blocks of the request: the first entry owns it, the others piggy-back (null event)
    bool first = true;
    for (auto& sb : st->stages) {
      parked_.push_back(Parked{first ? ev : nullptr, sb.slab});
      first = false;
    }
    st->stages.clear();
  }
  void sweep_parked(bool all = false) {
    std::vector<void*> to_free;
    {
      std::lock_guard<std::mutex> g(park_mu_);
      size_t i = 0;
      while (i < parked_.size()) {
        // entries come in groups [event, null, null...]; a group is released as a whole
        size_t j = i + 1;
        while (j < parked_.size() && parked_[j].ev == nullptr) ++j;
        cudaError_t qe = all ? cudaSuccess : cudaEventQuery(parked_[i].ev);
        if (qe == cudaErrorNotReady) cudaGetLastError();   // "not ready" is recorded as the thread's last error: clear it
        if (qe == cudaSuccess) {
          event_pool_.push_back(parked_[i].ev);
          for (size_t k = i; k < j; ++k) to_free.push_back(parked_[k].slab);
          parked_.erase(parked_.begin() + i, parked_.begin() + j);
        } else {
          i = j;
        }
      }
    }
    for (void* p : to_free) free(p);
  }
  void finish(CommRequest& r, CudaReqState* st);
  void check_error(const char* what);
  void launch_single(CommRequest& r, CudaReqState* st, cudaStream_t s);
};

// One process per GPU: run the process (and place the memory it allocates from now on - pinned staging buffers, the
// user's pinned tensors) on the NUMA node the GPU hangs off, unless the launcher already chose CPUs.  PCIe traffic that
// crosses the socket interconnect was the suspect for the end-to-end bandwidth halving from 1 to 2 ranks in round 1.
void CudaBackend::bind_to_gpu_numa_node() {
  if (inproc_ || !ctx_->env.tune.numa_bind) return;
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device_) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
  auto slurp = [](const std::string& path) {
    std::string out;
    if (FILE* f = fopen(path.c_str(), "r")) {
      char buf[4096];
      size_t n = fread(buf, 1, sizeof(buf) - 1, f);
      buf[n] = 0;
      out = buf;
      fclose(f);
    }
    return out;
  };
  const std::string dir = std::string("/sys/bus/pci/devices/") + bus;
  const std::string node_s = slurp(dir + "/numa_node");
  const int node = node_s.empty() ? -1 : atoi(node_s.c_str());
  if (node < 0) return;
  const std::string cpus = slurp("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
  cpu_set_t want, cur, both;
  CPU_ZERO(&want);
  for (const char* p = cpus.c_str(); *p && *p != '\n';) {          // "0-31,64-95"
    char* e = nullptr;
    long a = strtol(p, &e, 10), z = a;
    if (e == p) break;
    if (*e == '-') z = strtol(e + 1, &e, 10);
    for (long c = a; c <= z && c < CPU_SETSIZE; ++c) CPU_SET((int)c, &want);
    p = *e == ',' ? e + 1 : e;
  }
  if (CPU_COUNT(&want) == 0 || sched_getaffinity(0, sizeof(cur), &cur) != 0) return;
  CPU_AND(&both, &want, &cur);
  // already narrowed by the launcher (numactl, taskset, cgroup) to a subset - of this node or of another: leave it alone
  if (CPU_COUNT(&both) == 0 || CPU_COUNT(&cur) <= CPU_COUNT(&want)) {
    MLSLB_LOG(LOG_INFO, "NUMA: GPU %s is on node %d; affinity left as set by the launcher (%d CPUs)", bus, node, CPU_COUNT(&cur));
    return;
  }
  if (sched_setaffinity(0, sizeof(both), &both) == 0) {
    unsigned long mask[16] = {0};
    if (node < (int)(sizeof(mask) * 8)) {
      mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
      syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, sizeof(mask) * 8);
    }
    MLSLB_LOG(LOG_INFO, "NUMA: rank %d bound to node %d (%d CPUs) next to GPU %s", ctx_->rank, node, CPU_COUNT(&both), bus);
  }
}

void CudaBackend::init() {
  Bootstrap* b = ctx_->boot.get();
  inproc_ = b->inproc();
  int ndev = 0;
  MLSLB_CUDA(cudaGetDeviceCount(&ndev));
  MLSLB_ASSERT(ndev > 0, "no CUDA device");
  int lr = ctx_->env.local_rank >= 0 && !inproc_ ? ctx_->env.local_rank : b->rank();
  if (const char* d = getenv("MLSL_DEVICE")) device_ = atoi(d);
  else device_ = lr % ndev;
  set_device();
  bind_to_gpu_numa_node();
  cudaDeviceProp prop;
  MLSLB_CUDA(cudaGetDeviceProperties(&prop, device_));
  sm_count_ = prop.multiProcessorCount;
  slab_bytes_ = (size_t)(ctx_->env.heap_size_gb * 1024.0 * 1024.0 * 1024.0);
  if (slab_bytes_ < kHeaderBytes + ((size_t)64 << 20)) slab_bytes_ = kHeaderBytes + ((size_t)64 << 20);
  if (ctx_->env.check_mem_size) {
    size_t fr = 0, tot = 0;
    MLSLB_CUDA(cudaMemGetInfo(&fr, &tot));
    MLSLB_ASSERT(slab_bytes_ < fr, "MLSL_HEAP_SIZE_GB=%.2f does not fit in free device memory (%.2f GiB)",
                 ctx_->env.heap_size_gb, fr / 1073741824.0);
  }
  // Preferred: VMM slab (file-descriptor shared, NVLS multicast capable).  Fallback: cudaMalloc + CUDA IPC.
  {
    const char* mode = getenv("MLSL_SLAB");
    bool want_vmm = !inproc_ && b->size() > 1 && !(mode && !strcmp(mode, "ipc"));
    if (want_vmm) {
      const double max_gb = ctx_->env.heap_max_gb > 0 ? ctx_->env.heap_max_gb : std::max(32.0, 8.0 * ctx_->env.heap_size_gb);
      vmm_ = vmm_slab_create(b, device_, slab_bytes_, ctx_->env.use_nvls && b->size() <= kMaxDevRanks,
                             (size_t)(max_gb * 1073741824.0));
      if (vmm_.ok) {
        slab_ = vmm_.local;
        slab_bytes_ = vmm_.bytes;
        mc_ = vmm_.mc;
      }
      if (b->rank() == 0)
        MLSLB_LOG(LOG_INFO, "symmetric heap: %s, NVLS multicast: %s%s%s", vmm_.ok ? "VMM (POSIX fd shared)" : "CUDA IPC",
                  mc_ ? "yes" : "no", vmm_.why.empty() ? "" : " - ", vmm_.why.c_str());
    }
  }
  if (!vmm_.ok) {
    cudaError_t e = cudaMalloc((void**)&slab_, slab_bytes_);
    MLSLB_ASSERT(e == cudaSuccess, "cudaMalloc of the %.2f GiB symmetric heap failed: %s (lower MLSL_HEAP_SIZE_GB)",
                 slab_bytes_ / 1073741824.0, cudaGetErrorString(e));
  }
  MLSLB_CUDA(cudaMemset(slab_, 0, kHeaderBytes));
  MLSLB_CUDA(cudaDeviceSynchronize());
  heap_.reset(kHeaderBytes, slab_bytes_ - kHeaderBytes);
  int lo = 0, hi = 0;
  MLSLB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  stream_prio_hi_ = hi;
  streams_.assign(kPadRows, nullptr);
  MLSLB_CUDA(cudaHostAlloc((void**)&err_host_, 64, cudaHostAllocMapped | cudaHostAllocPortable));
  for (int i = 0; i < 16; ++i) err_host_[i] = 0;
  MLSLB_CUDA(cudaHostGetDevicePointer((void**)&err_dev_, (void*)err_host_, 0));
  if (const char* m = getenv("MLSL_STREAM_MODE")) inline_stream_ = !strcmp(m, "inline");
  MLSLB_CUDA(init_kernel_attributes());
  if (ctx_->env.stats || !ctx_->trace_prefix.empty())
    for (int i = 0; i < 128; ++i) {
      cudaEvent_t ev = nullptr;
      MLSLB_CUDA(cudaEventCreate(&ev));
      tevent_pool_.push_back(ev);
    }
  for (int i = 0; i < 256; ++i) {
    cudaEvent_t ev = nullptr;
    MLSLB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    event_pool_.push_back(ev);
  }

  // exchange slab addresses / IPC handles
  const int W = b->size();
  PeerInfo mine;
  memset(&mine, 0, sizeof(mine));
  mine.ptr = (unsigned long long)slab_;
  mine.pid = (int)getpid();
  mine.device = device_;
  if (!inproc_ && !vmm_.ok) MLSLB_CUDA(cudaIpcGetMemHandle(&mine.handle, slab_));
  std::vector<PeerInfo> all(W);
  b->allgather(&mine, all.data(), sizeof(PeerInfo));
  peer_slab_.assign(W, nullptr);
  peer_opened_.assign(W, false);
  for (int p = 0; p < W; ++p) {
    if (p == b->rank()) {
      peer_slab_[p] = slab_;
    } else if (vmm_.ok) {
      peer_slab_[p] = vmm_.peers[p];
    } else if (all[p].pid == mine.pid) {
      peer_slab_[p] = (char*)all[p].ptr;
      if (all[p].device != device_) {
        int can = 0;
        MLSLB_CUDA(cudaDeviceCanAccessPeer(&can, device_, all[p].device));
        MLSLB_ASSERT(can, "device %d cannot access peer device %d", device_, all[p].device);
        cudaError_t pe = cudaDeviceEnablePeerAccess(all[p].device, 0);
        if (pe == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else MLSLB_CUDA(pe);
      }
    } else {
      void* ptr = nullptr;
      cudaError_t oe = cudaIpcOpenMemHandle(&ptr, all[p].handle, cudaIpcMemLazyEnablePeerAccess);
      MLSLB_ASSERT(oe == cudaSuccess, "cudaIpcOpenMemHandle for rank %d failed: %s (are the GPUs peer-capable?)", p,
                   cudaGetErrorString(oe));
      peer_slab_[p] = (char*)ptr;
      peer_opened_[p] = true;
    }
  }
  // ranks sharing one physical device must all be co-resident while they spin on each other: bound the grid
  ranks_per_device_ = 1;
  {
    std::map<int, int> per_dev;
    for (int p = 0; p < W; ++p) per_dev[all[p].device]++;
    // CUDA_VISIBLE_DEVICES may remap ordinals per process; only same-process ranks are known to share
    if (inproc_) for (auto& kv : per_dev) ranks_per_device_ = std::max(ranks_per_device_, kv.second);
  }
  if (const char* v = getenv("MLSL_RANKS_PER_DEVICE")) ranks_per_device_ = std::max(1, atoi(v));
  // Loop-back ranks that share a GPU also share its (at most 32) hardware queues with torch's stream pool.  A kernel
  // that spins for a peer dead-locks as soon as that peer's next kernel is falsely serialised behind it on one queue,
  // so unless told otherwise every collective runs in line on the caller's stream: one stream per rank, program order.
  if (!getenv("MLSL_STREAM_MODE") && ranks_per_device_ > 1) inline_stream_ = true;
  if (vmm_.ok && b->ctl() && b->size() > 1) {
    grow_seen_.assign((size_t)W, 0);
    grow_watcher_ = std::thread([this] { grow_watch(); });
  }
  b->barrier();
  MLSLB_LOG(LOG_DEBUG, "cuda backend up: device %d slab %p (%zu bytes) ranks/device %d", device_, (void*)slab_,
            slab_bytes_, ranks_per_device_);
}

cudaStream_t CudaBackend::stream_for(int row, int lane) {
  std::lock_guard<std::mutex> g(mu_);
  cudaStream_t& s = streams_[row * 2 + lane];
  if (!s) {
    int lo = 0, hi = 0;
    MLSLB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    // communication outranks compute so a spinning collective is never starved of SMs by queued compute kernels
    MLSLB_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, lane ? hi : std::min(lo, hi + 1)));
  }
  return s;
}

// Channel ("endpoint") count as a pure function of the message size: every member must pick the same grid.
// `bytes` = what ONE rank produces (its slice of an all-reduce, its shard of a reduce-scatter, what it pulls in a
// gather).  NVLink round trips are ~2-3 us, so the only way to move data fast is to have it all in flight at once:
// one 16-byte vector per thread until the grid limit is reached, only then more vectors per thread.
int CudaBackend::pick_channels(size_t bytes) const {
  size_t want = ceil_div(std::max<size_t>(bytes, 1), (size_t)kCommThreads * 16);
  int cmax = ctx_->env.num_channels > 0 ? ctx_->env.num_channels : 96;   // MLSL_NUM_CHANNELS (<= 128)
  int cap = std::max(1, sm_count_ / std::max(1, ranks_per_device_));
  cmax = std::min(cmax, std::min(cap, kMaxChannels));
  if (ctx_->env.max_short_msg && bytes <= ctx_->env.max_short_msg * 4) return 1;
  return (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)cmax));
}

DevComm CudaBackend::make_comm(const ProcessGroup& g, int lane) const {
  DevComm dc;
  memset(&dc, 0, sizeof(dc));
  dc.nranks = g.size();
  dc.me = g.idx;
  const size_t row = g.row >= 0 ? (size_t)g.row : (size_t)kMaxGroupRows - 1;   // self groups: reserved last row
  dc.pad_off = (unsigned)((row * 2 + lane) * kPadRowBytes);
  dc.seq_off = (unsigned)(kSeqBase + (row * 2 + lane) * kSeqRowBytes);
  dc.timeout_ns = ctx_->env.watchdog_sec > 0 ? (unsigned long long)ctx_->env.watchdog_sec * 1000000000ull : 0ull;
  dc.err = err_dev_;
  for (int i = 0; i < g.size(); ++i) dc.slab[i] = peer_slab_[g.members[i]];
  dc.ll_off = (lane == 0 && g.row >= 0 && g.row < kLLRows) ? (unsigned)(kLLBase + (size_t)g.row * kLLRowBytes) : 0u;
  const bool mid = lane == 0 && g.row >= 0 && g.row < kMidRows;
  dc.mid_off = mid ? (unsigned)(kMidBase + (size_t)g.row * kMidRowBytes) : 0u;
  dc.mid_seq_off = mid ? (unsigned)(kMidSeqBase + (size_t)g.row * kMidSeqRowBytes) : 0u;
  // NVLS only for groups spanning every rank in world order (the multicast object covers exactly those devices)
  dc.mc = nullptr;
  // (and only from 4 ranks up: per direction NVLS moves S(1+1/N) bytes, the peer-to-peer kernel 2S(N-1)/N)
  const int nvls_min = (int)ctx_->env.tune.nvls_min_ranks;
  if (mc_ && g.size() == ctx_->world && g.size() >= nvls_min) {
    bool ident = true;
    for (int i = 0; i < g.size(); ++i) ident &= g.members[i] == i;
    if (ident) {
      dc.mc = mc_;
      dc.mc_bytes = vmm_.bytes;      // the multicast object covers the initial slab only
    }
  }
  return dc;
}

void CudaBackend::check_error(const char* what) {
  int code = *err_host_;
  if (code != 0) {
    ctx_->boot->poison(ctx_->rank);
    const unsigned long long off = ((unsigned long long)(unsigned)err_host_[4] << 32) | (unsigned)err_host_[3];
    const unsigned long long seen = ((unsigned long long)(unsigned)err_host_[6] << 32) | (unsigned)err_host_[5];
    MLSLB_ASSERT(false,
                 "device watchdog: %s: a peer never arrived (error code %d, reported by group index %d; channel %d, "
                 "waiting on peer index %d, signal word at slab offset 0x%llx held 0x%llx)",
                 what, code, code - 1000, err_host_[1], err_host_[2], off, seen);
  }
  if (ctx_->boot->poisoned()) MLSLB_ASSERT(false, "job poisoned by rank %d", (int)ctx_->boot->poisoned() - 1);
}

void CudaBackend::drop_stages(CudaReqState* st, bool) {
  for (auto& s : st->stages) free(s.slab);
  st->stages.clear();
}

void CudaBackend::finish(CommRequest& r, CudaReqState* st) {
  harvest(r, st, true);
  st->inflight = false;
  drop_stages(st, true);
  check_error(opkind_name(r.desc.kind));
}

void CudaBackend::launch(CommRequest& r) {
  set_device();
  auto* st = (CudaReqState*)r.backend_state;
  MLSLB_ASSERT(st != nullptr, "request was not prepared");
  const CommDesc& d = r.desc;
  ProcessGroup* g = d.group;
  const bool solo = !g || g->size() <= 1;
  cudaStream_t s = inline_stream_ ? ustream() : stream_for(solo || g->row < 0 ? kMaxGroupRows - 1 : g->row, r.lane);
  st->stream = s;
  if (!inline_stream_) MLSLB_CUDA(cudaStreamWaitEvent(s, st->ready, 0));
  // NVTX range per collective launch (header-only NVTX3: a no-op unless a profiler is attached), SURVEY 5.1
  const bool nvtx = ctx_->env.tune.nvtx != 0;
  if (nvtx) nvtxRangePushA(opkind_name(d.kind));
  struct NvtxPop {
    bool on;
    ~NvtxPop() { if (on) nvtxRangePop(); }
  } nvtx_pop{nvtx};
  if (!solo) loopback_rendezvous(*g, r.lane);
  const bool timing = want_timing();
  if (timing) {
    harvest(r, st, false);           // the previous run of a persistent request, if nobody asked in between
    if (!st->timed) {
      if (!st->t0) {
        st->t0 = take_timing_event();
        st->t1 = take_timing_event();
      }
      MLSLB_CUDA(cudaEventRecord(st->t0, s));
    }
  }
  launch_single(r, st, s);
  if (timing && !st->timed) {
    MLSLB_CUDA(cudaEventRecord(st->t1, s));
    st->timed = true;
  }
  st->recorded = false;
  if (!(eventless() && st->stages.empty())) {
    ensure_events(st);
    MLSLB_CUDA(cudaEventRecord(st->done, s));
    st->recorded = true;
  }
  st->inflight = true;
  r.state.store(CommRequest::LAUNCHED, std::memory_order_release);
}

void CudaBackend::launch_single(CommRequest& r, CudaReqState* st, cudaStream_t s) {
  const CommDesc& d = r.desc;
  ProcessGroup* g = d.group;
  const size_t es = dtype_size(d.dtype);
  const size_t n = d.count;
  // ---- single-rank groups: local semantics, no peers -------------------------------------------------------
  // MLSL_FORCE_KERNEL_SOLO=1: run a 1-rank group through the real collective kernels (handshake with itself,
  // pull from / push to its own buffers).  Lets ncu profile the kernels on one GPU - under the profiler kernels
  // are serialised, so ranks that wait for each other can never be captured.
  const bool force_solo = ctx_->env.tune.force_kernel_solo != 0;
  // (self groups have no signal row of their own: make_comm gives them the reserved last row)
  if (!g || g->size() <= 1 ? !((force_solo && g) || d.kind == OpKind::FUSED_UPDATE || d.kind == OpKind::GEMM_RS || d.kind == OpKind::AG_GEMM) : false) {
    size_t bytes = 0;
    switch (d.kind) {
      case OpKind::ALLREDUCE: case OpKind::REDUCE: case OpKind::REDUCE_SCATTER: case OpKind::ALLGATHER:
      case OpKind::GATHER: case OpKind::SCATTER: case OpKind::ALLTOALL: case OpKind::ALLGATHERV:
        bytes = n * es;
        break;
      default: break;
    }
    // host-resident all-reduce on one rank (the 1-GPU end-to-end case): same chunked H2D / D2H pipeline as with
    // peers, so the two PCIe directions overlap instead of one device-driven host-to-host copy
    if (d.kind == OpKind::ALLREDUCE && !d.compress && r.send != r.recv && launch_host_pipelined(r, DevComm{}, s, true)) return;
    const bool reducing = d.kind == OpKind::ALLREDUCE || d.kind == OpKind::REDUCE_SCATTER || d.kind == OpKind::REDUCE;
    if (reducing && bytes && r.recv && r.recv != r.send && owns(r.recv, bytes) && owns(r.send, bytes)) {
      // one kernel: copy with the scale epilogue fused
      MLSLB_CUDA(launch_scale_copy(d.dtype, r.recv, r.send, n, d.scale, s));
      return;
    }
    if (bytes && r.recv && r.recv != r.send) MLSLB_CUDA(cudaMemcpyAsync(r.recv, r.send, bytes, cudaMemcpyDefault, s));
    if ((d.kind == OpKind::ALLTOALLV || d.kind == OpKind::SENDRECV_LIST) && !d.send_counts.empty() && d.send_counts[0])
      MLSLB_CUDA(cudaMemcpyAsync((char*)r.recv + d.recv_offsets[0] * es, (char*)r.send + d.send_offsets[0] * es,
                                 d.send_counts[0] * es, cudaMemcpyDefault, s));
    if (d.scale != 1.0f && r.recv && (d.kind == OpKind::ALLREDUCE || d.kind == OpKind::REDUCE_SCATTER))
      MLSLB_CUDA(launch_scale(d.dtype, r.recv, n, d.scale, s));
    return;
  }
  const int P = g->size(), me = g->idx;
  DevComm dc = make_comm(*g, r.lane);
  const bool trace = ctx_->env.tune.trace_launch != 0;
  if (trace) {
    fprintf(stderr, "[launch %.6f] r%d %s row %d lane %d idx %d/%d count %zu send %p recv %p stream %p\n", now_ns() * 1e-9,
            ctx_->rank, opkind_name(d.kind), g->row, r.lane, me, P, n, r.send, r.recv, (void*)s);
    fflush(stderr);
  }

  // ---- tiny all-reduce: low-latency push path -------------------------------------------------------------------
  // The choice of kernel must be the same on every member, so it only looks at (kind, size, group).  The kernel reads
  // and writes any device-accessible memory at any alignment (torch tensors are used in place, no staging copy); a
  // buffer the GPU cannot address (pageable host memory) goes through a small slab scratch block first.
  const bool ll_enabled = ctx_->env.tune.ll != 0;
  const size_t mid_max = std::min<size_t>(kMidMaxBytes, (size_t)std::max<long>(0, ctx_->env.tune.mid_max_kb) << 10);
  const bool take_ll = d.kind == OpKind::ALLREDUCE && !d.compress && ll_enabled && dc.ll_off && n > 0 && n * es <= kLLMaxBytes;
  const bool take_mid = !take_ll && d.kind == OpKind::ALLREDUCE && !d.compress && ll_enabled && dc.mid_off && n > 0 &&
                        n * es <= mid_max;
  if (take_ll || take_mid) {
    const size_t bytes = n * es;
    auto dev_ok = [&](const void* p) { return owns(p, bytes) || is_device_pointer(p); };
    const void* sp = r.send;
    void* rp = r.recv;
    const bool in_place = r.send == r.recv;
    if (!dev_ok(rp)) {
      StageBuf sb{r.recv, alloc(bytes, 256), bytes, true};
      if (in_place) MLSLB_CUDA(cudaMemcpyAsync(sb.slab, r.recv, bytes, cudaMemcpyDefault, s));
      st->stages.push_back(sb);
      rp = sb.slab;
      if (in_place) sp = rp;
    }
    if (!in_place && !dev_ok(sp)) {
      StageBuf sb{(void*)r.send, alloc(bytes, 256), bytes, false};
      MLSLB_CUDA(cudaMemcpyAsync(sb.slab, r.send, bytes, cudaMemcpyDefault, s));
      st->stages.push_back(sb);
      sp = sb.slab;
    }
    if (take_ll) {
      MLSLB_CUDA(launch_allreduce_ll(dc, d.dtype, d.rop, sp, rp, n, d.scale, s));
    } else {
      // one-shot while the (P-1) copies every rank sends and receives stay small, two-shot (reduce-scatter + all-gather,
      // both flag-in-data) above; the grid is fixed (every launch counts on all kMidCtas launch counters)
      const bool two_shot = (size_t)(P - 1) * bytes > ((size_t)std::max<long>(0, ctx_->env.tune.mid_oneshot_kb) << 10);
      MLSLB_CUDA(launch_allreduce_mid(dc, d.dtype, d.rop, sp, rp, n, d.scale, two_shot, kMidCtas, s));
    }
    for (auto& sb : st->stages)
      if (sb.copy_out) copy_out(sb, s);
    return;
  }

  // ---- host-resident all-reduce: chunked H2D -> kernel -> D2H pipeline instead of staging the whole message ------
  if (d.kind == OpKind::ALLREDUCE && !d.compress && launch_host_pipelined(r, dc, s)) return;

  // ---- stage foreign buffers through the slab ------------------------------------------------------------------
  size_t sbytes = r.send_bytes(), rbytes = r.recv_bytes();
  if ((d.kind == OpKind::REDUCE || d.kind == OpKind::GATHER) && me != (int)d.root) rbytes = 0;
  if (d.kind == OpKind::SCATTER && me != (int)d.root) sbytes = 0;
  if (d.kind == OpKind::FUSED_UPDATE) rbytes = n * P * dtype_size(d.has_out_dtype ? d.out_dtype : d.dtype);
  if (d.kind == OpKind::AG_GEMM) rbytes = 0;   // Y and the gathered X are local: any device memory; the shard is staged if foreign
  if (d.kind == OpKind::ALLTOALL && d.strided.on) {   // whole unpacked tensors are the buffers
    sbytes = d.strided.src_total;
    if (d.strided.dst_direct) rbytes = d.strided.dst_total;
  }
  if (d.kind == OpKind::GEMM_RS) {
    sbytes = 0;   // A and W are only read by this rank's own TMA loads: any device memory will do
    rbytes = (size_t)d.gemm.M / P * d.gemm.N * (d.has_out_dtype && d.out_dtype == DType::F32 ? 4 : 2);
  }
  auto stage = [&](void* user, size_t bytes, bool copy_in, bool copy_out) -> char* {
    if (!user || bytes == 0 || owns(user, bytes)) return (char*)user;
    StageBuf sb{user, alloc(bytes, 256), bytes, copy_out};
    if (copy_in) MLSLB_CUDA(cudaMemcpyAsync(sb.slab, user, bytes, cudaMemcpyDefault, s));
    st->stages.push_back(sb);
    return (char*)sb.slab;
  };
  char* R = nullptr;
  char* S = nullptr;
  const bool in_place = r.send == r.recv;
  const bool recv_needs_input = in_place || d.kind == OpKind::BCAST || d.kind == OpKind::ALLGATHER ||
                                d.kind == OpKind::ALLGATHERV || d.kind == OpKind::FUSED_UPDATE;
  if (r.recv && rbytes) R = stage(r.recv, rbytes, recv_needs_input, true);
  if (r.send && sbytes) {
    char* us = (char*)r.send;
    char* ur = (char*)r.recv;
    if (ur && rbytes && us >= ur && us + sbytes <= ur + rbytes) S = R + (us - ur);
    else S = stage(r.send, sbytes, true, false);
  }
  // ops whose device kernel cannot run in place get a private receive area
  bool alias_fix = false;
  if ((d.kind == OpKind::ALLTOALL || d.kind == OpKind::ALLTOALLV || d.kind == OpKind::SENDRECV_LIST) && S && R == S)
    alias_fix = true;
  if (d.kind == OpKind::REDUCE_SCATTER && S && R >= S && R < S + sbytes) alias_fix = true;
  char* Rfinal = R;
  if (alias_fix) {
    StageBuf sb{R, alloc(rbytes, 256), rbytes, false};
    st->stages.push_back(sb);
    R = (char*)sb.slab;
  }
  const unsigned long long so = S ? (unsigned long long)(S - slab_) : 0ull;
  const unsigned long long ro = R ? (unsigned long long)(R - slab_) : 0ull;
  // work one rank performs, in output bytes (see pick_channels)
  size_t work = r.msg_bytes();
  switch (d.kind) {
    case OpKind::ALLREDUCE: work = ceil_div(n * es, (size_t)P); break;
    case OpKind::REDUCE_SCATTER: case OpKind::REDUCE: case OpKind::SCATTER: work = n * es; break;
    case OpKind::FUSED_UPDATE: work = n * 4; break;
    default: break;   // gather-like: everything a rank pulls
  }
  int ch = pick_channels(work);
  // the *v exchanges move a different number of bytes on every rank, but the grid (= the channels that shake hands) has
  // to be the same everywhere: a fixed grid for them
  if (d.kind == OpKind::ALLTOALLV || d.kind == OpKind::SENDRECV_LIST)
    ch = std::min(32, std::min(kMaxChannels, std::max(1, sm_count_ / std::max(1, ranks_per_device_))));

  switch (d.kind) {
    case OpKind::BARRIER: MLSLB_CUDA(launch_barrier(dc, s)); break;
    case OpKind::ALLREDUCE: {
      const bool can_quant = d.compress && d.dtype == DType::F32 && d.rop == RedOp::SUM && ((so | ro) & 15ull) == 0;
      if (can_quant) {
        size_t sb = allreduce_quant_stage_bytes(n);
        if (!st->qstage) st->qstage = alloc(sb, 256);
        if (st->residual_elems != n) {
          // from the slab, not cudaMalloc/cudaFree: those may synchronise the device while a peer rank of this
          // process is spinning inside a collective that needs THIS launch to make progress
          if (st->residual) free(st->residual);
          st->residual = (float*)alloc(std::max<size_t>(n, 1) * sizeof(float), 256);
          MLSLB_CUDA(cudaMemsetAsync(st->residual, 0, n * sizeof(float), s));
          st->residual_elems = n;
        }
        MLSLB_CUDA(launch_allreduce_quant(dc, so, ro, (unsigned long long)((char*)st->qstage - slab_), st->residual, n,
                                          d.scale, quant_channels(n), ctx_->env.tune.quant_mx != 0, s));
      } else {
        // very large messages go out as pipelined chunks (reference MLSL_LARGE_MSG_SIZE_MB / _CHUNKS,
        // src/comm_ep.cpp:645-656) so a higher-priority collective can slip in between them
        size_t chunks = 1;
        if (ctx_->env.large_msg_mb && n * es >= ctx_->env.large_msg_mb * (size_t)1048576 && ctx_->env.large_msg_chunks > 1 &&
            ctx_->env.msg_priority)
          chunks = (size_t)ctx_->env.large_msg_chunks;
        // The kernel walks the message pass by pass with all ranks in step (kernels.cu), so one launch covers any size;
        // MLSL_NVLS_CHUNK_MB > 0 brings back the round-1 split of giant multicast messages into several launches.
        const size_t nvls_chunk = (size_t)std::max<long>(0, ctx_->env.tune.nvls_chunk_mb) << 20;
        if (dc.mc && nvls_chunk && n * es >= nvls_chunk + nvls_chunk / 2) chunks = std::max(chunks, ceil_div(n * es, nvls_chunk));
        size_t per = round_up(ceil_div(n, chunks), 256);
        const int unroll = (int)ctx_->env.tune.ar_unroll;
        for (size_t off = 0; off < n; off += per) {
          size_t cnt = std::min(per, n - off);
          const int chn = ar_channels(ceil_div(cnt * es, (size_t)P));
          // hybrid (multicast + peer-to-peer at once) only where it can pay: big messages on a full grid
          int p2p_cta = 0;
          float p2p_frac = 0.f;
          if (dc.mc && ctx_->env.tune.ar_p2p_pct > 0 && cnt * es >= ((size_t)32 << 20) && chn >= 8) {
            p2p_cta = std::max(1, (int)(chn * ctx_->env.tune.ar_p2p_cta_pct / 100));
            p2p_frac = (float)ctx_->env.tune.ar_p2p_pct / 100.f;
          }
          MLSLB_CUDA(launch_allreduce(dc, d.dtype, d.rop, so + off * es, ro + off * es, cnt, d.scale, chn, unroll, p2p_cta, p2p_frac, s));
        }
      }
      break;
    }
    case OpKind::REDUCE_SCATTER:
    case OpKind::REDUCE: {
      DevComm rdc = dc;
      if (!(ctx_->env.tune.nvls_collectives & 1)) rdc.mc = nullptr;   // multimem.ld_reduce flavour off
      if (d.kind == OpKind::REDUCE_SCATTER)
        MLSLB_CUDA(launch_reduce_pull(rdc, d.dtype, d.rop, so, ro, (size_t)me * n, n, d.scale, true, ch, s));
      else
        MLSLB_CUDA(launch_reduce_pull(rdc, d.dtype, d.rop, so, ro, 0, n, d.scale, me == (int)d.root, ch, s));
      break;
    }
    case OpKind::ALLGATHER:
    case OpKind::ALLGATHERV:
    case OpKind::BCAST:
    case OpKind::ALLTOALL:
    case OpKind::ALLTOALLV:
    case OpKind::SENDRECV_LIST:
    case OpKind::GATHER:
    case OpKind::SCATTER: {
      CopyPlan plan;
      memset(&plan, 0, sizeof(plan));
      plan.elem_size = (int)es;
      auto add = [&](int peer, unsigned long long src, unsigned long long dst, unsigned long long bytes, int aux) {
        CopySeg& sg = plan.seg[plan.nseg++];
        sg.peer = peer;
        sg.src_off = src;
        sg.dst_off = dst;
        sg.bytes = bytes;
        sg.use_aux = aux;
      };
      unsigned long long pub_send = so;
      switch (d.kind) {
        case OpKind::ALLGATHER:
          for (int p = 0; p < P; ++p) add(p, 0, (unsigned long long)p * n * es, n * es, 0);
          break;
        case OpKind::ALLGATHERV: {
          unsigned long long off = 0;
          for (int p = 0; p < P; ++p) {
            add(p, 0, off * es, d.recv_counts[p] * es, 0);
            off += d.recv_counts[p];
          }
          break;
        }
        case OpKind::BCAST:
          pub_send = ro;   // one buffer: publish it as the source
          if (me != (int)d.root) add((int)d.root, 0, 0, n * es, 0);
          break;
        case OpKind::ALLTOALL:
          if (d.strided.on) {
            const CommDesc::Strided& sd = d.strided;
            for (int p = 0; p < P; ++p) {
              add(p, sd.src_off, sd.dst_direct ? sd.dst_off[(size_t)p] : (unsigned long long)p * n * es, sd.rows * sd.row_bytes, 0);
              CopySeg& sg = plan.seg[plan.nseg - 1];
              sg.rows = sd.rows;
              sg.row_bytes = sd.row_bytes;
              sg.src_stride = sd.src_stride;
              sg.dst_stride = sd.dst_direct ? sd.dst_stride : sd.row_bytes;
            }
            break;
          }
          for (int p = 0; p < P; ++p) add(p, (unsigned long long)me * n * es, (unsigned long long)p * n * es, n * es, 0);
          break;
        case OpKind::ALLTOALLV:
        case OpKind::SENDRECV_LIST:
          for (int p = 0; p < P; ++p) {
            plan.aux_out[p] = d.send_offsets[p];
            if (d.recv_counts[p]) add(p, 0, d.recv_offsets[p] * es, d.recv_counts[p] * es, 1);
          }
          break;
        case OpKind::GATHER:
          if (me == (int)d.root)
            for (int p = 0; p < P; ++p) add(p, 0, (unsigned long long)p * n * es, n * es, 0);
          break;
        case OpKind::SCATTER:
          add((int)d.root, (unsigned long long)me * n * es, 0, n * es, 0);
          break;
        default: break;
      }
      // NVLS push where it pays: the root of a bcast sends the message once instead of P-1 times (all-gather: opt-in,
      // bit 4 - every byte still has to arrive at every member, the push only saves outbound traffic)
      const long nv = ctx_->env.tune.nvls_collectives;
      if (dc.mc && d.kind == OpKind::BCAST && (nv & 2)) {
        plan.mc_mode = 1;
        plan.mc_root = (int)d.root;
        plan.mc_bytes = n * es;
      } else if (dc.mc && d.kind == OpKind::ALLGATHER && (nv & 4)) {
        plan.mc_mode = 2;
        plan.mc_bytes = n * es;
      }
      if (d.kind == OpKind::ALLTOALL) plan.pairs_concurrent = ctx_->env.alltoall_split == 0;
      if (d.kind == OpKind::ALLTOALLV || d.kind == OpKind::SENDRECV_LIST) plan.pairs_concurrent = ctx_->env.alltoallv_split == 0;
      // big messages: the copy engine (cp.async.bulk rings) moves the segments instead of the threads
      unsigned long long pulled = 0;
      for (int i = 0; i < plan.nseg; ++i) pulled += plan.seg[i].bytes;
      const long bk = ctx_->env.tune.bulk_copy_kb;
      const bool bulk = bk > 0 && pulled >= ((unsigned long long)bk << 10) && !plan.pairs_concurrent &&   // the ring deals pieces itself
                        !(d.kind == OpKind::ALLTOALL && d.strided.on);                                  // rectangles: the threads copy
      MLSLB_CUDA(launch_pull_copy(dc, plan, pub_send, ro, ch, bulk, s));
      break;
    }
    case OpKind::FUSED_UPDATE: {
      const CommDesc::FusedUpdate& f = d.fused;
      FusedUpdateArgs a;
      a.optimizer = f.optimizer;
      a.lr = f.lr;
      a.momentum = f.momentum;
      a.beta1 = f.beta1;
      a.beta2 = f.beta2;
      a.eps = f.eps;
      a.weight_decay = f.weight_decay;
      a.grad_scale = d.scale;
      a.bc1 = f.optimizer == 1 ? 1.f - powf(f.beta1, (float)f.step) : 1.f;
      a.bc2 = f.optimizer == 1 ? 1.f - powf(f.beta2, (float)f.step) : 1.f;
      a.master = (float*)f.master;
      a.state1 = (float*)f.state1;
      a.state2 = (float*)f.state2;
      MLSLB_CUDA(launch_fused_update(dc, d.dtype, d.has_out_dtype ? d.out_dtype : d.dtype, so, ro, n, a, ch, s));
      break;
    }
    case OpKind::AG_GEMM: {
      const CommDesc::GemmRs& gm = d.gemm;
      const char* why = ag_gemm_check(gm.M, gm.N, gm.K, P);
      MLSLB_ASSERT(why == nullptr, "AllGatherGemm(M=%d, N=%d, K=%d, P=%d): %s", gm.M, gm.N, gm.K, P, why);
      MLSLB_ASSERT(is_device_pointer(gm.w) && is_device_pointer(d.gathered) && is_device_pointer(r.recv),
                   "AllGatherGemm: W, the gathered buffer and the output must be device memory");
      const int cap = std::min(std::max(1, sm_count_ / std::max(1, ranks_per_device_)), kMaxChannels);
      int copy_ctas = 0, total_ctas = 0;
      ag_gemm_grid(gm.M, gm.N, P, cap, &copy_ctas, &total_ctas);
      if (!st->qstage) {   // launch counters + tile flags: zero once, the kernel keeps them consistent afterwards
        const size_t sb = ag_gemm_scratch_bytes(gm.M, kMaxChannels);
        st->qstage = alloc(sb, 256);
        MLSLB_CUDA(cudaMemsetAsync(st->qstage, 0, sb, s));
      }
      const bool out32 = d.has_out_dtype && d.out_dtype == DType::F32;
      MLSLB_CUDA(launch_ag_gemm(dc, so, gm.w, d.gathered, r.recv, out32, gm.M, gm.N, gm.K, copy_ctas, total_ctas, st->qstage, s));
      break;
    }
    case OpKind::GEMM_RS: {
      const CommDesc::GemmRs& gm = d.gemm;
      const char* why = gemm_rs_check(gm.M, gm.N, gm.K, P);
      MLSLB_ASSERT(why == nullptr, "GemmReduceScatter(M=%d, N=%d, K=%d, P=%d): %s", gm.M, gm.N, gm.K, P, why);
      MLSLB_ASSERT(d.dtype == DType::BF16, "GemmReduceScatter: inputs must be bf16");
      size_t sb = gemm_rs_stage_bytes(gm.M, gm.N);
      if (!st->qstage) st->qstage = alloc(sb, 1024);
      const bool out32 = d.has_out_dtype && d.out_dtype == DType::F32;
      const int gcap = std::min(std::max(1, sm_count_ / std::max(1, ranks_per_device_)), kMaxChannels);
      // MLSL_GEMM_2CTA=1: experimental cta_group::2 kernel (256 x 256 tiles on CTA pairs); same choice on every rank
      const bool two_cta = ctx_->env.tune.gemm_2cta != 0;
      if (two_cta && gcap >= 2 && gemm_rs2_check(gm.M, gm.N, gm.K, P) == nullptr) {
        const int gch = gemm_rs2_channels(gm.M, gm.N, gcap);
        MLSLB_CUDA(launch_gemm_rs2(dc, gm.a, gm.w, (unsigned long long)((char*)st->qstage - slab_), R, out32, gm.M, gm.N, gm.K, gch, s));
      } else {
        const int gch = gemm_rs_channels(gm.M, gm.N, gcap);
        MLSLB_CUDA(launch_gemm_rs(dc, gm.a, gm.w, (unsigned long long)((char*)st->qstage - slab_), R, out32, gm.M, gm.N, gm.K, gch, s));
      }
      break;
    }
  }

  // ---- copy results out of the staging buffers -------------------------------------------------------------------
  if (alias_fix) MLSLB_CUDA(cudaMemcpyAsync(Rfinal, R, rbytes, cudaMemcpyDefault, s));
  for (auto& sb : st->stages)
    if (sb.copy_out && sb.user == r.recv) copy_out(sb, s);
}

// End-to-end path for buffers that live in HOST memory (the reference's only kind of buffer): the message is cut
// into chunks that flow through three slab buffers - chunk i+1 is on its way up over PCIe while chunk i is being
// all-reduced over NVLink and chunk i-1 travels back down, so a step costs ~max(H2D, D2H) instead of their sum.
// Every rank derives the same chunking from the message size, so the per-chunk kernels pair up across ranks.
bool CudaBackend::launch_host_pipelined(CommRequest& r, const DevComm& dc, cudaStream_t s, bool solo) {
  const CommDesc& d = r.desc;
  const size_t es = dtype_size(d.dtype), bytes = d.count * es;
  if (bytes < ((size_t)64 << 20) || !r.send || !r.recv) return false;
  if (!ctx_->env.tune.host_pipeline) return false;
  cudaPointerAttributes as, ar;
  if (cudaPointerGetAttributes(&as, r.send) != cudaSuccess || cudaPointerGetAttributes(&ar, r.recv) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  auto is_host = [](const cudaPointerAttributes& a) { return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeUnregistered; };
  // the decision must be identical on every rank: only taken when BOTH buffers are host memory (SPMD programs pass
  // the same kind of buffer everywhere)
  if (!is_host(as) || !is_host(ar)) return false;
  if (!pipe_buf_[0]) {
    pipe_bufs_ = (int)std::min<long>(kPipeBufs, std::max<long>(2, ctx_->env.tune.pipe_bufs));
    pipe_chunk_ = (size_t)std::max<long>(1, ctx_->env.tune.pipe_chunk_mb) << 20;
    for (int b = 0; b < pipe_bufs_; ++b) {
      pipe_buf_[b] = (char*)alloc(pipe_chunk_, 4096);
      pipe_h2d_[b] = take_event();
      pipe_ar_[b] = take_event();
      pipe_d2h_[b] = take_event();
    }
    pipe_start_ = take_event();
    int lo = 0, hi = 0;
    MLSLB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    MLSLB_CUDA(cudaStreamCreateWithPriority(&h2d_stream_, cudaStreamNonBlocking, hi));
    MLSLB_CUDA(cudaStreamCreateWithPriority(&d2h_stream_, cudaStreamNonBlocking, hi));
  }
  const int P = solo ? 1 : dc.nranks;
  MLSLB_CUDA(cudaEventRecord(pipe_start_, s));
  MLSLB_CUDA(cudaStreamWaitEvent(h2d_stream_, pipe_start_, 0));
  const size_t chunk_elems = pipe_chunk_ / es;
  int i = 0;
  for (size_t off = 0; off < d.count; off += chunk_elems, ++i) {
    const int b = i % pipe_bufs_;
    const size_t cnt = std::min(chunk_elems, d.count - off);
    if (pipe_used_[b]) MLSLB_CUDA(cudaStreamWaitEvent(h2d_stream_, pipe_d2h_[b], 0));   // buffer drained by its last user
    MLSLB_CUDA(cudaMemcpyAsync(pipe_buf_[b], (const char*)r.send + off * es, cnt * es, cudaMemcpyHostToDevice, h2d_stream_));
    MLSLB_CUDA(cudaEventRecord(pipe_h2d_[b], h2d_stream_));
    MLSLB_CUDA(cudaStreamWaitEvent(s, pipe_h2d_[b], 0));
    const unsigned long long o = (unsigned long long)(pipe_buf_[b] - slab_);
    if (solo) {   // single-rank group: nothing to reduce, the chunk only passes through the GPU (scaled if asked for)
      if (d.scale != 1.0f) MLSLB_CUDA(launch_scale(d.dtype, pipe_buf_[b], cnt, d.scale, s));
    } else {
      MLSLB_CUDA(launch_allreduce(dc, d.dtype, d.rop, o, o, cnt, d.scale, ar_channels(ceil_div(cnt * es, (size_t)P)),
                                  (int)ctx_->env.tune.ar_unroll, 0, 0.f, s));
    }
    MLSLB_CUDA(cudaEventRecord(pipe_ar_[b], s));
    MLSLB_CUDA(cudaStreamWaitEvent(d2h_stream_, pipe_ar_[b], 0));
    MLSLB_CUDA(cudaMemcpyAsync((char*)r.recv + off * es, pipe_buf_[b], cnt * es, cudaMemcpyDeviceToHost, d2h_stream_));
    MLSLB_CUDA(cudaEventRecord(pipe_d2h_[b], d2h_stream_));
    pipe_used_[b] = true;
  }
  // completion of the request = the last copies down
  for (int b = 0; b < pipe_bufs_; ++b)
    if (pipe_used_[b]) MLSLB_CUDA(cudaStreamWaitEvent(s, pipe_d2h_[b], 0));
  return true;
}

}  // namespace

std::unique_ptr<Backend> make_cuda_backend(RankContext* ctx) {
  if (!cuda_backend_available()) return nullptr;
  return std::unique_ptr<Backend>(new CudaBackend(ctx));
}

bool cuda_backend_available() {
  static int cached = -1;
  if (cached < 0) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) cudaGetLastError();
    cached = (e == cudaSuccess && n > 0) ? 1 : 0;
  }
  return cached == 1;
}

void cuda_fill_sysinfo(SysInfo& s) {
  if (!cuda_backend_available()) return;
  int n = 0, dev = 0;
  cudaGetDeviceCount(&n);
  cudaGetDevice(&dev);
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return;
  s.gpus = n;
  s.sms = p.multiProcessorCount;
  s.gpu_name = p.name;
  s.cc_major = p.major;
  s.cc_minor = p.minor;
  s.peer_access = n <= 1;
  if (n > 1) {
    int can = 0;
    cudaDeviceCanAccessPeer(&can, dev, (dev + 1) % n);
    s.peer_access = can != 0;
  }
  // multicast (NVLS) support: CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED = 132, queried through the driver entry point
  typedef int (*cuDeviceGetAttribute_t)(int*, int, int);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuDeviceGetAttribute", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn) {
    int v = 0;
    if (((cuDeviceGetAttribute_t)fn)(&v, 132, dev) == 0) s.multicast = v != 0;
  } else {
    cudaGetLastError();
  }
}

}  // namespace mlslb
