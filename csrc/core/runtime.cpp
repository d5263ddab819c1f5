#include "runtime.hpp"

#include <pthread.h>
#include <sched.h>
#include <unistd.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>

#include "log.hpp"

namespace mlslb {

// ============================================================================================================
// CommRequest
// ============================================================================================================
CommRequest::CommRequest(RankContext* c, DType dt, int64_t uid, CommDesc::CompType ct) : ctx(c) {
  desc.dtype = dt;
  desc.op_uid = uid;
  desc.comp_type = ct;
}

CommRequest::~CommRequest() {
  if (ctx && ctx->backend && backend_state) ctx->backend->release(*this);
}

// Buffer-size rules.  Same contract as the reference's CommRequestImpl::Setup (src/comm_ep.cpp:568-766): the
// comm buffer is [send region | recv region] for ops we run out of place and just the message for in-place ops.
// We run ReduceScatter / AlltoAll(v) / Gather out of place (what the reference does whenever endpoints are
// active) because a peer-memory kernel cannot overwrite a slice other GPUs are still reading.
void CommRequest::setup() {
  const size_t dt = dtype_size(desc.dtype);
  const size_t P = desc.group ? (size_t)desc.group->size() : 1;
  const size_t n = desc.count;
  oop_default_ = false;
  switch (desc.kind) {
    case OpKind::BARRIER:
      send_bytes_ = recv_bytes_ = buf_bytes_ = msg_bytes_ = 0;
      break;
    case OpKind::BCAST:
    case OpKind::REDUCE:
    case OpKind::ALLREDUCE:
      send_bytes_ = recv_bytes_ = buf_bytes_ = msg_bytes_ = n * dt;
      break;
    case OpKind::ALLGATHER:
      send_bytes_ = n * dt;
      recv_bytes_ = n * P * dt;
      buf_bytes_ = recv_bytes_;          // in place: rank r's contribution lives at r*n
      msg_bytes_ = recv_bytes_;
      break;
    case OpKind::ALLGATHERV: {
      size_t tot = 0;
      for (size_t c : desc.recv_counts) tot += c;
      send_bytes_ = n * dt;
      recv_bytes_ = buf_bytes_ = msg_bytes_ = tot * dt;
      break;
    }
    case OpKind::REDUCE_SCATTER:
      send_bytes_ = n * P * dt;
      recv_bytes_ = n * dt;
      buf_bytes_ = send_bytes_ + recv_bytes_;
      msg_bytes_ = send_bytes_;
      oop_default_ = true;
      break;
    case OpKind::ALLTOALL:
      send_bytes_ = recv_bytes_ = n * P * dt;
      buf_bytes_ = 2 * send_bytes_;
      msg_bytes_ = send_bytes_;
      oop_default_ = true;
      break;
    case OpKind::ALLTOALLV: {
      size_t s = 0, r = 0;
      for (size_t i = 0; i < desc.send_counts.size(); ++i)
        s = std::max(s, desc.send_offsets[i] + desc.send_counts[i]);
      for (size_t i = 0; i < desc.recv_counts.size(); ++i)
        r = std::max(r, desc.recv_offsets[i] + desc.recv_counts[i]);
      send_bytes_ = s * dt;
      recv_bytes_ = r * dt;
      buf_bytes_ = send_bytes_ + recv_bytes_;
      msg_bytes_ = send_bytes_;
      oop_default_ = true;
      break;
    }
    case OpKind::GATHER:
      send_bytes_ = n * dt;
      recv_bytes_ = n * P * dt;
      buf_bytes_ = 0;
      msg_bytes_ = recv_bytes_;
      oop_default_ = true;
      break;
    case OpKind::SCATTER:
      send_bytes_ = n * P * dt;
      recv_bytes_ = n * dt;
      buf_bytes_ = 0;
      msg_bytes_ = send_bytes_;
      oop_default_ = true;
      break;
    case OpKind::SENDRECV_LIST: {
      size_t s = 0, r = 0;
      for (size_t i = 0; i < desc.send_counts.size(); ++i)
        s = std::max(s, desc.send_offsets[i] + desc.send_counts[i]);
      for (size_t i = 0; i < desc.recv_counts.size(); ++i)
        r = std::max(r, desc.recv_offsets[i] + desc.recv_counts[i]);
      send_bytes_ = s * dt;
      recv_bytes_ = r * dt;
      buf_bytes_ = send_bytes_ + recv_bytes_;
      msg_bytes_ = send_bytes_;
      oop_default_ = true;
      break;
    }
    case OpKind::FUSED_UPDATE:
      send_bytes_ = n * P * dt;          // full gradient buffer (n = owned elements)
      recv_bytes_ = n * P * dtype_size(desc.has_out_dtype ? desc.out_dtype : desc.dtype);
      buf_bytes_ = 0;
      msg_bytes_ = send_bytes_;
      break;
    case OpKind::GEMM_RS:
      send_bytes_ = 0;
      recv_bytes_ = (size_t)desc.gemm.M / P * (size_t)desc.gemm.N * dtype_size(desc.has_out_dtype ? desc.out_dtype : desc.dtype);
      buf_bytes_ = 0;
      msg_bytes_ = (size_t)desc.gemm.M * (size_t)desc.gemm.N * 2;
      break;
    case OpKind::AG_GEMM:
      send_bytes_ = (size_t)desc.gemm.M / P * (size_t)desc.gemm.K * 2;   // this rank's shard, read by every peer
      recv_bytes_ = 0;                                                     // Y and the gathered X are local only
      buf_bytes_ = 0;
      msg_bytes_ = (size_t)desc.gemm.M * (size_t)desc.gemm.K * 2;
      break;
  }
  // priority lane: large gradient messages of the earliest operations overtake the rest (the intent of the
  // reference's newest-first Rabenseifner progress, eplib/allreduce_pr.c:76-79: first-layer gradients first).
  // With progress threads the launch ORDER is prioritised instead (ProgressEngine::choose: newest big gradient first,
  // the same order on every rank); the static lane is the fallback when collectives are launched by the caller.
  lane = 0;
  const bool dynamic_order = ctx->progress && ctx->progress->servers() > 0;
  if (!dynamic_order && ctx->env.msg_priority && desc.comp_type == CommDesc::PARAM_GRAD && msg_bytes_ >= ctx->env.msg_priority_threshold) {
    int ops = ctx->session_ops_hint > 0 ? ctx->session_ops_hint : 4;
    int cut = std::max(1, ops / 4);
    if (desc.op_uid >= 0 && (desc.op_uid % (int64_t)std::max(ops, 1)) < cut && ctx->env.msg_priority_mode == 1) lane = 1;
  }
  setup_done = true;
  if (ctx->backend) ctx->backend->prepare(*this);
}

void CommRequest::start(void* s, void* r) {
  MLSLB_ASSERT(setup_done, "request started before Setup()");
  MLSLB_ASSERT(state.load(std::memory_order_acquire) == IDLE, "%s request started while still in flight",
               opkind_name(desc.kind));
  send = s;
  recv = r;
  // In-place all-gather convention (MPI_IN_PLACE): rank i's contribution already sits in its slot of the
  // receive buffer.
  if (s == r && desc.group && (desc.kind == OpKind::ALLGATHER || desc.kind == OpKind::ALLGATHERV)) {
    size_t off = 0;
    if (desc.kind == OpKind::ALLGATHER) off = (size_t)desc.group->idx * desc.count;
    else for (int p = 0; p < desc.group->idx; ++p) off += desc.recv_counts[(size_t)p];
    send = (char*)r + off * dtype_size(desc.dtype);
  }
  if (ctx->env.pointer_check) {
    if (send_bytes_) ctx->check_pointer(s, send_bytes_, "send buffer");
    if (recv_bytes_ && r) ctx->check_pointer(r, recv_bytes_, "recv buffer");
  }
  start_ns = now_ns();
  group_seq = 0;            // the ticket on the group's row is drawn when the command is LAUNCHED (ProgressEngine): message
                            // prioritisation may launch in another order than the program started, on every rank alike
  state.store(QUEUED, std::memory_order_release);
  ctx->backend->on_start(*this);
  ctx->progress->submit(this);
}

static void spin_until_launched(CommRequest* r) {
  uint64_t spins = 0, t0 = 0;
  while (r->state.load(std::memory_order_acquire) == CommRequest::QUEUED) {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
    if ((++spins & 0xfff) == 0) {
      sched_yield();
      if (r->ctx->boot && r->ctx->boot->poisoned())
        MLSLB_ASSERT(false, "job poisoned by rank %d", (int)r->ctx->boot->poisoned() - 1);
      if (!t0) t0 = now_ns();
      int wd = r->ctx->env.watchdog_sec;
      if (wd > 0 && now_ns() - t0 > (uint64_t)wd * 1000000000ull) {
        if (r->ctx->boot) r->ctx->boot->poison(r->ctx->rank);
        MLSLB_ASSERT(false, "watchdog: %s on group row %d seq %llu was never launched (a peer is missing?)",
                     opkind_name(r->desc.kind), r->desc.group ? r->desc.group->row : -1,
                     (unsigned long long)r->group_seq);
      }
    }
  }
}

static void rethrow_failed(CommRequest* r) {
  const std::string msg = r->error;
  r->state.store(CommRequest::IDLE, std::memory_order_release);
  MLSLB_ASSERT(false, "%s failed on the progress thread: %s", opkind_name(r->desc.kind), msg.c_str());
}

void* CommRequest::wait() {
  int st = state.load(std::memory_order_acquire);
  if (st == IDLE) return recv;      // nothing in flight (already completed through Test)
  spin_until_launched(this);
  if (state.load(std::memory_order_acquire) == FAILED) rethrow_failed(this);
  ctx->backend->wait(*this);
  done_ns = now_ns();
  if (!ctx->trace_prefix.empty()) ctx->trace_request(*this);
  state.store(IDLE, std::memory_order_release);
  return recv;
}

void* CommRequest::test(bool* done) {
  int st = state.load(std::memory_order_acquire);
  if (st == IDLE) {
    *done = true;
    return recv;
  }
  if (st == QUEUED) {
    *done = false;
    return nullptr;
  }
  if (st == FAILED) rethrow_failed(this);
  if (ctx->backend->test(*this)) {
    done_ns = now_ns();
    if (!ctx->trace_prefix.empty()) ctx->trace_request(*this);
    state.store(IDLE, std::memory_order_release);
    *done = true;
    return recv;
  }
  *done = false;
  return nullptr;
}

// ============================================================================================================
// ProgressEngine
// ============================================================================================================
static void pin_thread(const std::string& affinity, int idx) {
  // "MLSL_SERVER_AFFINITY=5,6,7": server i runs on the i-th listed cpu.  Default (reference eplib/env.c:207-218):
  // the last cores in reverse order.
  int ncpu = (int)sysconf(_SC_NPROCESSORS_ONLN);
  int cpu = -1;
  if (!affinity.empty()) {
    std::vector<int> list;
    const char* p = affinity.c_str();
    while (*p) {
      list.push_back(atoi(p));
      const char* c = strchr(p, ',');
      if (!c) break;
      p = c + 1;
    }
    if (!list.empty()) cpu = list[idx % list.size()];
  }
  if (cpu < 0 || cpu >= ncpu) return;   // unpinned by default: ranks share the node, let the OS place servers
  cpu_set_t set;
  CPU_ZERO(&set);
  CPU_SET(cpu, &set);
  pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
}

static void draw_ticket(CommRequest* r) {
  ProcessGroup* g = r->desc.group;
  r->group_seq = (!g || g->size() <= 1 || g->is_self) ? 0 : ++g->seq[r->lane];
}

ProgressEngine::ProgressEngine(RankContext* ctx, int num_servers) : ctx_(ctx) {
  for (int i = 0; i < num_servers; ++i) {
    servers_.emplace_back(new Server());
    Server* s = servers_.back().get();
    s->th = std::thread([this, s, i] { run(s, i); });
  }
}

ProgressEngine::~ProgressEngine() {
  for (auto& s : servers_) {
    Command c;
    c.kind = Command::STOP;
    {
      std::lock_guard<std::mutex> g(s->mu);
      while (!s->ring.push(c)) sched_yield();
    }
  }
  for (auto& s : servers_)
    if (s->th.joinable()) s->th.join();
}

void ProgressEngine::submit(CommRequest* r) {
  if (servers_.empty()) {
    draw_ticket(r);
    ctx_->backend->launch(*r);
    launched_.fetch_add(1, std::memory_order_relaxed);
    return;
  }
  int row = r->desc.group ? std::max(r->desc.group->row, 0) : 0;
  // rows spread over the servers first (different groups progress independently), the priority lane of a row goes to
  // the neighbouring server
  Server* s = servers_[(size_t)(row + r->lane) % servers_.size()].get();
  Command c;
  c.kind = Command::EXEC;
  c.req = r;
  std::lock_guard<std::mutex> g(s->mu);
  s->submitted.fetch_add(1, std::memory_order_relaxed);
  while (!s->ring.push(c)) sched_yield();   // back-pressure (the reference only detects overflow server-side)
}

void ProgressEngine::drain() {
  for (auto& s : servers_)
    while (s->completed.load(std::memory_order_acquire) < s->submitted.load(std::memory_order_acquire)) sched_yield();
}

void ProgressEngine::suspend() {
  const uint64_t gen = suspend_gen_.fetch_add(1) + 1;
  for (auto& s : servers_) {
    Command c;
    c.kind = Command::SUSPEND;
    c.arg = gen;
    std::lock_guard<std::mutex> g(s->mu);
    while (!s->ring.push(c)) sched_yield();
  }
}

void ProgressEngine::resume() {
  // generation based: a resume that overtakes a not-yet-processed SUSPEND still cancels it
  resume_gen_.store(suspend_gen_.load(std::memory_order_acquire), std::memory_order_release);
}

std::vector<int64_t> ProgressEngine::recent_launches() {
  std::lock_guard<std::mutex> g(recent_mu_);
  return recent_;
}

void ProgressEngine::exec(Server* s, CommRequest* r) {
  draw_ticket(r);
  {
    std::lock_guard<std::mutex> g(recent_mu_);
    if (recent_.size() >= 256) recent_.erase(recent_.begin());
    recent_.push_back(r->desc.op_uid);
  }
  try {
    ctx_->backend->launch(*r);
  } catch (const std::exception& e) {
    // the error belongs to the caller of Wait / Test: keep it on the request and fail it (a LAUNCHED state would
    // leave host / net waiters spinning for a completion that never comes)
    fprintf(stderr, "(r%d) progress thread: %s\n", ctx_->rank, e.what());
    if (ctx_->boot) ctx_->boot->poison(ctx_->rank);
    r->error = e.what();
    r->state.store(CommRequest::FAILED, std::memory_order_release);
  }
  launched_.fetch_add(1, std::memory_order_relaxed);
  s->completed.fetch_add(1, std::memory_order_release);
}

// Message prioritisation (reference eplib/allreduce_pr.c:76-79 progresses its big all-reduces newest first, so the
// gradients of the FIRST layers - produced last by the backward pass, needed first by the next forward pass - overtake
// the bulk; selected at eplib/cqueue.c:1999-2012 by size).  On a GPU the order is fixed once kernels sit in a stream, so
// the choice is made here, before the launch: big gradient messages are held back while the row already has
// kPrioWindow of them in flight, and when a slot frees up the NEWEST queued one goes next; everything else (activations,
// small messages) is launched at once, oldest first.  All members must launch a row's collectives in the same order:
// the group's first member decides and appends the choice to the row's order log in the shared control block, the
// others replay the log.  SPMD programs start the same collectives before they wait, so a follower always finds the
// logged command in its own queue eventually.
constexpr size_t kPrioWindow = 2;

static uint64_t order_key(const CommRequest* r) {
  uint64_t k = (uint64_t)(r->desc.op_uid + 1) * 0x9e3779b97f4a7c15ull;
  k ^= ((uint64_t)r->desc.kind << 56) ^ ((uint64_t)r->desc.comp_type << 48) ^ (uint64_t)r->desc.count;
  return k | 1ull;   // never 0
}

bool ProgressEngine::prioritised(const CommRequest* r) const {
  return ctx_->env.msg_priority && r->desc.comp_type == CommDesc::PARAM_GRAD && r->msg_bytes() >= ctx_->env.msg_priority_threshold &&
         r->desc.group && r->desc.group->size() > 1 && r->desc.group->row >= 0;
}

CommRequest* ProgressEngine::choose(Server* s) {
  if (s->pending.empty()) return nullptr;
  BootCtl* c = ctx_->boot ? ctx_->boot->ctl() : nullptr;
  if (!ctx_->env.msg_priority || !c) {          // plain FIFO
    CommRequest* r = s->pending.front();
    s->pending.erase(s->pending.begin());
    return r;
  }
  // retire finished prioritised launches from the in-flight windows
  for (size_t i = 0; i < s->inflight.size();) {
    if (ctx_->backend->peek_done(*s->inflight[i])) s->inflight.erase(s->inflight.begin() + i);
    else ++i;
  }
  // one ordering domain per (row, lane); walk the domains that have something queued, oldest command first
  for (size_t qi = 0; qi < s->pending.size(); ++qi) {
    CommRequest* first = s->pending[qi];
    ProcessGroup* g = first->desc.group;
    if (!g || g->size() <= 1 || g->row < 0) {   // local command: no peer has to agree
      s->pending.erase(s->pending.begin() + qi);
      return first;
    }
    const int dom = g->row * 2 + first->lane;
    bool seen = false;
    for (size_t k = 0; k < qi; ++k) seen |= s->pending[k]->desc.group == g && s->pending[k]->lane == first->lane;
    if (seen) continue;                          // this domain was already looked at through an older command
    auto in_domain = [&](CommRequest* r) { return r->desc.group == g && r->lane == first->lane; };
    size_t pick = SIZE_MAX;
    if (g->idx == 0) {
      // leader: the oldest normal command if there is one, else the newest prioritised one when the window has room
      for (size_t k = qi; k < s->pending.size() && pick == SIZE_MAX; ++k)
        if (in_domain(s->pending[k]) && !prioritised(s->pending[k])) pick = k;
      if (pick == SIZE_MAX) {
        size_t fl = 0;
        for (CommRequest* r : s->inflight) fl += in_domain(r);
        if (fl < kPrioWindow) {
          for (size_t k = s->pending.size(); k-- > qi;)
            if (in_domain(s->pending[k])) {
              pick = ctx_->env.msg_priority_mode == 1 ? k : SIZE_MAX;
              if (pick != SIZE_MAX) break;
            }
          if (pick == SIZE_MAX) pick = qi;       // mode 0: oldest first
        }
      }
      if (pick == SIZE_MAX) continue;
      const uint64_t n = c->order_head[dom].load(std::memory_order_relaxed);
      // never lap a follower that still has to read the entry this one would overwrite
      uint64_t slowest = n;
      for (int m : g->members)
        if (m != ctx_->rank) slowest = std::min(slowest, c->order_pos[dom][m].load(std::memory_order_acquire));
      if (n - slowest >= kOrderLogDepth) continue;
      c->order_log[dom][n % kOrderLogDepth].store(order_key(s->pending[pick]), std::memory_order_relaxed);
      c->order_head[dom].store(n + 1, std::memory_order_release);
    } else {
      const uint64_t pos = c->order_pos[dom][ctx_->rank].load(std::memory_order_relaxed);
      if (c->order_head[dom].load(std::memory_order_acquire) <= pos) continue;      // the leader has not decided yet
      const uint64_t key = c->order_log[dom][pos % kOrderLogDepth].load(std::memory_order_relaxed);
      for (size_t k = qi; k < s->pending.size() && pick == SIZE_MAX; ++k)
        if (in_domain(s->pending[k]) && order_key(s->pending[k]) == key) pick = k;
      if (pick == SIZE_MAX) continue;            // not started by this rank's program yet
      c->order_pos[dom][ctx_->rank].store(pos + 1, std::memory_order_release);
    }
    CommRequest* r = s->pending[pick];
    s->pending.erase(s->pending.begin() + pick);
    if (prioritised(r)) s->inflight.push_back(r);
    return r;
  }
  return nullptr;
}

void ProgressEngine::run(Server* s, int idx) {
  pin_thread(ctx_->env.server_affinity, idx);
  set_log_rank(ctx_->rank);
  uint64_t idle = 0;
  for (;;) {
    Command c;
    bool got = false;
    while (s->ring.pop(c)) {
      got = true;
      if (c.kind == Command::STOP) {
        for (CommRequest* r : s->pending) exec(s, r);   // nothing may stay queued behind a shutdown
        s->pending.clear();
        return;
      }
      if (c.kind == Command::SUSPEND) {
        s->parked.store(true, std::memory_order_release);
        while (resume_gen_.load(std::memory_order_acquire) < c.arg) usleep(100);
        s->parked.store(false, std::memory_order_release);
        continue;
      }
      if (c.kind == Command::EXEC) s->pending.push_back(c.req);
    }
    bool launched = false;
    while (CommRequest* r = choose(s)) {
      exec(s, r);
      launched = true;
    }
    if (got || launched) {
      idle = 0;
      continue;
    }
    // commands held back by the ordering rules keep the thread polling; an empty queue lets it back off
    if (++idle < 20000 || !s->pending.empty()) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
      if (!s->pending.empty() && (idle & 0xfff) == 0) {
        sched_yield();
        if (ctx_->boot && ctx_->boot->poisoned()) {      // a dead peer can never log / start the command: fail what waits
          for (CommRequest* r : s->pending) {
            r->error = "job poisoned while the command waited for its turn";
            r->state.store(CommRequest::FAILED, std::memory_order_release);
            s->completed.fetch_add(1, std::memory_order_release);
          }
          s->pending.clear();
        }
      }
    } else if (idle < 200000) {
      sched_yield();
    } else {
      usleep(50);
    }
    if (ctx_->boot && (idle & 0x3ff) == 0) ctx_->boot->heartbeat();
  }
}

// ============================================================================================================
// Backend defaults
// ============================================================================================================
void Backend::pack_blocks(const BlockDesc* blocks, size_t nblocks, size_t local_fm_count, DType dt, const void* src,
                          void* dst, bool unpack) {
  const size_t es = dtype_size(dt);
  for (size_t b = 0; b < nblocks; ++b) {
    const BlockDesc& k = blocks[b];
    const size_t row = k.fm_cnt * k.fm_size * es;   // one minibatch row of the block is contiguous on both sides
    for (size_t mb = 0; mb < k.mb_cnt; ++mb) {
      size_t local_off = ((mb + k.mb_off) * local_fm_count + k.fm_off) * k.fm_size * es;
      size_t comm_off = (k.buf_off + mb * k.fm_cnt * k.fm_size) * es;
      if (!unpack) memcpy((char*)dst + comm_off, (const char*)src + local_off, row);
      else memcpy((char*)dst + local_off, (const char*)src + comm_off, row);
    }
  }
}

void Backend::rma_copy(void* dst, const void* src, size_t bytes) { memcpy(dst, src, bytes); }
void Backend::copy_from_host(void* dst, const void* src, size_t bytes) { memcpy(dst, src, bytes); }

// ============================================================================================================
// PointerChecker
// ============================================================================================================
void PointerChecker::add(const void* p, size_t len) {
  std::lock_guard<std::mutex> g(mu_);
  ranges_.insert({(uintptr_t)p, (uintptr_t)p + len});
}
void PointerChecker::remove(const void* p) {
  std::lock_guard<std::mutex> g(mu_);
  for (auto it = ranges_.begin(); it != ranges_.end(); ++it)
    if (it->first == (uintptr_t)p) {
      ranges_.erase(it);
      return;
    }
}
bool PointerChecker::check(const void* p, size_t len) const {
  std::lock_guard<std::mutex> g(mu_);
  uintptr_t a = (uintptr_t)p, b = a + len;
  auto it = ranges_.upper_bound({a, UINTPTR_MAX});
  if (it == ranges_.begin()) return false;
  --it;
  return a >= it->first && b <= it->second;
}
size_t PointerChecker::count() const {
  std::lock_guard<std::mutex> g(mu_);
  return ranges_.size();
}

// ============================================================================================================
// RankContext: groups, request storage
// ============================================================================================================
// Collective over `parent`: ranks passing the same colour end up in the same new group, ordered by parent index
// (MPI_Comm_split(parent, colour, key=rank) semantics, which is what the reference's CreateProcessGroup(colour)
// does; reference src/comm_ep.cpp:1791-1830).  The exchange also agrees on (a) a signal row that is free on
// every member and (b) a ticket base above anything a member has ever used, so flags on a recycled row can
// never alias an older group's.
ProcessGroup* RankContext::create_group_by_color(ProcessGroup* parent, int color) {
  MLSLB_ASSERT(parent != nullptr, "null parent group");
  struct Msg {
    int32_t color;
    int32_t pad;
    uint64_t row_used;
    uint64_t hwm;
  } mine{color, 0, row_used, std::max(seq_hwm, parent->hwm())};
  std::vector<Msg> all(parent->size());
  if (parent->size() > 1) {
    boot->group_allgather(parent->members, parent->row, ++parent->ctl_seq, &mine, all.data(), sizeof(Msg));
  } else {
    all[0] = mine;
  }
  std::vector<int> members;
  uint64_t used = 0, base = 0;
  for (int i = 0; i < parent->size(); ++i)
    if (all[i].color == color) {
      members.push_back(parent->members[i]);
      used |= all[i].row_used;
      base = std::max(base, all[i].hwm);
    }
  return create_group_from_members(members, used, base);
}

ProcessGroup* RankContext::create_group_from_members(const std::vector<int>& members, uint64_t used, uint64_t base) {
  ProcessGroup* g = new ProcessGroup();
  g->ctx = this;
  g->members = members;
  g->idx = -1;
  for (size_t i = 0; i < members.size(); ++i) {
    for (size_t j = 0; j < i; ++j) MLSLB_ASSERT(members[j] != members[i], "rank %d is listed twice in a group", members[i]);
    if (members[i] == rank) g->idx = (int)i;
  }
  if (g->idx < 0) {
    delete g;
    MLSLB_ASSERT(false, "rank %d is not a member of the group it is creating", rank);
  }
  g->is_self = g->members.size() == 1;
  if (g->is_self) {
    g->row = -1;
  } else {
    used |= row_used;
    int row = -1;
    for (int r = 1; r < kMaxGroupRows - 1; ++r)   // the last row is reserved for single-rank (self) groups
      if (!(used & (1ull << r))) {
        row = r;
        break;
      }
    if (row < 0) delete g;
    MLSLB_ASSERT(row > 0, "out of process-group rows (max %d live groups)", kMaxGroupRows);
    g->row = row;
    row_used |= 1ull << row;
    base = std::max(base, seq_hwm);
    g->seq[0] = g->seq[1] = g->ctl_seq = base;
    seq_hwm = std::max(seq_hwm, base);
  }
  if (boot && boot->ctl() && g->row >= 0)          // the row's launch-order log continues where the previous group on it stopped
    for (int l = 0; l < 2; ++l)
      boot->ctl()->order_pos[g->row * 2 + l][rank].store(boot->ctl()->order_head[g->row * 2 + l].load(std::memory_order_acquire),
                                                         std::memory_order_release);
  if (backend) backend->group_created(*g);
  return g;
}

void RankContext::group_barrier(ProcessGroup* g) {
  if (!g || g->size() <= 1) return;
  boot->group_allgather(g->members, g->row, ++g->ctl_seq, nullptr, nullptr, 0);
}

void RankContext::free_group(ProcessGroup* g) {
  if (!g || g == world_group || g == self_group) return;
  if (g->size() > 1) {
    // like MPI_Comm_free this is collective: nobody recycles the row while a peer still polls it
    if (progress) progress->drain();
    group_barrier(g);
  }
  if (backend) backend->group_destroyed(*g);
  seq_hwm = std::max(seq_hwm, g->hwm());
  if (g->row > 0) row_used &= ~(1ull << g->row);
  delete g;
}

void RankContext::trace_request(const CommRequest& r) {
  TraceEvent e{r.start_ns, r.done_ns, (int)r.desc.kind, r.desc.group ? r.desc.group->row : -1, r.lane, r.msg_bytes(), r.device_ns_last};
  std::lock_guard<std::mutex> g(trace_mu);
  if (trace.size() < (size_t)1 << 20) trace.push_back(e);   // bounded: ~32 MB per rank
}

void RankContext::trace_dump() {
  if (trace_prefix.empty()) return;
  std::vector<TraceEvent> ev;
  {
    std::lock_guard<std::mutex> g(trace_mu);
    ev.swap(trace);
  }
  const std::string path = trace_prefix + "." + std::to_string(rank) + ".json";
  FILE* f = fopen(path.c_str(), "w");
  if (!f) {
    MLSLB_LOG(LOG_ERROR, "cannot write trace file %s", path.c_str());
    return;
  }
  // Chrome trace event format ("X" = complete event, microseconds); one process per rank, one thread per (row, lane)
  fprintf(f, "{\"traceEvents\":[\n");
  fprintf(f, "{\"ph\":\"M\",\"pid\":%d,\"name\":\"process_name\",\"args\":{\"name\":\"mlsl rank %d (%s)\"}}", rank, rank,
          backend ? backend->name() : "finalized");
  for (const TraceEvent& e : ev)
    fprintf(f, ",\n{\"ph\":\"X\",\"pid\":%d,\"tid\":%d,\"ts\":%.3f,\"dur\":%.3f,\"name\":\"%s\",\"args\":{\"bytes\":%zu,\"row\":%d,\"lane\":%d,\"device_us\":%.3f}}",
            rank, (e.row < 0 ? 0 : e.row) * 2 + e.lane, e.t0 / 1000.0, (e.t1 > e.t0 ? e.t1 - e.t0 : 0) / 1000.0,
            opkind_name((OpKind)e.kind), e.bytes, e.row, e.lane, e.device_ns / 1000.0);
  fprintf(f, "\n]}\n");
  fclose(f);
  MLSLB_LOG(LOG_INFO, "wrote %zu trace events to %s", ev.size(), path.c_str());
}

void RankContext::register_request(CommRequest* r) {
  std::lock_guard<std::mutex> g(req_mu);
  inflight.insert(r);
}

void RankContext::remove_request(CommRequest* r) {
  {
    std::lock_guard<std::mutex> g(req_mu);
    inflight.erase(r);
  }
  if (r->one_shot) delete r;
}

void RankContext::check_pointer(const void* p, size_t len, const char* what) {
  MLSLB_ASSERT(ptrcheck.check(p, len), "pointer check: %s [%p, +%zu) is not inside memory obtained from Environment::Alloc",
               what, p, len);
}

// ============================================================================================================
// Context lifecycle
// ============================================================================================================
struct InprocWorldReg {
  std::vector<std::unique_ptr<Bootstrap>> boots;
};
static std::mutex g_reg_mu;
static std::map<int, std::unique_ptr<InprocWorldReg>> g_worlds;
static int g_next_world = 1;
static RankContext g_process_ctx;
static thread_local RankContext* tls_ctx = nullptr;
static thread_local std::unique_ptr<Bootstrap> tls_boot;

RankContext* process_context() { return &g_process_ctx; }
RankContext* current_context() { return tls_ctx ? tls_ctx : &g_process_ctx; }
bool thread_is_inproc_rank() { return tls_ctx != nullptr; }

int inproc_world_create(int nranks) {
  std::lock_guard<std::mutex> g(g_reg_mu);
  int id = g_next_world++;
  auto reg = std::unique_ptr<InprocWorldReg>(new InprocWorldReg());
  reg->boots = Bootstrap::create_inproc(nranks);
  g_worlds[id] = std::move(reg);
  return id;
}

void inproc_world_destroy(int id) {
  std::lock_guard<std::mutex> g(g_reg_mu);
  g_worlds.erase(id);
}

void inproc_bind_thread(int id, int rank) {
  std::lock_guard<std::mutex> g(g_reg_mu);
  auto it = g_worlds.find(id);
  MLSLB_ASSERT(it != g_worlds.end(), "unknown in-process world %d", id);
  MLSLB_ASSERT(rank >= 0 && rank < (int)it->second->boots.size() && it->second->boots[rank],
               "in-process rank %d unavailable", rank);
  MLSLB_ASSERT(tls_ctx == nullptr, "thread already bound to a virtual rank");
  tls_boot = std::move(it->second->boots[rank]);
  tls_ctx = new RankContext();
}

void inproc_unbind_thread() {
  if (!tls_ctx) return;
  if (tls_ctx->initialized) context_finalize(tls_ctx);
  delete tls_ctx;
  tls_ctx = nullptr;
  tls_boot.reset();
}

std::unique_ptr<Bootstrap> take_thread_bootstrap() { return std::move(tls_boot); }

static std::string derive_job_key(const EnvConfig& e) {
  if (!e.job_id.empty()) return e.job_id;
  std::string k;
  if (const char* v = getenv("TORCHELASTIC_RUN_ID")) k += v;
  if (const char* v = getenv("MASTER_PORT")) k += std::string("p") + v;
  if (k.empty()) k = "default";
  for (auto& ch : k)
    if (!isalnum((unsigned char)ch)) ch = '_';
  return k;
}

static void poison_on_fail() {
  RankContext* c = current_context();
  if (c && c->boot) c->boot->poison(c->rank);
}

void context_init(RankContext* ctx) {
  ctx->env = parse_env();
  set_log_level(ctx->env.log_level);
  ctx->init_pid = (int)getpid();
  if (const char* t = getenv("MLSL_TRACE_FILE")) ctx->trace_prefix = t;
  if (auto b = take_thread_bootstrap()) {
    ctx->boot = std::move(b);
  } else {
    int world = ctx->env.world > 0 ? ctx->env.world : 1;
    int rank = ctx->env.rank >= 0 ? ctx->env.rank : 0;
    const bool net = ctx->env.backend == "net" || ctx->env.backend == "tcp";
    // torchrun tells how many ranks share this node: a job that spans nodes cannot meet in /dev/shm
    if (const char* lws = getenv("LOCAL_WORLD_SIZE")) {
      const int local_world = atoi(lws);
      MLSLB_ASSERT(net || local_world <= 0 || local_world >= world,
                   "this job spans nodes (%d of %d ranks are local): the %s backend covers one node - use MLSL_BACKEND=net "
                   "(TCP, host memory) across nodes", local_world, world, ctx->env.backend == "cuda" ? "cuda" : "host");
    }
    if (world == 1) {
      auto v = Bootstrap::create_inproc(1);
      ctx->boot = std::move(v[0]);
    } else if (net) {
      ctx->boot = Bootstrap::create_tcp(ctx->env.master_addr, ctx->env.master_port, rank, world);
    } else {
      ctx->boot = Bootstrap::create_shm(derive_job_key(ctx->env), rank, world);
    }
  }
  ctx->rank = ctx->boot->rank();
  ctx->world = ctx->boot->size();
  set_log_rank(ctx->rank);
  set_fail_hook(poison_on_fail);

  // base groups
  ctx->world_group = new ProcessGroup();
  ctx->world_group->ctx = ctx;
  for (int r = 0; r < ctx->world; ++r) ctx->world_group->members.push_back(r);
  ctx->world_group->idx = ctx->rank;
  ctx->world_group->row = 0;
  ctx->world_group->is_world = true;
  ctx->world_group->is_self = ctx->world == 1;
  ctx->row_used = 1;
  ctx->self_group = new ProcessGroup();
  ctx->self_group->ctx = ctx;
  ctx->self_group->members.push_back(ctx->rank);
  ctx->self_group->idx = 0;
  ctx->self_group->row = -1;
  ctx->self_group->is_self = true;
  ctx->global_group = ctx->world_group;

  // Backend selection.  MLSL_BACKEND=cuda | host; unset ("auto") means host for native programs: sources written
  // against the reference fill Environment::Alloc memory and the communication buffers with CPU code, which only the
  // host backend's memory allows.  The Python layer picks cuda itself when a GPU is visible (mlsl_b200/api.py).
  std::string want = ctx->env.backend;
  if (want == "auto") want = "host";
  if (want == "cuda") {
    ctx->backend = make_cuda_backend(ctx);
    MLSLB_ASSERT(ctx->backend != nullptr, "MLSL_BACKEND=cuda requested but no usable CUDA device / extension");
  } else if ((want == "net" || want == "tcp") && ctx->boot->is_tcp()) {
    ctx->backend = make_net_backend(ctx);
  } else {
    MLSLB_ASSERT(want == "host" || want == "net" || want == "tcp", "unknown MLSL_BACKEND '%s' (host|cuda|net)", want.c_str());
    ctx->backend = make_host_backend(ctx);   // also a single-rank or in-process "net" job: nothing crosses a wire
  }
  ctx->backend->group_created(*ctx->world_group);

  // Servers: the reference switches its endpoint servers off on a single node unless MLSL_NUM_SERVERS is set
  // (src/comm_ep.cpp:1585-1602).  Same rule here, for both backends.
  // MLSL_CHECK_SINGLE_NODE=0 switches that rule off: the reference's default of 4 servers applies (host path).
  int ns = ctx->env.num_servers;
  if (ns < 0) {
    if (ctx->world <= 1) {
      ns = 0;
    } else if (ctx->backend->is_device()) {
      // one process per GPU: ONE progress thread per rank launches the collectives on the communication streams - the
      // API thread only records an event and pushes a ring entry (the reference's endpoint server, as a thread).  Ranks
      // that share a GPU (loop-back tests) and inline-stream mode launch from the caller.
      ns = ctx->backend->default_servers();
    } else if (ctx->boot->is_tcp()) {
      // jobs that span nodes: Start() must not run a whole TCP collective on the caller (ranks that start collectives on
      // overlapping groups in different orders would wait for each other): one progress thread, like the reference's servers
      ns = 1;
    } else if (!ctx->env.check_single_node) {
      ns = 4;
    } else {
      // the reference's own single-node rule (src/comm_ep.cpp:1585-1602): no servers unless asked for.  A collective
      // then runs inside Start() on the calling thread - 2 us for a small all-reduce on 4 ranks; a hop through a
      // progress thread costs 20-400 us on a busy or virtualised node (measured, profiles/host_backend_vs_reference_cpu.txt).
      // MLSL_NUM_SERVERS=1 buys true asynchronous progress when there are cores to spare.
      ns = 0;
    }
  }
  if (ctx->boot->is_tcp() && ns > 1) ns = 1;   // one queue keeps the issue order of collectives that share connections
  ctx->progress.reset(new ProgressEngine(ctx, ns));
  ctx->initialized = true;
  install_signal_handlers(ctx);
  ctx->boot->barrier();
}

void context_finalize(RankContext* ctx) {
  if (!ctx->initialized) return;
  ctx->initialized = false;   // never torn down twice, also when a step below fails
  // Every step runs even when an earlier one failed (a poisoned job throws at the first barrier): a failing rank must
  // still stop its threads and give its memory back; the first error is re-thrown at the end.
  std::string first_error;
  auto step = [&](auto&& fn) {
    try {
      fn();
    } catch (const std::exception& e) {
      if (first_error.empty()) first_error = e.what();
    }
  };
  if (!ctx->boot->inproc()) remove_signal_handlers();
  step([&] { ctx->trace_dump(); });
  step([&] { io_shutdown(ctx); });
  step([&] { if (ctx->progress) ctx->progress->drain(); });
  step([&] { ctx->boot->barrier(); });
  step([&] { ctx->progress.reset(); });
  if (ctx->global_group != ctx->world_group) {
    ProcessGroup* g = ctx->global_group;
    ctx->global_group = ctx->world_group;
    step([&] { ctx->free_group(g); });
  }
  if (ctx->backend) {
    step([&] { ctx->backend->group_destroyed(*ctx->world_group); });
    step([&] { ctx->backend->finalize(); });
    step([&] { ctx->backend.reset(); });
  }
  delete ctx->world_group;
  delete ctx->self_group;
  ctx->world_group = ctx->self_group = ctx->global_group = nullptr;
  step([&] { ctx->boot->barrier(); });
  ctx->boot.reset();
  ctx->row_used = 0;
  ctx->quant = QuantConfig();
  MLSLB_ASSERT(first_error.empty(), "finalize after a failure: %s", first_error.c_str());
}

}  // namespace mlslb
