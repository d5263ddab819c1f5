// Net backend: collectives between ranks that share nothing but a network (MLSL_BACKEND=net).
//
// The reference reaches other nodes through MPI (src/comm_ep.cpp issues MPI_I* calls on endpoint communicators); this
// backend is the self-contained counterpart for CPU clusters and for the control-plane-only parts of a multi-node job:
// a full mesh of TCP connections between the ranks, one generic primitive - every member sends at most one byte range to
// and receives at most one from every other member, all of them progressed together with non-blocking sockets - and
// every collective expressed as one or two such exchanges plus a local, fixed-order reduction (so results are bitwise
// identical on all ranks):
//   all-gather(v), all-to-all(v), gather, scatter, bcast, send/recv list, barrier : one exchange
//   reduce-scatter : exchange of slices, then reduce          reduce : gather to the root, then reduce
//   all-reduce     : reduce-scatter + all-gather              fused update : reduce-scatter + optimizer + all-gather
// Messages carry a tag (signal row, lane, sequence number, step), so collectives of different groups may be in flight on
// the same connection; an early message is parked until its collective asks for it.  Buffers are ordinary host memory.
//
// On top of that (round 2): an exchange may carry several tagged messages per peer and run a callback per arrival, which may
// queue further sends - large reductions travel as a pipeline of pieces; ranks of one node talk through shared-memory byte
// rings instead of their socket (same stream semantics); groups with several members per node run all-reduce, all-gather,
// reduce-scatter, broadcast and the fused update in two levels (node-local step, 1/L of the bytes between nodes, node-local
// step), the compressed all-reduce with fp8 only on the wire between nodes.  Each of those two-level collectives is ONE
// exchange from two pieces per block on: what arrives on one level is reduced / passed on to the other level inside the
// arrival callback, so shared-memory work of early pieces overlaps the wire time of later ones (MLSL_NET_HIER_PIPELINE=0:
// the levels as separate exchanges).
#include <fcntl.h>
#include <poll.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/uio.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>

#include "heap.hpp"
#include "log.hpp"
#include "numeric.hpp"
#include <functional>

#include "quant.hpp"
#include "runtime.hpp"
#include "tcp_control.hpp"

namespace mlslb {

namespace {

struct Seg {
  int peer;      // global rank
  char* ptr;
  size_t bytes;
  uint64_t tag = 0;   // 0: the tag of the exchange; chunked collectives give every chunk its own
};

constexpr size_t kOneShotBytes = 32 << 10;   // all-reduce up to this size: one exchange of whole vectors
// A message nobody waits for yet is read into parking space only up to this size (eager); a larger one stays in the
// socket until its receive is posted and is then read straight into the user's buffer (rendezvous by TCP flow control).
// Safe because a pair of ranks starts its common collectives in the same order: nothing this rank needs first can be
// queued behind the held message on that connection.
constexpr size_t kEagerBytes = 32 << 10;
constexpr size_t kBcastSplitBytes = 128 << 10;   // broadcast from this size on: scatter + all-gather
constexpr int kTreeMinRanks = 4;                 // small broadcasts and barriers of this many members take log P steps

struct WireHdr {
  uint64_t tag, bytes;
};
struct MeshHello {           // first frame of a data connection
  uint32_t rank, pad;
  uint64_t token;
};

// One direction of a same-node pair: a byte stream through shared memory with the semantics of the socket it replaces
// (bounded, in order, the writer stalls when it is full).  Single producer, single consumer.
struct ShmRing {
  alignas(64) std::atomic<uint64_t> head;   // bytes written so far (producer)
  alignas(64) std::atomic<uint64_t> tail;   // bytes read so far (consumer)
  alignas(64) uint64_t cap;                 // power of two
  alignas(64) char data[1];

  size_t write(const char* p, size_t n) {
    const uint64_t h = head.load(std::memory_order_relaxed), t = tail.load(std::memory_order_acquire);
    n = std::min<size_t>(n, cap - (h - t));
    if (!n) return 0;
    const size_t o = h & (cap - 1), first = std::min<size_t>(n, cap - o);
    memcpy(data + o, p, first);
    if (n > first) memcpy(data, p + first, n - first);
    head.store(h + n, std::memory_order_release);
    return n;
  }
  size_t read(char* p, size_t n) {
    const uint64_t t = tail.load(std::memory_order_relaxed), h = head.load(std::memory_order_acquire);
    n = std::min<size_t>(n, h - t);
    if (!n) return 0;
    const size_t o = t & (cap - 1), first = std::min<size_t>(n, cap - o);
    memcpy(p, data + o, first);
    if (n > first) memcpy(p + first, data, n - first);
    tail.store(t + n, std::memory_order_release);
    return n;
  }
  bool readable() const { return head.load(std::memory_order_acquire) != tail.load(std::memory_order_relaxed); }
};

class Mesh {
 public:
  void init(RankContext* ctx) {
    ctx_ = ctx;
    rank_ = ctx->rank;
    world_ = ctx->world;
    fds_.assign(world_, -1);
    peers_.resize(world_);
    if (world_ == 1) return;
    int port = 0;
    struct Addr {
      char ip[48];
      int port;
      char node[80];     // which machine (and which launcher on it) the rank runs on: equal keys -> shared memory
    } mine, zero;
    memset(&zero, 0, sizeof(zero));
    mine = zero;
    const std::string& key = ctx->boot->key();                  // "master:port"
    const size_t colon = key.rfind(':');
    std::string my_ip = tcp_local_address_towards(key.substr(0, colon), atoi(key.c_str() + colon + 1));
    // the reference's ways to pick the interface (eplib/server.c:228-330), most explicit first
    const EnvConfig& ec = ctx->env;
    if (!ec.iface_name.empty() || ec.iface_idx >= 0) {
      const std::string a = tcp_address_of_interface(ec.iface_name, ec.iface_idx);
      MLSLB_ASSERT(!a.empty(), "no IPv4 interface matches MLSL_IFACE_NAME=%s / MLSL_IFACE_IDX=%d", ec.iface_name.c_str(), ec.iface_idx);
      my_ip = a;
    }
    if (!ec.hostname.empty() && ec.hostname_type != 0) {
      const std::string a = tcp_resolve_to_ip(ec.hostname);
      MLSLB_ASSERT(!a.empty(), "MLSL_HOSTNAME=%s does not resolve", ec.hostname.c_str());
      my_ip = a;
    }
    if (!ec.net_addr.empty()) my_ip = ec.net_addr;                // explicit address of this rank's interface
    // an interface that was asked for by name / address is also the only one the data port listens on
    const bool chosen = !ec.net_addr.empty() || !ec.iface_name.empty() || ec.iface_idx >= 0 || (!ec.hostname.empty() && ec.hostname_type != 0);
    int lfd = tcp_listen(chosen ? my_ip.c_str() : "*", 0, world_ + 8, &port);
    my_ip_ = my_ip;
    snprintf(mine.ip, sizeof(mine.ip), "%s", my_ip.c_str());
    mine.port = port;
    snprintf(mine.node, sizeof(mine.node), "%s", node_key(ec.node_rank).c_str());
    eager_bytes_ = (size_t)std::max(0l, ec.net_eager_kb) << 10;
    rate_Bps_ = ec.net_emulate_gbit * 1e9 / 8;
    std::vector<Addr> all(world_);
    ctx->boot->allgather(&mine, all.data(), sizeof(Addr));
    node_of_.assign(world_, 0);                                 // node index of every rank (order of first appearance)
    {
      std::vector<std::string> seen;
      for (int p = 0; p < world_; ++p) {
        all[p].node[sizeof(all[p].node) - 1] = 0;
        auto it = std::find(seen.begin(), seen.end(), std::string(all[p].node));
        node_of_[p] = (int)(it - seen.begin());
        if (it == seen.end()) seen.push_back(all[p].node);
      }
    }
    // connect to every lower rank (the listen backlog completes the handshake even before the peer accepts) ...
    for (int p = 0; p < rank_; ++p) {
      int fd = tcp_connect_retry(all[p].ip, all[p].port, 60);
      tcp_tune(fd, ctx->env.net_sockbuf_kb);
      MeshHello me{(uint32_t)rank_, 0, tcp_job_token()};
      try {
        tcp_send_all(fd, &me, sizeof(me));
      } catch (const Error& e) {
        MLSLB_ASSERT(false, "data mesh: hello to rank %d failed: %s", p, e.what());
      }
      fds_[p] = fd;
    }
    // ... and accept every higher one
    for (int accepted = 0; accepted < world_ - 1 - rank_;) {
      pollfd pl{lfd, POLLIN, 0};
      const int pr = poll(&pl, 1, 120000);
      MLSLB_ASSERT(pr > 0, "data mesh: %d of the %d higher ranks never connected", world_ - 1 - rank_ - accepted, world_ - 1 - rank_);
      int fd = accept4(lfd, nullptr, nullptr, SOCK_CLOEXEC);
      MLSLB_ASSERT(fd >= 0, "accept(): %s", strerror(errno));
      tcp_tune(fd, ctx->env.net_sockbuf_kb);
      MeshHello who{0, 0, 0};
      bool ok = true;
      try {
        timeval tv{10, 0};                                       // a stray that never speaks must not stall the job
        setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
        tcp_recv_all(fd, &who, sizeof(who));
        timeval none{0, 0};
        setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &none, sizeof(none));
      } catch (const Error&) {
        ok = false;
      }
      // somebody who is not a higher rank of THIS job (another job on a recycled port, a scanner): turn it away, keep waiting
      if (!ok || who.token != tcp_job_token() || (int)who.rank <= rank_ || (int)who.rank >= world_ || fds_[who.rank] >= 0) {
        MLSLB_LOG(LOG_ERROR, "data mesh: turned away a connection (claims rank %u)", who.rank);
        close(fd);
        continue;
      }
      fds_[who.rank] = fd;
      ++accepted;
    }
    close(lfd);
    // ranks of one node talk through shared memory: the lower rank of a pair creates the segment and names it over the
    // socket, the higher one maps it and answers; a pair that cannot share it (containers with separate /dev/shm) stays on TCP
    if (ec.net_shm) {
      struct Offer {
        uint32_t magic;
        char name[60];
      };
      const size_t ring_bytes = ring_capacity(ec.net_shm_ring_kb);
      const size_t seg_bytes = 2 * (offsetof(ShmRing, data) + ring_bytes);
      auto ring_at = [&](char* base, int k) { return (ShmRing*)(base + (size_t)k * (offsetof(ShmRing, data) + ring_bytes)); };
      std::vector<std::pair<int, std::string>> offered;
      for (int p = rank_ + 1; p < world_; ++p) {
        if (strcmp(all[p].node, mine.node) != 0) continue;
        Offer o;
        memset(&o, 0, sizeof(o));
        o.magic = 0x4d53484d;
        snprintf(o.name, sizeof(o.name), "/mlslb_n%08x_%d_%d_%d", (unsigned)std::hash<std::string>()(key), (int)getpid(), rank_, p);
        int fd = shm_open(o.name, O_CREAT | O_EXCL | O_RDWR, 0600);
        char* base = nullptr;
        if (fd >= 0 && ftruncate(fd, (off_t)seg_bytes) == 0) {
          void* m = mmap(nullptr, seg_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
          if (m != MAP_FAILED) base = (char*)m;
        }
        if (fd >= 0) close(fd);
        if (base) {
          for (int k = 0; k < 2; ++k) {
            ShmRing* r = ring_at(base, k);
            r->head.store(0, std::memory_order_relaxed);
            r->tail.store(0, std::memory_order_relaxed);
            r->cap = ring_bytes;
          }
          std::atomic_thread_fence(std::memory_order_release);
          peers_[p].tx = ring_at(base, 0);      // lower -> higher
          peers_[p].rx = ring_at(base, 1);
          peers_[p].seg = base;
          peers_[p].seg_bytes = seg_bytes;
        } else {
          o.magic = 0;                          // tell the peer to stay on TCP
        }
        tcp_send_all(fds_[p], &o, sizeof(o));
        offered.emplace_back(p, base ? std::string(o.name) : std::string());
      }
      for (int p = 0; p < rank_; ++p) {
        if (strcmp(all[p].node, mine.node) != 0) continue;
        Offer o;
        tcp_recv_all(fds_[p], &o, sizeof(o));
        uint32_t ok = 0;
        if (o.magic == 0x4d53484d) {
          o.name[sizeof(o.name) - 1] = 0;
          int fd = shm_open(o.name, O_RDWR, 0600);
          if (fd >= 0) {
            void* m = mmap(nullptr, seg_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
            close(fd);
            if (m != MAP_FAILED && ring_at((char*)m, 0)->cap == ring_bytes) {
              peers_[p].rx = ring_at((char*)m, 0);
              peers_[p].tx = ring_at((char*)m, 1);
              peers_[p].seg = (char*)m;
              peers_[p].seg_bytes = seg_bytes;
              ok = 1;
            } else if (m != MAP_FAILED) {
              munmap(m, seg_bytes);
            }
          }
        }
        tcp_send_all(fds_[p], &ok, sizeof(ok));
      }
      for (auto& of : offered) {
        uint32_t ok = 0;
        tcp_recv_all(fds_[of.first], &ok, sizeof(ok));
        if (!of.second.empty()) shm_unlink(of.second.c_str());     // both sides hold their mapping (or never will)
        Peer& P = peers_[of.first];
        if (!ok && P.seg) {
          munmap(P.seg, P.seg_bytes);
          P.seg = nullptr;
          P.tx = P.rx = nullptr;
        }
      }
      for (int p = 0; p < world_; ++p)
        if (peers_[p].tx) {
          shm_peers_.push_back(p);
        }
    }
    for (int p = 0; p < world_; ++p)
      if (fds_[p] >= 0) {
        int fl = fcntl_nonblock(fds_[p]);
        (void)fl;
      }
    ctx->boot->barrier();
  }

  int shm_peer_count() const { return (int)shm_peers_.size(); }
  int node_of(int rank) const { return node_of_.empty() ? 0 : node_of_[rank]; }
  const std::string& address() const { return my_ip_; }

  void shutdown_all() {
    for (int& fd : fds_)
      if (fd >= 0) {
        close(fd);
        fd = -1;
      }
    for (Peer& P : peers_)
      if (P.seg) {
        munmap(P.seg, P.seg_bytes);
        P.seg = nullptr;
        P.tx = P.rx = nullptr;
      }
    shm_peers_.clear();
  }

  // Returns when every send and receive has completed.  Plain collectives name every peer at most once per direction; the
  // chunked ones post several messages per peer (own tag each): messages to one peer leave in list order, receives complete
  // in the order the peer sent them.  `on_recv(i)` runs when recvs[i] is complete (inside the progress loop: the sockets keep
  // flowing through the kernel's buffers while it computes) and may queue further sends with add_send().
  using RecvFn = std::function<void(size_t)>;
  void exchange(uint64_t tag, const std::vector<Seg>& sends, const std::vector<Seg>& recvs, const RecvFn* on_recv = nullptr) {
    std::lock_guard<std::mutex> g(mu_);
    Xchg x;
    x.tag = tag;
    x.on_recv = on_recv;
    struct Scope {
      Mesh* m;
      ~Scope() { m->cur_ = nullptr; }
    } scope{this};
    cur_ = &x;
    for (const Seg& s : sends) x.outs.push_back(Out{s.peer, WireHdr{s.tag ? s.tag : tag, s.bytes}, 0, 0, s.ptr});
    // expected arrivals; something that came early is already parked
    std::vector<size_t> early;
    std::vector<int> rpeers;
    for (size_t i = 0; i < recvs.size(); ++i) {
      const Seg& r = recvs[i];
      const uint64_t t = r.tag ? r.tag : tag;
      Peer& P = peers_[r.peer];
      if (std::find(rpeers.begin(), rpeers.end(), r.peer) == rpeers.end()) rpeers.push_back(r.peer);
      auto it = P.parked.find(t);
      if (it != P.parked.end()) {
        MLSLB_ASSERT(it->second.size() == r.bytes, "message of %zu bytes from rank %d where %zu were expected", it->second.size(),
                     r.peer, r.bytes);
        if (r.bytes) memcpy(r.ptr, it->second.data(), r.bytes);
        P.parked.erase(it);
        early.push_back(i);
      } else {
        MLSLB_ASSERT(P.expect.find(t) == P.expect.end(), "two receives from rank %d under one tag", r.peer);
        P.expect[t] = Expect{r.ptr, r.bytes, i};
        ++x.pending_in;
      }
    }
    // start with the next rank, not with rank 0: if everybody served the peers in index order, all first messages would
    // converge on the same receiver (stable: the messages of one peer keep their order)
    std::stable_sort(x.outs.begin(), x.outs.end(), [&](const Out& a, const Out& b) {
      return (a.peer - rank_ + world_) % world_ < (b.peer - rank_ + world_) % world_;
    });
    x.pending_out = x.outs.size();
    blocked_.assign(world_, 0);
    if (on_recv)
      for (size_t i : early) (*on_recv)(i);
    const uint64_t t0 = now_ns();
    // first try without sleeping: small messages usually go out and come in at once
    push_ready(x);
    for (int spin = 0; spin < 200 && x.pending_in; ++spin)
      for (int p : rpeers) {
        if (peers_[p].expect.empty()) continue;      // everything of this peer is in: no system call for it
        x.pending_in -= drain(p);
        if (x.fresh) push_ready(x);
      }
    std::vector<pollfd> pfds;
    uint64_t idle = 0;
    while (x.pending_in || x.pending_out) {
      push_ready(x);
      // shared-memory peers have no descriptor to sleep on: look at their rings every round, and keep the rounds short while
      // something is expected from / still has to go to one of them
      bool shm_busy = x.shm_stall;
      for (int p : shm_peers_) {
        Peer& P = peers_[p];
        if (P.expect.empty()) continue;
        shm_busy = true;
        if (P.rx->readable() && !(P.held && !P.expect.count(P.hdr.tag))) {
          x.pending_in -= drain(p);
        }
      }
      if (io_progress_) {          // bytes moved through a ring (or a socket) since the last round: not idle
        io_progress_ = false;
        idle = 0;
      }
      if (x.fresh) continue;
      if (!x.pending_in && !x.pending_out) break;
      long timeout_us = x.paced ? (tokens_ < 1.0 ? 1000 : 0) : 100000;
      if (shm_busy) {
        timeout_us = 0;
        if (++idle > 64) sched_yield();
        if (idle > 20000) timeout_us = 100;    // a slow peer: stop burning the core (the sockets still wake us at once)
      }
      pfds.clear();
      // always listen on every connection: a peer may already be sending for a later collective
      for (int p = 0; p < world_; ++p)
        if (fds_[p] >= 0) {
          const Peer& P = peers_[p];
          const bool idle_hold = P.held && !P.expect.count(P.hdr.tag);   // readable, but nobody to read for: don't spin
          if (P.rx) continue;                                           // (its socket is silent after the set-up)
          pfds.push_back(pollfd{fds_[p], (short)((idle_hold ? 0 : POLLIN) | (blocked_[p] ? POLLOUT : 0)), 0});
        }
      timespec ts{timeout_us / 1000000, (timeout_us % 1000000) * 1000};
      int rc = ppoll(pfds.data(), (nfds_t)pfds.size(), &ts, nullptr);
      if (rc < 0 && errno != EINTR) MLSLB_ASSERT(false, "poll(): %s", strerror(errno));
      if (ctx_->boot->poisoned()) MLSLB_ASSERT(false, "job poisoned by rank %d during a network collective", (int)ctx_->boot->poisoned() - 1);
      const int wd = ctx_->env.watchdog_sec;
      if (wd > 0 && now_ns() - t0 > (uint64_t)wd * 1000000000ull) {
        ctx_->boot->poison(rank_);
        MLSLB_ASSERT(false, "watchdog: network collective (tag %llx) did not complete in %d s", (unsigned long long)tag, wd);
      }
      if (rc <= 0) continue;
      idle = 0;
      for (pollfd& pf : pfds) {
        if (pf.revents & (POLLERR | POLLHUP | POLLNVAL)) {
          if (!(pf.revents & POLLIN)) MLSLB_ASSERT(false, "connection to a peer broke during a collective");
        }
        int peer = -1;
        for (int p = 0; p < world_; ++p)
          if (fds_[p] == pf.fd) peer = p;
        if (pf.revents & POLLOUT) blocked_[peer] = 0;
        if (pf.revents & POLLIN) x.pending_in -= drain(peer);
      }
    }
  }

  // Only inside an on_recv callback: one more message of the running exchange (e.g. the reduced chunk that can leave now).
  void add_send(const Seg& s) {
    MLSLB_ASSERT(cur_ != nullptr, "add_send() outside an exchange");
    cur_->outs.push_back(Out{s.peer, WireHdr{s.tag ? s.tag : cur_->tag, s.bytes}, 0, 0, s.ptr});
    ++cur_->pending_out;
    cur_->fresh = true;
  }

 private:
  struct Expect {
    char* ptr;
    size_t bytes;
    size_t idx;     // position in the exchange's receive list
  };
  struct Out {
    int peer;
    WireHdr hdr;
    size_t hdr_sent, sent;
    const char* ptr;
    bool done() const { return hdr_sent == sizeof(WireHdr) && sent == hdr.bytes; }
  };
  struct Xchg {     // the running exchange
    uint64_t tag = 0;
    std::vector<Out> outs;
    size_t first_live = 0, pending_out = 0, pending_in = 0;
    bool fresh = false;          // add_send() queued something since the last push
    bool paced = false;          // link emulation: the last push ran out of tokens
    bool shm_stall = false;      // a shared-memory ring was full at the last push
    const RecvFn* on_recv = nullptr;
  };
  struct Peer {
    // incoming parser
    WireHdr hdr;
    size_t hdr_got = 0, got = 0;
    char* dst = nullptr;                  // where the current payload goes (user buffer or parking space)
    std::vector<char> stash;              // parking space of the message being read, if nobody waits for it yet
    bool to_stash = false;
    bool held = false;                    // header read, large payload left in the socket until its receive is posted
    std::map<uint64_t, Expect> expect;    // tag -> waiting receive of the running exchange
    std::map<uint64_t, std::vector<char>> parked;
    std::vector<char> ra;                 // read-ahead of the socket (sock_read)
    size_t ra_lo = 0, ra_hi = 0;
    // same node: the byte streams of the pair run through shared memory instead of the socket
    ShmRing *tx = nullptr, *rx = nullptr;
    char* seg = nullptr;
    size_t seg_bytes = 0;
  };

  static int fcntl_nonblock(int fd);
  static size_t ring_capacity(long kb) {      // per direction of a same-node pair, rounded up to a power of two
    size_t cap = 4096;
    while (cap < ((size_t)std::max(4l, kb) << 10)) cap <<= 1;
    return cap;
  }
  // Same key = same machine and same launcher: host name + boot id, plus the node rank the launcher gave (several "nodes" on
  // one machine - the test set-up - stay separate).
  static std::string node_key(const std::string& node_rank) {
    char host[64] = "?";
    gethostname(host, sizeof(host) - 1);
    std::string k = host;
    if (FILE* f = fopen("/proc/sys/kernel/random/boot_id", "r")) {
      char id[48] = "";
      if (fgets(id, sizeof(id), f)) k += std::string(":") + std::string(id).substr(0, 8);
      fclose(f);
    }
    if (!node_rank.empty()) k += ":" + node_rank;
    return k;
  }
  size_t eager_bytes() const { return eager_bytes_; }

  // Send what the sockets take.  Per peer strictly in list order (the byte stream carries one message after the other); a
  // peer whose socket is full is skipped until poll() reports it writable again.
  void push_ready(Xchg& x) {
    x.fresh = x.paced = x.shm_stall = false;
    busy_.assign(world_, 0);
    while (x.first_live < x.outs.size() && x.outs[x.first_live].done()) ++x.first_live;
    for (size_t i = x.first_live; i < x.outs.size(); ++i) {
      Out& o = x.outs[i];
      if (o.done() || busy_[o.peer]) continue;
      if (blocked_[o.peer]) {
        busy_[o.peer] = 1;
        continue;
      }
      paced_ = false;
      quantum_left_ = 64 << 10;
      push(o.peer, o.hdr, o.hdr_sent, o.ptr, o.sent);
      if (o.done()) {
        --x.pending_out;
      } else if (paced_) {      // out of tokens or of its turn, not out of socket space: go on in a moment, nothing to poll for
        x.paced = true;
        busy_[o.peer] = 1;
        if (tokens_ < 1.0) break;
      } else if (peers_[o.peer].tx) {   // ring full: the reader is not there yet, look again in a moment (nothing to poll)
        busy_[o.peer] = 1;
        x.shm_stall = true;
      } else {
        busy_[o.peer] = blocked_[o.peer] = 1;
      }
    }
  }

  void push(int peer, const WireHdr& h, size_t& hdr_sent, const char* ptr, size_t& sent) {
    if (ShmRing* ring = peers_[peer].tx) {
      while (hdr_sent < sizeof(WireHdr)) {
        const size_t k = ring->write((const char*)&h + hdr_sent, sizeof(WireHdr) - hdr_sent);
        if (!k) return;
        hdr_sent += k;
        io_progress_ = true;
      }
      while (sent < h.bytes) {
        const size_t k = ring->write(ptr + sent, h.bytes - sent);
        if (!k) return;
        sent += k;
        io_progress_ = true;
      }
      return;
    }
    const int fd = fds_[peer];
    while (hdr_sent < sizeof(WireHdr) || sent < h.bytes) {
      // header and payload leave in one call (one segment for small messages instead of a 16-byte packet of its own)
      iovec iov[2];
      int niov = 0;
      if (hdr_sent < sizeof(WireHdr)) iov[niov++] = iovec{(char*)&h + hdr_sent, sizeof(WireHdr) - hdr_sent};
      size_t pay = h.bytes - sent;
      if (rate_Bps_ > 0) {      // link emulation: this rank's egress is paced by a token bucket (3 ms of burst)
        const uint64_t now = now_ns();
        tokens_ = std::min(rate_Bps_ * 3e-3, tokens_ + (double)(now - last_refill_ns_) * 1e-9 * rate_Bps_);
        last_refill_ns_ = now;
        if (tokens_ < 1.0) {
          paced_ = true;
          return;
        }
        if (quantum_left_ == 0) {    // the peers take turns, like sockets sharing one NIC
          paced_ = true;
          return;
        }
        pay = std::min({pay, (size_t)tokens_, quantum_left_});
      }
      if (pay) iov[niov++] = iovec{(char*)ptr + sent, pay};
      msghdr mh;
      memset(&mh, 0, sizeof(mh));
      mh.msg_iov = iov;
      mh.msg_iovlen = (size_t)niov;
      ssize_t n = sendmsg(fd, &mh, MSG_NOSIGNAL);
      if (n < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) return;
      if (n < 0 && errno == EINTR) continue;
      MLSLB_ASSERT(n > 0, "send() to rank %d: %s", peer, strerror(errno));
      size_t k = (size_t)n;
      if (hdr_sent < sizeof(WireHdr)) {
        const size_t h_part = std::min(k, sizeof(WireHdr) - hdr_sent);
        hdr_sent += h_part;
        k -= h_part;
      }
      sent += k;
      if (rate_Bps_ > 0) {
        tokens_ -= (double)k;
        quantum_left_ -= std::min(quantum_left_, k);
      }
    }
  }

  // recv() with read-ahead: a short read (a header, a small payload) fetches up to kReadAhead bytes from the socket in one
  // system call and serves what follows from there - header and payload of a small message cost one call instead of two,
  // several small messages one; long reads go straight to their destination.  Same return convention as recv().
  static constexpr size_t kReadAhead = 16 << 10;
  ssize_t sock_read(Peer& P, int fd, char* dst, size_t want) {
    if (P.ra_lo < P.ra_hi) {
      const size_t k = std::min(want, P.ra_hi - P.ra_lo);
      memcpy(dst, P.ra.data() + P.ra_lo, k);
      P.ra_lo += k;
      return (ssize_t)k;
    }
    if (want >= kReadAhead / 2) return recv(fd, dst, want, 0);
    if (P.ra.empty()) P.ra.resize(kReadAhead);
    const ssize_t n = recv(fd, P.ra.data(), kReadAhead, 0);
    if (n <= 0) return n;
    const size_t k = std::min(want, (size_t)n);
    memcpy(dst, P.ra.data(), k);
    P.ra_lo = k;
    P.ra_hi = (size_t)n;
    return (ssize_t)k;
  }

  // read what is available from `peer`; returns how many receives of the running exchange completed
  size_t drain(int peer) {
    Peer& P = peers_[peer];
    const int fd = fds_[peer];
    size_t completed = 0;
    for (;;) {
      if (P.hdr_got < sizeof(WireHdr)) {
        ssize_t n = P.rx ? (ssize_t)P.rx->read((char*)&P.hdr + P.hdr_got, sizeof(WireHdr) - P.hdr_got)
                         : sock_read(P, fd, (char*)&P.hdr + P.hdr_got, sizeof(WireHdr) - P.hdr_got);
        if (P.rx && n == 0) return completed;
        if (n < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) return completed;
        if (n < 0 && errno == EINTR) continue;
        MLSLB_ASSERT(n > 0, "rank %d closed its connection in the middle of a job", peer);
        P.hdr_got += (size_t)n;
        if (P.hdr_got < sizeof(WireHdr)) continue;
        P.got = 0;
        auto it = P.expect.find(P.hdr.tag);
        if (it != P.expect.end()) {
          MLSLB_ASSERT(it->second.bytes == P.hdr.bytes, "message of %llu bytes from rank %d where %zu were expected",
                       (unsigned long long)P.hdr.bytes, peer, it->second.bytes);
          P.dst = it->second.ptr;
          P.to_stash = false;
        } else if (P.hdr.bytes > eager_bytes()) {
          P.held = true;
        } else {
          P.stash.resize(P.hdr.bytes);
          P.dst = P.stash.data();
          P.to_stash = true;
        }
      }
      if (P.held) {
        auto it = P.expect.find(P.hdr.tag);
        if (it == P.expect.end()) return completed;
        MLSLB_ASSERT(it->second.bytes == P.hdr.bytes, "message of %llu bytes from rank %d where %zu were expected",
                     (unsigned long long)P.hdr.bytes, peer, it->second.bytes);
        P.dst = it->second.ptr;
        P.to_stash = false;
        P.held = false;
      }
      while (P.got < P.hdr.bytes) {
        ssize_t n = P.rx ? (ssize_t)P.rx->read(P.dst + P.got, P.hdr.bytes - P.got) : sock_read(P, fd, P.dst + P.got, P.hdr.bytes - P.got);
        if (P.rx && n == 0) return completed;
        if (n < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) return completed;
        if (n < 0 && errno == EINTR) continue;
        MLSLB_ASSERT(n > 0, "rank %d closed its connection in the middle of a message", peer);
        io_progress_ = true;
        P.got += (size_t)n;
      }
      // message complete
      if (P.to_stash) {
        // its collective may have started while the payload was still trickling in
        auto it = P.expect.find(P.hdr.tag);
        if (it != P.expect.end()) {
          MLSLB_ASSERT(it->second.bytes == P.hdr.bytes, "message of %llu bytes from rank %d where %zu were expected",
                       (unsigned long long)P.hdr.bytes, peer, it->second.bytes);
          if (P.hdr.bytes) memcpy(it->second.ptr, P.stash.data(), P.hdr.bytes);
          const size_t idx = it->second.idx;
          P.expect.erase(it);
          ++completed;
          if (cur_ && cur_->on_recv) (*cur_->on_recv)(idx);
        } else {
          P.parked[P.hdr.tag] = std::move(P.stash);
        }
        P.stash.clear();
        P.hdr_got = 0;
        P.dst = nullptr;
      } else {
        auto it = P.expect.find(P.hdr.tag);
        const size_t idx = it->second.idx;
        P.expect.erase(it);
        ++completed;
        P.hdr_got = 0;       // the parser is ready for the next message before the callback runs
        P.dst = nullptr;
        if (cur_ && cur_->on_recv) (*cur_->on_recv)(idx);
      }
      // nothing else is expected from this peer now: what it sends next belongs to a later collective and is read then
      // (saves the receive call that would only report an empty socket)
      if (P.expect.empty()) return completed;
    }
  }

  RankContext* ctx_ = nullptr;
  int rank_ = 0, world_ = 1;
  std::vector<int> fds_;
  std::vector<Peer> peers_;
  std::vector<char> blocked_, busy_;   // per peer: socket full (wait for POLLOUT) / already served in this push pass
  std::vector<int> shm_peers_, node_of_;
  std::string my_ip_;
  Xchg* cur_ = nullptr;
  // MLSL_NET_EMULATE_GBIT=<x>: pace this rank's egress to x Gbit/s - what a collective does on a slower link than loop-back
  // can be measured on one machine (bench / test knob, off by default)
  double rate_Bps_ = 0.0, tokens_ = 0.0;
  size_t eager_bytes_ = kEagerBytes;
  uint64_t last_refill_ns_ = 0;
  bool paced_ = false, io_progress_ = false;
  size_t quantum_left_ = 0;
  std::mutex mu_;
};

int Mesh::fcntl_nonblock(int fd) {
  int fl = fcntl(fd, F_GETFL, 0);
  return fcntl(fd, F_SETFL, fl | O_NONBLOCK);
}

struct NetReqState {
  std::vector<float> residual;
};

class NetBackend final : public Backend {
 public:
  explicit NetBackend(RankContext* ctx) : ctx_(ctx) { mesh_.init(ctx); }
  const char* name() const override { return "net"; }
  void* alloc(size_t bytes, size_t align) override {
    void* p = aligned_host_alloc(bytes, align, ctx_->env.thp_threshold_mb << 20);
    MLSLB_ASSERT(p != nullptr, "allocation of %zu bytes failed", bytes);
    ctx_->ptrcheck.add(p, bytes);
    return p;
  }
  void free(void* p) override {
    if (!p) return;
    ctx_->ptrcheck.remove(p);
    ::free(p);
  }
  bool owns(const void*, size_t) const override { return true; }   // any host buffer can go on the wire
  void prepare(CommRequest&) override {}
  void release(CommRequest& r) override {   // the error-feedback residual of a quantised request lives as long as the request
    delete (NetReqState*)r.backend_state;
    r.backend_state = nullptr;
  }
  void launch(CommRequest& r) override {
    execute(r);
    r.state.store(CommRequest::DONE, std::memory_order_release);
  }
  bool test(CommRequest& r) override { return r.state.load(std::memory_order_acquire) >= CommRequest::DONE; }
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
  uint64_t heap_offset(const void*) const override {
    MLSLB_ASSERT(false, "RMA windows need a shared address space: not available on the net backend");
    return 0;
  }
  void* peer_heap_ptr(int, uint64_t) override {
    MLSLB_ASSERT(false, "RMA windows need a shared address space: not available on the net backend");
    return nullptr;
  }
  void finalize() override {
    try {
      ctx_->boot->barrier();
    } catch (const std::exception&) {
    }
    mesh_.shutdown_all();
  }
  std::string describe() const override {
    return "net backend (TCP mesh, " + std::to_string(ctx_->world) + " ranks, " + std::to_string(mesh_.shm_peer_count()) +
           " same-node peers over shared memory" + (mesh_.address().empty() ? "" : ", data address " + mesh_.address()) + ")";
  }

 private:
  RankContext* ctx_;
  Mesh mesh_;
  void execute(CommRequest& r);
  void quantized_allreduce(CommRequest& r, const ProcessGroup& g);
  void quantized_exchange(const float* x, float* y, size_t n, const std::vector<int>& ranks, int me, std::vector<float>& residual,
                          float scale, const std::function<uint64_t(int)>& tag, int step0);
  bool hierarchical_allreduce(CommRequest& r, const ProcessGroup& g, const std::function<uint64_t(int)>& tag);
  bool hierarchical_gather_scatter(CommRequest& r, const ProcessGroup& g, const std::function<uint64_t(int)>& tag);
  bool hierarchical_bcast(CommRequest& r, const ProcessGroup& g, const std::function<uint64_t(int)>& tag);
  struct NodeMap;
  void hierarchical_allreduce_pipelined(CommRequest& r, const ProcessGroup& g, const std::function<uint64_t(int)>& tag, const NodeMap& nm,
                                        size_t piece_bytes);
  void hier_allgather(const NodeMap& nm, const ProcessGroup& g, const std::function<uint64_t(int)>& tag, int step0, const char* S, char* R,
                      size_t blk);
  void hier_reduce_scatter(const NodeMap& nm, const ProcessGroup& g, const std::function<uint64_t(int)>& tag, int step0, DType dtype, RedOp rop,
                           const char* S, char* R, size_t n, float scale, int step_end = 256);
  // Where the members of a group run: N nodes with L members each (member positions per node, in member order), and this
  // rank's place.  false unless the group is regular (same L > 1 on each of N > 1 nodes).
  struct NodeMap {
    int N = 0, L = 0, my_node = 0, li = 0;
    std::vector<std::vector<int>> on_node;
  };
  bool node_map(const ProcessGroup& g, NodeMap& m) const {
    std::vector<int> nodes;
    for (int p = 0; p < g.size(); ++p) {
      const int nd = mesh_.node_of(g.members[p]);
      size_t k = std::find(nodes.begin(), nodes.end(), nd) - nodes.begin();
      if (k == nodes.size()) {
        nodes.push_back(nd);
        m.on_node.emplace_back();
      }
      m.on_node[k].push_back(p);
    }
    m.N = (int)nodes.size();
    m.L = (int)m.on_node[0].size();
    if (m.N < 2 || m.L < 2) return false;
    for (auto& v : m.on_node)
      if ((int)v.size() != m.L) return false;
    for (int k = 0; k < m.N; ++k)
      for (int j = 0; j < m.L; ++j)
        if (m.on_node[k][j] == g.idx) {
          m.my_node = k;
          m.li = j;
        }
    return true;
  }
};

// Receive space for the slices of a reduction: grow-only and uninitialised (a fresh std::vector would zero-fill and
// page-fault the whole message size on every call).
static char* net_scratch(size_t bytes) {
  thread_local std::unique_ptr<char[]> buf;
  thread_local size_t cap = 0;
  if (bytes > cap) {
    cap = bytes + bytes / 4;
    buf.reset(new char[cap]);
  }
  return buf.get();
}

// (the quantised exchange has its own: it runs inside the two-level all-reduce, whose shard lives in the scratch above)
static char* net_scratch_q(size_t bytes) {
  thread_local std::unique_ptr<char[]> buf;
  thread_local size_t cap = 0;
  if (bytes > cap) {
    cap = bytes + bytes / 4;
    buf.reset(new char[cap]);
  }
  return buf.get();
}

void NetBackend::execute(CommRequest& r) {
  const CommDesc& d = r.desc;
  ProcessGroup* gp = d.group;
  const size_t dt = dtype_size(d.dtype), n = d.count;
  char* S = (char*)r.send;
  char* R = (char*)r.recv;
  if (!gp || gp->size() <= 1) {   // local semantics, like the other backends
    size_t bytes = 0;
    switch (d.kind) {
      case OpKind::ALLREDUCE: case OpKind::REDUCE: case OpKind::REDUCE_SCATTER: case OpKind::ALLGATHER: case OpKind::GATHER:
      case OpKind::SCATTER: case OpKind::ALLTOALL: case OpKind::ALLGATHERV:
        bytes = n * dt;
        break;
      default: break;
    }
    if (bytes && R && R != S) memcpy(R, S, bytes);
    if ((d.kind == OpKind::ALLTOALLV || d.kind == OpKind::SENDRECV_LIST) && !d.send_counts.empty() && d.send_counts[0])
      memmove(R + d.recv_offsets[0] * dt, S + d.send_offsets[0] * dt, d.send_counts[0] * dt);
    if (d.kind == OpKind::FUSED_UPDATE) {
      const DType pdt = d.has_out_dtype ? d.out_dtype : d.dtype;
      std::vector<float> gs(n);
      for (size_t i = 0; i < n; ++i) gs[i] = (d.dtype == DType::F32 ? ((const float*)S)[i] : bf16_to_f32(((const uint16_t*)S)[i])) * d.scale;
      host_optimizer_step(d.fused, pdt, (char*)d.fused.param, gs.data(), n);
    } else if (d.scale != 1.0f && R && (d.kind == OpKind::ALLREDUCE || d.kind == OpKind::REDUCE_SCATTER)) {
      std::vector<const void*> one{R};
      host_reduce(d.dtype, R, one, n, RedOp::SUM, d.scale);
    }
    return;
  }
  const ProcessGroup& g = *gp;
  const int P = g.size(), me = g.idx;
  auto tag = [&](int step) {
    return ((uint64_t)(uint8_t)g.row << 56) | ((uint64_t)(r.lane & 0xff) << 48) | ((r.group_seq & 0xffffffffffull) << 8) |
           (uint64_t)(step & 0xff);
  };
  auto peer = [&](int p) { return g.members[p]; };
  std::vector<Seg> snd, rcv;
  auto go = [&](int step) {
    mesh_.exchange(tag(step), snd, rcv);
    snd.clear();
    rcv.clear();
  };
  // reduce `P` equally sized slices sitting in tmp (slice p from member p; my own contribution read from `own`)
  // (host_reduce works element by element in a fixed member order: `dst` may be the same memory as `own`)
  auto reduce_slices = [&](char* dst, const char* own, const char* tmp, size_t elems, float scale) {
    std::vector<const void*> srcs(P);
    for (int p = 0; p < P; ++p) srcs[p] = p == me ? (const void*)own : (const void*)(tmp + (size_t)p * elems * dt);
    if (elems) host_reduce(d.dtype, dst, srcs, elems, d.rop, scale);
  };

  // Large reductions travel in chunks (the reference cuts messages of 128 MiB and more into epSize x MLSL_LARGE_MSG_CHUNKS
  // independent non-blocking requests and raises the chunk count to 128 on Ethernet: reference src/comm_ep.cpp:96-97,649-653,
  // src/mlsl.cpp:667; here the pieces additionally form a pipeline): every slice is cut into K pieces with a tag each, a piece is reduced as soon as all
  // members' copies of it are there (while the later pieces are still on the wire) and - all-reduce - its result leaves for
  // the other members right away, so the reduction and the second exchange hide behind the first.  Pieces are reduced in
  // member order like whole slices: the values do not depend on the chunking.
  const size_t chunk_bytes = std::max<size_t>(4096, (size_t)std::max(0l, ctx_->env.net_chunk_kb) << 10);
  constexpr int kMaxChunks = 96;                       // two phases of tags fit the 8-bit step field (2 + 2 * 96 < 256)
  auto chunk_elems = [&](size_t slice_elems) {         // elements per piece (0: not worth chunking)
    if (slice_elems * dt < 2 * chunk_bytes) return (size_t)0;
    size_t ce = std::max(chunk_bytes / dt, ceil_div(slice_elems, (size_t)kMaxChunks));
    return (ce + 63) & ~(size_t)63;
  };
  // slice p of the send side starts at element s_lo(p) and has s_len(p) elements; mine is reduced into `out`; with
  // `gather` the reduced pieces go to everybody (all-reduce: slice p of R starts at s_lo(p) too)
  auto chunked_reduce = [&](size_t ce, size_t per, const std::function<size_t(int)>& s_lo, const std::function<size_t(int)>& s_len,
                            char* out, bool gather) {
    char* tmp = net_scratch((size_t)(P + 1) * per * dt);   // (+1: room for a staged output behind the slices)
    auto clen = [&](int p, int c) {                    // length of piece c of slice p
      const size_t L = s_len(p), b = (size_t)c * ce;
      return b >= L ? (size_t)0 : std::min(ce, L - b);
    };
    const int Kall = (int)ceil_div(per, ce);
    struct Meta {
      int phase, p, c;
    };
    std::vector<Meta> meta;
    std::vector<int> arrived(Kall, 0);
    for (int p = 0; p < P; ++p) {
      if (p == me) continue;
      for (int c = 0; c < Kall; ++c) {
        if (clen(p, c)) snd.push_back(Seg{peer(p), S + (s_lo(p) + (size_t)c * ce) * dt, clen(p, c) * dt, tag(2 + c)});
        if (clen(me, c)) {
          rcv.push_back(Seg{peer(p), tmp + ((size_t)p * per + (size_t)c * ce) * dt, clen(me, c) * dt, tag(2 + c)});
          meta.push_back(Meta{0, p, c});
        }
      }
      if (gather)
        for (int c = 0; c < Kall; ++c)
          if (clen(p, c)) {
            rcv.push_back(Seg{peer(p), R + (s_lo(p) + (size_t)c * ce) * dt, clen(p, c) * dt, tag(2 + kMaxChunks + c)});
            meta.push_back(Meta{1, p, c});
          }
    }
    std::vector<const void*> srcs(P);
    auto reduce_piece = [&](int c) {
      const size_t off = (size_t)c * ce;
      for (int p = 0; p < P; ++p)
        srcs[p] = p == me ? (const void*)(S + (s_lo(me) + off) * dt) : (const void*)(tmp + ((size_t)p * per + off) * dt);
      host_reduce(d.dtype, out + off * dt, srcs, clen(me, c), d.rop, d.scale);
      if (gather)
        for (int p = 0; p < P; ++p)
          if (p != me) mesh_.add_send(Seg{peer(p), out + off * dt, clen(me, c) * dt, tag(2 + kMaxChunks + c)});
    };
    Mesh::RecvFn on_recv = [&](size_t i) {
      const Meta& m = meta[i];
      if (m.phase == 0 && ++arrived[m.c] == P - 1) reduce_piece(m.c);
    };
    mesh_.exchange(tag(0), snd, rcv, &on_recv);
    snd.clear();
    rcv.clear();
  };

  switch (d.kind) {
    case OpKind::BARRIER:
      if (P >= kTreeMinRanks) {
        // dissemination: in round k everybody signals the member 2^k places ahead and waits for the one 2^k places behind;
        // after ceil(log2 P) rounds every member has (transitively) heard from every other - log P messages per rank
        // instead of P - 1
        for (int k = 0, dist = 1; dist < P; ++k, dist <<= 1) {
          snd.push_back(Seg{peer((me + dist) % P), nullptr, 0});
          rcv.push_back(Seg{peer((me - dist + P) % P), nullptr, 0});
          go(k);
        }
        break;
      }
      for (int p = 0; p < P; ++p)
        if (p != me) {
          snd.push_back(Seg{peer(p), nullptr, 0});
          rcv.push_back(Seg{peer(p), nullptr, 0});
        }
      go(0);
      break;
    case OpKind::BCAST: {
      const int root = (int)d.root;
      if (hierarchical_bcast(r, g, tag)) break;
      if (n * dt < kBcastSplitBytes && P >= kTreeMinRanks) {
        // small, four members or more: binomial tree.  Member v (counted from the root) gets the buffer from v minus its
        // lowest set bit and passes it to v + 2^k for every 2^k below that bit, farthest first: the root's connection
        // carries log2 P copies instead of P - 1 and the copies of one level travel at the same time
        const int v = (me - root + P) % P;
        int low = 1;
        if (v)
          while (!(v & low)) low <<= 1;
        else
          while (low < P) low <<= 1;
        std::vector<Seg> kids;
        for (int b = low >> 1; b >= 1; b >>= 1)
          if (v + b < P) kids.push_back(Seg{peer((v + b + root) % P), R, n * dt, tag(1)});
        if (v) {      // one exchange: the copies for the children are queued the moment the parent's message is complete
          rcv.push_back(Seg{peer((v - low + root) % P), R, n * dt, tag(1)});
          Mesh::RecvFn forward = [&](size_t) {
            for (const Seg& k : kids) mesh_.add_send(k);
          };
          mesh_.exchange(tag(0), snd, rcv, &forward);
          rcv.clear();
        } else {
          snd = kids;
          go(0);
        }
        break;
      }
      if (n * dt < kBcastSplitBytes || P < 3) {   // small: the root sends the whole buffer to everybody
        if (me == root) {
          for (int p = 0; p < P; ++p)
            if (p != me) snd.push_back(Seg{peer(p), R, n * dt});
        } else {
          rcv.push_back(Seg{peer(root), R, n * dt});
        }
        go(0);
        break;
      }
      // large: the root scatters 1/P slices, then the owners all-gather them - every link carries (P-1)/P of the
      // message instead of the root's link carrying P-1 copies
      const size_t per = ceil_div(n, (size_t)P);
      auto lo = [&](int p) { return std::min(n, (size_t)p * per); };
      auto len = [&](int p) { return std::min(n, lo(p) + per) - lo(p); };
      if (me == root) {
        for (int p = 0; p < P; ++p)
          if (p != me && len(p)) snd.push_back(Seg{peer(p), R + lo(p) * dt, len(p) * dt});
      } else if (len(me)) {
        rcv.push_back(Seg{peer(root), R + lo(me) * dt, len(me) * dt});
      }
      go(0);
      for (int p = 0; p < P; ++p)
        if (p != me) {
          if (p != root && len(me)) snd.push_back(Seg{peer(p), R + lo(me) * dt, len(me) * dt});
          if (me != root && len(p)) rcv.push_back(Seg{peer(p), R + lo(p) * dt, len(p) * dt});
        }
      go(1);
      break;
    }
    case OpKind::ALLGATHER:
    case OpKind::ALLGATHERV: {
      if (d.kind == OpKind::ALLGATHER && hierarchical_gather_scatter(r, g, tag)) break;
      std::vector<size_t> cnt(P, n), off(P, 0);
      if (d.kind == OpKind::ALLGATHERV) cnt.assign(d.recv_counts.begin(), d.recv_counts.end());
      for (int p = 1; p < P; ++p) off[p] = off[p - 1] + cnt[p - 1];
      if (R + off[me] * dt != S) memmove(R + off[me] * dt, S, cnt[me] * dt);
      for (int p = 0; p < P; ++p)
        if (p != me) {
          snd.push_back(Seg{peer(p), R + off[me] * dt, cnt[me] * dt});
          rcv.push_back(Seg{peer(p), R + off[p] * dt, cnt[p] * dt});
        }
      go(0);
      break;
    }
    case OpKind::ALLTOALL:
    case OpKind::ALLTOALLV:
    case OpKind::SENDRECV_LIST: {
      const bool v = d.kind != OpKind::ALLTOALL;
      auto sc = [&](int p) { return v ? d.send_counts[p] : n; };
      auto so = [&](int p) { return v ? d.send_offsets[p] : (size_t)p * n; };
      auto rc = [&](int p) { return v ? d.recv_counts[p] : n; };
      auto ro = [&](int p) { return v ? d.recv_offsets[p] : (size_t)p * n; };
      std::vector<char> stage;   // in place: outgoing data must survive the incoming writes
      const char* src = S;
      if (S == R) {
        size_t total = 0;
        for (int p = 0; p < P; ++p) total = std::max(total, so(p) + sc(p));
        stage.assign(S, S + total * dt);
        src = stage.data();
      }
      if (sc(me)) memmove(R + ro(me) * dt, src + so(me) * dt, sc(me) * dt);
      for (int p = 0; p < P; ++p)
        if (p != me) {
          if (sc(p) || !v) snd.push_back(Seg{peer(p), (char*)src + so(p) * dt, sc(p) * dt});
          if (rc(p) || !v) rcv.push_back(Seg{peer(p), R + ro(p) * dt, rc(p) * dt});
        }
      go(0);
      break;
    }
    case OpKind::GATHER:
      if (me == (int)d.root) {
        memmove(R + (size_t)me * n * dt, S, n * dt);
        for (int p = 0; p < P; ++p)
          if (p != me) rcv.push_back(Seg{peer(p), R + (size_t)p * n * dt, n * dt});
      } else {
        snd.push_back(Seg{peer((int)d.root), S, n * dt});
      }
      go(0);
      break;
    case OpKind::SCATTER:
      if (me == (int)d.root) {
        for (int p = 0; p < P; ++p)
          if (p != me) snd.push_back(Seg{peer(p), S + (size_t)p * n * dt, n * dt});
        memmove(R, S + (size_t)me * n * dt, n * dt);
      } else {
        rcv.push_back(Seg{peer((int)d.root), R, n * dt});
      }
      go(0);
      break;
    case OpKind::REDUCE_SCATTER: {
      if (hierarchical_gather_scatter(r, g, tag)) break;
      if (const size_t ce = chunk_elems(n)) {      // (the choice must not depend on the rank or on its buffers)
        // in place the result lands on slice 0 of the send buffer, which a rank other than member 0 still has to send: only
        // an output on the own slice or outside the send buffer can be written while the exchange runs - else it is staged
        const bool out_safe = R == S + (size_t)me * n * dt || R + n * dt <= S || R >= S + (size_t)P * n * dt;
        char* out = out_safe ? R : net_scratch((size_t)(P + 1) * n * dt) + (size_t)P * n * dt;
        chunked_reduce(ce, n, [&](int p) { return (size_t)p * n; }, [&](int) { return n; }, out, false);
        if (!out_safe) memcpy(R, out, n * dt);
        break;
      }
      char* tmp = net_scratch((size_t)P * n * dt);
      for (int p = 0; p < P; ++p)
        if (p != me) {
          snd.push_back(Seg{peer(p), S + (size_t)p * n * dt, n * dt});
          rcv.push_back(Seg{peer(p), tmp + (size_t)p * n * dt, n * dt});
        }
      go(0);
      // in place (R == S) the result lands on slice 0 of the send buffer: either exactly on my input slice (me == 0) or
      // on a slice that has already been sent and is not an input of this reduction
      reduce_slices(R, S + (size_t)me * n * dt, tmp, n, d.scale);
      break;
    }
    case OpKind::REDUCE: {
      if (me == (int)d.root) {
        char* tmp = net_scratch((size_t)P * n * dt);
        for (int p = 0; p < P; ++p)
          if (p != me) rcv.push_back(Seg{peer(p), tmp + (size_t)p * n * dt, n * dt});
        go(0);
        reduce_slices(R, S, tmp, n, 1.0f);
      } else {
        snd.push_back(Seg{peer((int)d.root), S, n * dt});
        go(0);
      }
      break;
    }
    case OpKind::ALLREDUCE: {
      if (d.compress && d.dtype == DType::F32 && d.rop == RedOp::SUM) {
        if (hierarchical_allreduce(r, g, tag)) break;     // exact inside the node, quantised between the nodes
        quantized_allreduce(r, g);
        break;
      }
      const size_t one_shot = (size_t)std::max(0l, ctx_->env.net_oneshot_kb) << 10;
      if (hierarchical_allreduce(r, g, tag)) break;
      if (n * dt <= one_shot) {
        // small message: everybody sends the whole vector to everybody and reduces locally in member order (bitwise
        // identical everywhere) - one exchange instead of two
        char* tmp = net_scratch((size_t)P * n * dt);
        for (int p = 0; p < P; ++p)
          if (p != me) {
            snd.push_back(Seg{peer(p), S, n * dt});
            rcv.push_back(Seg{peer(p), tmp + (size_t)p * n * dt, n * dt});
          }
        go(0);
        reduce_slices(R, S, tmp, n, d.scale);
        break;
      }
      // reduce-scatter over ceil(n / P) sized slices, then all-gather of the reduced slices
      const size_t per = ceil_div(n, (size_t)P);
      auto lo = [&](int p) { return std::min(n, (size_t)p * per); };
      auto len = [&](int p) { return std::min(n, lo(p) + per) - lo(p); };
      if (const size_t ce = chunk_elems(per)) {
        chunked_reduce(ce, per, lo, len, R + lo(me) * dt, true);
        break;
      }
      char* tmp = net_scratch((size_t)P * per * dt);
      for (int p = 0; p < P; ++p)
        if (p != me) {
          snd.push_back(Seg{peer(p), S + lo(p) * dt, len(p) * dt});
          rcv.push_back(Seg{peer(p), tmp + (size_t)p * per * dt, len(me) * dt});
        }
      go(0);
      {
        std::vector<const void*> srcs(P);
        for (int p = 0; p < P; ++p) srcs[p] = p == me ? (const void*)(S + lo(me) * dt) : (const void*)(tmp + (size_t)p * per * dt);
        if (len(me)) host_reduce(d.dtype, R + lo(me) * dt, srcs, len(me), d.rop, d.scale);
      }
      for (int p = 0; p < P; ++p)
        if (p != me) {
          snd.push_back(Seg{peer(p), R + lo(me) * dt, len(me) * dt});
          rcv.push_back(Seg{peer(p), R + lo(p) * dt, len(p) * dt});
        }
      go(1);
      break;
    }
    case OpKind::FUSED_UPDATE: {
      MLSLB_ASSERT(d.dtype == DType::F32 || d.dtype == DType::BF16, "fused update: gradient dtype must be f32/bf16");
      const DType pdt = d.has_out_dtype ? d.out_dtype : d.dtype;
      MLSLB_ASSERT(pdt == DType::F32 || pdt == DType::BF16, "fused update: parameter dtype must be f32/bf16");
      const size_t pdts = dtype_size(pdt);
      {
        // groups with several members per node: the same reduce-scatter -> shared optimizer step -> all-gather, each in two
        // levels.  bf16 gradients are widened first: the sums are fp32 on both levels like in the flat form, and the wire
        // between the nodes still carries fewer bytes (fp32 partial sums of 1/L of the slices instead of every bf16 slice)
        NodeMap nm;
        const long hier_kb = ctx_->env.net_hier_kb;
        if (hier_kb >= 0 && (size_t)P * n * dt >= (size_t)hier_kb << 10 && n && node_map(g, nm)) {
          std::vector<float> gsum(n), wide;
          const char* grads = S;
          if (d.dtype == DType::BF16) {
            wide.resize((size_t)P * n);
            const uint16_t* h = (const uint16_t*)S;
            for (size_t i = 0; i < (size_t)P * n; ++i) wide[i] = bf16_to_f32(h[i]);
            grads = (const char*)wide.data();
          }
          hier_reduce_scatter(nm, g, tag, 100, DType::F32, RedOp::SUM, grads, (char*)gsum.data(), n, d.scale, 209);
          char* param = (char*)d.fused.param;
          host_optimizer_step(d.fused, pdt, param + (size_t)me * n * pdts, gsum.data(), n);
          hier_allgather(nm, g, tag, 209, param + (size_t)me * n * pdts, param, n * pdts);
          break;
        }
      }
      char* tmp = net_scratch((size_t)P * n * dt);
      for (int p = 0; p < P; ++p)
        if (p != me) {
          snd.push_back(Seg{peer(p), S + (size_t)p * n * dt, n * dt});
          rcv.push_back(Seg{peer(p), tmp + (size_t)p * n * dt, n * dt});
        }
      go(0);
      std::vector<float> gsum(n);
      if (d.dtype == DType::F32) {            // the vectorised reduction (same member order as the loop below)
        std::vector<const void*> srcs(P);
        for (int p = 0; p < P; ++p) srcs[p] = p == me ? (const void*)(S + (size_t)me * n * dt) : (const void*)(tmp + (size_t)p * n * dt);
        if (n) host_reduce(DType::F32, gsum.data(), srcs, n, RedOp::SUM, d.scale);
      } else {
        for (size_t i = 0; i < n; ++i) {
          float a = 0.f;
          for (int p = 0; p < P; ++p) {
            const char* sp = (p == me ? S + (size_t)me * n * dt : tmp + (size_t)p * n * dt) + i * dt;
            a += bf16_to_f32(*(const uint16_t*)sp);
          }
          gsum[i] = a * d.scale;
        }
      }
      char* param = (char*)d.fused.param;
      host_optimizer_step(d.fused, pdt, param + (size_t)me * n * pdts, gsum.data(), n);
      for (int p = 0; p < P; ++p)
        if (p != me) {
          snd.push_back(Seg{peer(p), param + (size_t)me * n * pdts, n * pdts});
          rcv.push_back(Seg{peer(p), param + (size_t)p * n * pdts, n * pdts});
        }
      go(1);
      break;
    }
    case OpKind::GEMM_RS:
    case OpKind::AG_GEMM:
      MLSLB_ASSERT(false, "%s is a device-only fused op", opkind_name(d.kind));
      break;
  }
}

// Two-level all-reduce for groups that have the same number L > 1 of members on each of N > 1 nodes: reduce-scatter among
// the members of a node (shared-memory rings), all-reduce of the 1/L shard among the members with the same local index
// (one per node, over the wire), all-gather inside the node.  A rank then sends 2 (n / L)(N - 1) / N bytes between nodes
// instead of the 2 n (P - L) / P of the flat exchange - L times less on the link that is the slow one.  (The reference
// leaves this to the MPI library underneath; Intel MPI's shm + fabric collectives are topology aware in the same way.)
// Every element is reduced along one fixed chain (node members in member order, then nodes in order) by exactly one rank
// and then copied, so all ranks end up with identical bits.  Returns false when the group / size does not qualify; the
// decision only uses facts every member knows (who runs where, the count), so all members decide alike.
bool NetBackend::hierarchical_allreduce(CommRequest& r, const ProcessGroup& g, const std::function<uint64_t(int)>& tag) {
  const long hier_kb = ctx_->env.net_hier_kb;   // < 0: never
  const CommDesc& d = r.desc;
  const size_t dt = dtype_size(d.dtype), n = d.count;
  if (hier_kb < 0 || n * dt < (size_t)hier_kb << 10) return false;
  NodeMap nm;
  if (!node_map(g, nm) || n < (size_t)nm.L * nm.N) return false;
  const int N = nm.N, L = nm.L, my_node = nm.my_node, li = nm.li;
  const std::vector<std::vector<int>>& on_node = nm.on_node;
  char* S = (char*)r.send;
  char* R = (char*)r.recv;
  auto peer = [&](int p) { return g.members[p]; };
  const size_t per_l = ceil_div(n, (size_t)L);                       // slice of local member j: [lo_l(j), lo_l(j) + len_l(j))
  auto lo_l = [&](int j) { return std::min(n, (size_t)j * per_l); };
  auto len_l = [&](int j) { return std::min(n, lo_l(j) + per_l) - lo_l(j); };
  const size_t mine = len_l(li);
  const size_t per_n = ceil_div(per_l, (size_t)N);                   // sub-slice of node k inside a shard
  auto lo_n = [&](size_t shard_len, int k) { return std::min(shard_len, (size_t)k * per_n); };
  auto len_n = [&](size_t shard_len, int k) { return std::min(shard_len, lo_n(shard_len, k) + per_n) - lo_n(shard_len, k); };
  const bool quant = d.compress && d.dtype == DType::F32 && d.rop == RedOp::SUM;
  const size_t piece_bytes = std::max<size_t>(4096, (size_t)std::max(0l, ctx_->env.net_chunk_kb) << 10);
  if (!quant && per_l * dt >= 2 * piece_bytes && ctx_->env.net_hier_pipeline) {
    hierarchical_allreduce_pipelined(r, g, tag, nm, piece_bytes);
    return true;
  }
  char* scratch = net_scratch(((size_t)L * per_l + per_l + (size_t)N * per_n) * dt);
  char* tmpA = scratch;                                              // L slices received inside the node
  char* shard = scratch + (size_t)L * per_l * dt;                    // my 1/L of the node's sum, then of the total
  char* tmpB = shard + per_l * dt;                                   // N sub-slices received from the other nodes
  std::vector<Seg> snd, rcv;
  std::vector<const void*> srcs;
  // A: reduce-scatter inside the node
  for (int j = 0; j < L; ++j) {
    if (j == li) continue;
    const int p = on_node[my_node][j];
    if (len_l(j)) snd.push_back(Seg{peer(p), S + lo_l(j) * dt, len_l(j) * dt});
    if (mine) rcv.push_back(Seg{peer(p), tmpA + (size_t)j * per_l * dt, mine * dt});
  }
  mesh_.exchange(tag(200), snd, rcv);
  snd.clear();
  rcv.clear();
  if (mine) {
    srcs.assign(L, nullptr);
    for (int j = 0; j < L; ++j) srcs[j] = j == li ? (const void*)(S + lo_l(li) * dt) : (const void*)(tmpA + (size_t)j * per_l * dt);
    host_reduce(d.dtype, shard, srcs, mine, d.rop, 1.0f);
  }
  // B: all-reduce of the shard among the members with my local index, one per node
  if (d.compress && d.dtype == DType::F32 && d.rop == RedOp::SUM) {
    // gradient compression where it pays - on the wire between nodes: block-scaled fp8 with error feedback for the shard
    // (the node-local sums before and the copies after are exact); the residual belongs to the request as in the flat form
    if (!r.backend_state) r.backend_state = new NetReqState();
    NetReqState* st = (NetReqState*)r.backend_state;
    std::vector<int> column(N);
    for (int k = 0; k < N; ++k) column[k] = peer(on_node[k][li]);
    if (mine) quantized_exchange((const float*)shard, (float*)shard, mine, column, my_node, st->residual, d.scale, tag, 201);
  } else {
  const size_t sub = len_n(mine, my_node);
  for (int k = 0; k < N; ++k) {
    if (k == my_node) continue;
    const int p = on_node[k][li];
    if (len_n(mine, k)) snd.push_back(Seg{peer(p), shard + lo_n(mine, k) * dt, len_n(mine, k) * dt});
    if (sub) rcv.push_back(Seg{peer(p), tmpB + (size_t)k * per_n * dt, sub * dt});
  }
  mesh_.exchange(tag(201), snd, rcv);
  snd.clear();
  rcv.clear();
  if (sub) {
    srcs.assign(N, nullptr);
    for (int k = 0; k < N; ++k)
      srcs[k] = k == my_node ? (const void*)(shard + lo_n(mine, my_node) * dt) : (const void*)(tmpB + (size_t)k * per_n * dt);
    host_reduce(d.dtype, shard + lo_n(mine, my_node) * dt, srcs, sub, d.rop, d.scale);     // (element-wise: in place is fine)
  }
  for (int k = 0; k < N; ++k) {
    if (k == my_node) continue;
    const int p = on_node[k][li];
    if (sub) snd.push_back(Seg{peer(p), shard + lo_n(mine, my_node) * dt, sub * dt});
    if (len_n(mine, k)) rcv.push_back(Seg{peer(p), shard + lo_n(mine, k) * dt, len_n(mine, k) * dt});
  }
  mesh_.exchange(tag(202), snd, rcv);
  snd.clear();
  rcv.clear();
  }
  // C: all-gather inside the node, straight into the result
  if (mine) memcpy(R + lo_l(li) * dt, shard, mine * dt);
  for (int j = 0; j < L; ++j) {
    if (j == li) continue;
    const int p = on_node[my_node][j];
    if (mine) snd.push_back(Seg{peer(p), shard, mine * dt});
    if (len_l(j)) rcv.push_back(Seg{peer(p), R + lo_l(j) * dt, len_l(j) * dt});
  }
  mesh_.exchange(tag(203), snd, rcv);
  return true;
}

// The two-level all-reduce cut into pieces: the shard (1/L of the message) is divided into up to 48 pieces, and every piece
// runs through the four steps on its own - node-local reduce-scatter, reduce-scatter among the nodes, all-gather among the
// nodes, node-local all-gather - driven by the arrivals inside ONE exchange: when the L - 1 local copies of a piece are there
// it is summed and its sub-slices leave for the other nodes, when their N - 1 copies of my sub-slice are there it is summed
// (scale applied) and passed around, and when a piece is complete it goes to the local members.  The shared-memory steps of
// later pieces run while earlier pieces are on the wire.  Same chains of additions as the one-piece form: same bits.
void NetBackend::hierarchical_allreduce_pipelined(CommRequest& r, const ProcessGroup& g, const std::function<uint64_t(int)>& tag,
                                                  const NodeMap& nm, size_t piece_bytes) {
  const CommDesc& d = r.desc;
  const size_t dt = dtype_size(d.dtype), n = d.count;
  const int N = nm.N, L = nm.L, my_node = nm.my_node, li = nm.li;
  char* S = (char*)r.send;
  char* R = (char*)r.recv;
  auto peer = [&](int p) { return g.members[p]; };
  constexpr int kMaxPieces = 16;                                      // (four steps x 16 tags fit the 8-bit step field with room to spare)
  const size_t per_l = ceil_div(n, (size_t)L);
  auto lo_l = [&](int j) { return std::min(n, (size_t)j * per_l); };
  auto len_l = [&](int j) { return std::min(n, lo_l(j) + per_l) - lo_l(j); };
  size_t ce = std::max(piece_bytes / dt, ceil_div(per_l, (size_t)kMaxPieces));
  ce = (ce + 63) & ~(size_t)63;
  const int K = (int)ceil_div(per_l, ce);
  auto plen = [&](int j, int c) {                                      // length of piece c inside slice j
    const size_t Lj = len_l(j), b = (size_t)c * ce;
    return b >= Lj ? (size_t)0 : std::min(ce, Lj - b);
  };
  const size_t per_n = ceil_div(ce, (size_t)N);                       // sub-slice of node k inside a piece
  auto slo = [&](size_t m, int k) { return std::min(m, (size_t)k * per_n); };
  auto slen = [&](size_t m, int k) { return std::min(m, slo(m, k) + per_n) - slo(m, k); };
  auto step = [&](int phase, int c) { return 2 + phase * kMaxPieces + c; };
  char* scratch = net_scratch(((size_t)L * per_l + per_l + (size_t)K * N * per_n) * dt);
  char* tmpA = scratch;                                               // [j]: local member j's copy of my slice
  char* shard = scratch + (size_t)L * per_l * dt;                     // my slice: node sum, then total
  char* tmpB = shard + per_l * dt;                                    // [c][k]: node k's copy of my sub-slice of piece c
  struct Meta {
    int phase, c;
  };
  std::vector<Seg> snd, rcv;
  std::vector<Meta> meta;
  std::vector<int> gotA(K, 0), gotB1(K, 0), gotB2(K, 0), needB1(K, 0), needB2(K, 0);
  for (int c = 0; c < K; ++c) {
    const size_t m = plen(li, c);
    // step 0 (inside the node): piece c of everybody's slice
    for (int j = 0; j < L; ++j) {
      if (j == li) continue;
      const int p = nm.on_node[my_node][j];
      if (plen(j, c)) snd.push_back(Seg{peer(p), S + (lo_l(j) + (size_t)c * ce) * dt, plen(j, c) * dt, tag(step(0, c))});
      if (m) {
        rcv.push_back(Seg{peer(p), tmpA + ((size_t)j * per_l + (size_t)c * ce) * dt, m * dt, tag(step(0, c))});
        meta.push_back(Meta{0, c});
      }
    }
    if (!m) continue;
    for (int k = 0; k < N; ++k) {
      if (k == my_node) continue;
      const int p = nm.on_node[k][li];
      if (slen(m, my_node)) {                                          // step 1: their copies of my sub-slice
        rcv.push_back(Seg{peer(p), tmpB + ((size_t)c * N + k) * per_n * dt, slen(m, my_node) * dt, tag(step(1, c))});
        meta.push_back(Meta{1, c});
        ++needB1[c];
      }
      if (slen(m, k)) {                                                // step 2: their finished sub-slices
        rcv.push_back(Seg{peer(p), shard + ((size_t)c * ce + slo(m, k)) * dt, slen(m, k) * dt, tag(step(2, c))});
        meta.push_back(Meta{2, c});
        ++needB2[c];
      }
    }
  }
  for (int j = 0; j < L; ++j) {                                        // step 3 (inside the node): finished pieces of the others
    if (j == li) continue;
    const int p = nm.on_node[my_node][j];
    for (int c = 0; c < K; ++c)
      if (plen(j, c)) {
        rcv.push_back(Seg{peer(p), R + (lo_l(j) + (size_t)c * ce) * dt, plen(j, c) * dt, tag(step(3, c))});
        meta.push_back(Meta{3, c});
      }
  }
  std::vector<const void*> srcs;
  std::vector<char> localDone(K, 0), subDone(K, 0), pieceDone(K, 0);
  // One rule for every arrival: a piece moves on as far as its inputs allow (an arrival can run ahead of the step before it -
  // the other nodes may be faster with piece c than this node's members - so every step checks the one before it too).
  auto advance = [&](int c) {
    const size_t m = plen(li, c), off = (size_t)c * ce;
    if (!localDone[c] && gotA[c] == L - 1) {                          // the node's copies of piece c of my slice are here
      srcs.assign(L, nullptr);
      for (int j = 0; j < L; ++j)
        srcs[j] = j == li ? (const void*)(S + (lo_l(li) + off) * dt) : (const void*)(tmpA + ((size_t)j * per_l + off) * dt);
      host_reduce(d.dtype, shard + off * dt, srcs, m, d.rop, 1.0f);
      for (int k = 0; k < N; ++k)
        if (k != my_node && slen(m, k))
          mesh_.add_send(Seg{peer(nm.on_node[k][li]), shard + (off + slo(m, k)) * dt, slen(m, k) * dt, tag(step(1, c))});
      localDone[c] = 1;
    }
    if (localDone[c] && !subDone[c] && gotB1[c] == needB1[c]) {       // every node's copy of my sub-slice of piece c is here
      const size_t sub = slen(m, my_node);
      if (sub) {
        srcs.assign(N, nullptr);
        for (int k = 0; k < N; ++k)
          srcs[k] = k == my_node ? (const void*)(shard + (off + slo(m, my_node)) * dt) : (const void*)(tmpB + ((size_t)c * N + k) * per_n * dt);
        host_reduce(d.dtype, shard + (off + slo(m, my_node)) * dt, srcs, sub, d.rop, d.scale);     // (element-wise: in place is fine)
        for (int k = 0; k < N; ++k)
          if (k != my_node) mesh_.add_send(Seg{peer(nm.on_node[k][li]), shard + (off + slo(m, my_node)) * dt, sub * dt, tag(step(2, c))});
      }
      subDone[c] = 1;
    }
    if (subDone[c] && !pieceDone[c] && gotB2[c] == needB2[c]) {       // piece c of my shard is final: to the result, to the node
      memcpy(R + (lo_l(li) + off) * dt, shard + off * dt, m * dt);
      for (int j = 0; j < L; ++j)
        if (j != li) mesh_.add_send(Seg{peer(nm.on_node[my_node][j]), shard + off * dt, m * dt, tag(step(3, c))});
      pieceDone[c] = 1;
    }
  };
  Mesh::RecvFn on_recv = [&](size_t i) {
    const Meta& mt = meta[i];
    if (mt.phase == 0) ++gotA[mt.c];
    else if (mt.phase == 1) ++gotB1[mt.c];
    else if (mt.phase == 2) ++gotB2[mt.c];
    else return;
    advance(mt.c);
  };
  mesh_.exchange(tag(0), snd, rcv, &on_recv);
}

// The same two levels for all-gather and reduce-scatter (the pair behind the distributed weight update): between nodes only
// the members with the same local index talk, (N - 1) n bytes per rank instead of (P - L) n; the node-local step moves the N
// blocks every member holds / needs through shared memory (packed into one message per local peer).
bool NetBackend::hierarchical_gather_scatter(CommRequest& r, const ProcessGroup& g, const std::function<uint64_t(int)>& tag) {
  const long hier_kb = ctx_->env.net_hier_kb;
  const CommDesc& d = r.desc;
  const size_t dt = dtype_size(d.dtype), n = d.count;      // n = elements of ONE block (per member)
  if (hier_kb < 0 || (size_t)g.size() * n * dt < (size_t)hier_kb << 10 || n == 0) return false;
  NodeMap nm;
  if (!node_map(g, nm)) return false;
  if (d.kind == OpKind::ALLGATHER) {
    hier_allgather(nm, g, tag, 204, (const char*)r.send, (char*)r.recv, n * dt);
    return true;
  }
  if (d.kind == OpKind::REDUCE_SCATTER) {
    hier_reduce_scatter(nm, g, tag, 100, d.dtype, d.rop, (const char*)r.send, (char*)r.recv, n, d.scale);
    return true;
  }
  return false;
}

// between nodes: my block goes to the members with my local index, theirs arrive at their places in R; inside the node:
// everybody passes on the N blocks of its column, packed [node 0 .. node N-1]          (tags step0, step0 + 1)
void NetBackend::hier_allgather(const NodeMap& nm, const ProcessGroup& g, const std::function<uint64_t(int)>& tag, int step0, const char* S,
                                char* R, size_t blk) {
  const int N = nm.N, L = nm.L, my_node = nm.my_node, li = nm.li;
  auto peer = [&](int p) { return g.members[p]; };
  std::vector<Seg> snd, rcv;
  if (R + (size_t)g.idx * blk != S) memmove(R + (size_t)g.idx * blk, S, blk);
  // From two pieces per block on, both levels run inside ONE exchange: a piece that arrives from another node is handed to
  // the local members straight from its place in R (and lands in its place in theirs: no packing on either side), so the
  // shared-memory step of early pieces runs while later ones are on the wire.  Tags: piece c of the wire step = step0 + c,
  // piece c of node k's block inside the node = step0 + k C + c (column peers and local peers are different ranks).
  // (a quarter of the reductions' piece: nothing is computed between the two levels, so finer pieces only shorten the ramp)
  const size_t piece_bytes = std::max<size_t>(4096, (size_t)std::max(0l, ctx_->env.net_chunk_kb) << 8);
  const int tags = 255 - step0;
  if (ctx_->env.net_hier_pipeline && N > 1 && L > 1 && N <= tags && blk >= 2 * piece_bytes) {
    const int C = (int)std::min<size_t>(ceil_div(blk, piece_bytes), (size_t)std::min(16, tags / N));
    const size_t ce = (ceil_div(blk, (size_t)C) + 63) & ~(size_t)63;
    auto plo = [&](int c) { return std::min(blk, (size_t)c * ce); };
    auto plen = [&](int c) { return std::min(blk, plo(c) + ce) - plo(c); };
    struct Meta {
      int k, c;      // k < 0: from a local member, nothing to pass on
    };
    std::vector<Meta> meta;
    for (int c = 0; c < C; ++c) {
      if (!plen(c)) continue;
      char* mine = R + (size_t)g.idx * blk + plo(c);
      for (int k = 0; k < N; ++k) {
        if (k == my_node) continue;
        const int p = nm.on_node[k][li];
        snd.push_back(Seg{peer(p), mine, plen(c), tag(step0 + c)});
        rcv.push_back(Seg{peer(p), R + (size_t)p * blk + plo(c), plen(c), tag(step0 + c)});
        meta.push_back(Meta{k, c});
      }
      for (int j = 0; j < L; ++j) {
        if (j == li) continue;
        const int p = nm.on_node[my_node][j];
        snd.push_back(Seg{peer(p), mine, plen(c), tag(step0 + my_node * C + c)});
        for (int k = 0; k < N; ++k) {
          rcv.push_back(Seg{peer(p), R + (size_t)nm.on_node[k][j] * blk + plo(c), plen(c), tag(step0 + k * C + c)});
          meta.push_back(Meta{-1, c});
        }
      }
    }
    Mesh::RecvFn on_recv = [&](size_t i) {
      const Meta& mt = meta[i];
      if (mt.k < 0) return;
      char* src = R + (size_t)nm.on_node[mt.k][li] * blk + plo(mt.c);
      for (int j = 0; j < L; ++j)
        if (j != li) mesh_.add_send(Seg{peer(nm.on_node[my_node][j]), src, plen(mt.c), tag(step0 + mt.k * C + mt.c)});
    };
    mesh_.exchange(tag(step0), snd, rcv, &on_recv);
    return;
  }
  for (int k = 0; k < N; ++k) {
    if (k == my_node) continue;
    const int p = nm.on_node[k][li];
    snd.push_back(Seg{peer(p), R + (size_t)g.idx * blk, blk});
    rcv.push_back(Seg{peer(p), R + (size_t)p * blk, blk});
  }
  mesh_.exchange(tag(step0), snd, rcv);
  snd.clear();
  rcv.clear();
  char* scratch = net_scratch((size_t)L * N * blk);
  char* mine = scratch + (size_t)li * N * blk;
  for (int k = 0; k < N; ++k) memcpy(mine + (size_t)k * blk, R + (size_t)nm.on_node[k][li] * blk, blk);
  for (int j = 0; j < L; ++j) {
    if (j == li) continue;
    const int p = nm.on_node[my_node][j];
    snd.push_back(Seg{peer(p), mine, (size_t)N * blk});
    rcv.push_back(Seg{peer(p), scratch + (size_t)j * N * blk, (size_t)N * blk});
  }
  mesh_.exchange(tag(step0 + 1), snd, rcv);
  for (int j = 0; j < L; ++j)
    if (j != li)
      for (int k = 0; k < N; ++k) memcpy(R + (size_t)nm.on_node[k][j] * blk, scratch + ((size_t)j * N + k) * blk, blk);
}

// inside the node: local member j collects, from every local member, the blocks meant for column j (packed by node) and adds
// them up - N partial sums per rank; between nodes: the partial sum for each member of my column goes to that member, which
// adds what arrives (scale applied there)                                            (tags step0, step0 + 1)
void NetBackend::hier_reduce_scatter(const NodeMap& nm, const ProcessGroup& g, const std::function<uint64_t(int)>& tag, int step0, DType dtype,
                                     RedOp rop, const char* S, char* R, size_t n, float scale, int step_end) {
  const int N = nm.N, L = nm.L, my_node = nm.my_node, li = nm.li;
  const size_t dt = dtype_size(dtype), blk = n * dt;
  auto peer = [&](int p) { return g.members[p]; };
  std::vector<Seg> snd, rcv;
  // From two pieces per block on, both levels run inside ONE exchange.  The blocks leave for the local members straight from
  // the input (no packing); when the L - 1 local copies of piece c of the block for node k's member of my column are here it
  // is summed and leaves for that member (k = my node: it stays), and when the N - 1 partial sums of piece c of MY block are
  // here the piece is finished (scale applied) - the node-local additions of later pieces run while earlier partial sums are
  // on the wire.  Same chains of additions as the two-exchange form below: same bits.
  // Tags [step0, step_end): piece c inside the node = step0 + k C + c, between the nodes = step0 + c (different peers).
  // (pieces of MLSL_NET_CHUNK_KB / 4 like the all-gather: 4 MiB on 2 x 4 ranks 4.3 -> 3.2 ms, larger sizes unchanged)
  const size_t piece_bytes = std::max<size_t>(4096, (size_t)std::max(0l, ctx_->env.net_chunk_kb) << 8);
  const int tags = std::min(step_end, 256) - step0;
  if (ctx_->env.net_hier_pipeline && N > 1 && L > 1 && N <= tags && blk >= 2 * piece_bytes) {
    const int C = (int)std::min<size_t>(ceil_div(blk, piece_bytes), (size_t)std::min(16, tags / N));
    const size_t ce = (ceil_div(n, (size_t)C) + 63) & ~(size_t)63;                  // elements per piece
    auto plo = [&](int c) { return std::min(n, (size_t)c * ce); };
    auto plen = [&](int c) { return std::min(n, plo(c) + ce) - plo(c); };
    const size_t P = (size_t)N * L;
    const bool alias = R < S + P * blk && S < R + blk;                             // in place: the input is still being sent
    char* scratch = net_scratch(((size_t)L * N + N + N + 1) * blk);
    char* in_pack = scratch;                                   // [j][k]: local member j's copy of the block for on_node[k][li]
    char* partial = in_pack + (size_t)L * N * blk;             // [k]: the node's sum of that block
    char* from_nodes = partial + (size_t)N * blk;              // [k]: node k's sum of MY block
    char* out = alias ? from_nodes + (size_t)N * blk : R;
    struct Meta {
      int wire, k, c;
    };
    std::vector<Meta> meta;
    std::vector<int> got_local((size_t)N * C, 0), got_wire(C, 0);
    std::vector<char> local_done((size_t)N * C, 0), done(C, 0);
    for (int c = 0; c < C; ++c) {
      if (!plen(c)) continue;
      for (int kk = 1; kk <= N; ++kk) {                        // the blocks that have to cross the wire first, my node's last
        const int k = (my_node + kk) % N;
        for (int j = 0; j < L; ++j) {
          if (j == li) continue;
          const int p = nm.on_node[my_node][j];
          snd.push_back(Seg{peer(p), const_cast<char*>(S) + (size_t)nm.on_node[k][j] * blk + plo(c) * dt, plen(c) * dt, tag(step0 + k * C + c)});
          rcv.push_back(Seg{peer(p), in_pack + ((size_t)j * N + k) * blk + plo(c) * dt, plen(c) * dt, tag(step0 + k * C + c)});
          meta.push_back(Meta{0, k, c});
        }
        if (k != my_node) {
          rcv.push_back(Seg{peer(nm.on_node[k][li]), from_nodes + (size_t)k * blk + plo(c) * dt, plen(c) * dt, tag(step0 + c)});
          meta.push_back(Meta{1, k, c});
        }
      }
    }
    std::vector<const void*> srcs;
    auto finish = [&](int c) {
      if (done[c] || !local_done[(size_t)my_node * C + c] || got_wire[c] != N - 1) return;
      srcs.assign(N, nullptr);
      for (int k = 0; k < N; ++k) srcs[k] = (k == my_node ? partial : from_nodes) + (size_t)k * blk + plo(c) * dt;
      host_reduce(dtype, out + plo(c) * dt, srcs, plen(c), rop, scale);
      done[c] = 1;
    };
    Mesh::RecvFn on_recv = [&](size_t i) {
      const Meta& mt = meta[i];
      const int c = mt.c, k = mt.k;
      if (mt.wire) {
        ++got_wire[c];
        finish(c);
        return;
      }
      if (++got_local[(size_t)k * C + c] != L - 1) return;
      srcs.assign(L, nullptr);
      for (int j = 0; j < L; ++j)
        srcs[j] = j == li ? (const void*)(S + (size_t)nm.on_node[k][li] * blk + plo(c) * dt)
                          : (const void*)(in_pack + ((size_t)j * N + k) * blk + plo(c) * dt);
      host_reduce(dtype, partial + (size_t)k * blk + plo(c) * dt, srcs, plen(c), rop, 1.0f);
      local_done[(size_t)k * C + c] = 1;
      if (k != my_node) mesh_.add_send(Seg{peer(nm.on_node[k][li]), partial + (size_t)k * blk + plo(c) * dt, plen(c) * dt, tag(step0 + c)});
      else finish(c);
    };
    mesh_.exchange(tag(step0), snd, rcv, &on_recv);
    if (alias) memcpy(R, out, blk);
    return;
  }
  char* scratch = net_scratch(((size_t)L * N + (size_t)L * N + N + N) * blk);
  char* out_pack = scratch;                                   // [j][k]: what I send to local member j
  char* in_pack = scratch + (size_t)L * N * blk;              // [j][k]: what local member j sent me
  char* partial = in_pack + (size_t)L * N * blk;              // [k]: node-local sum of the block for member on_node[k][li]
  char* from_nodes = partial + (size_t)N * blk;               // [k]: partial sums for ME from the other nodes
  for (int j = 0; j < L; ++j)
    for (int k = 0; k < N; ++k) memcpy(out_pack + ((size_t)j * N + k) * blk, S + (size_t)nm.on_node[k][j] * blk, blk);
  for (int j = 0; j < L; ++j) {
    if (j == li) continue;
    const int p = nm.on_node[my_node][j];
    snd.push_back(Seg{peer(p), out_pack + (size_t)j * N * blk, (size_t)N * blk});
    rcv.push_back(Seg{peer(p), in_pack + (size_t)j * N * blk, (size_t)N * blk});
  }
  mesh_.exchange(tag(step0), snd, rcv);
  snd.clear();
  rcv.clear();
  std::vector<const void*> srcs(L);
  for (int j = 0; j < L; ++j) srcs[j] = j == li ? (const void*)(out_pack + (size_t)li * N * blk) : (const void*)(in_pack + (size_t)j * N * blk);
  host_reduce(dtype, partial, srcs, (size_t)N * n, rop, 1.0f);
  for (int k = 0; k < N; ++k) {
    if (k == my_node) continue;
    const int p = nm.on_node[k][li];
    snd.push_back(Seg{peer(p), partial + (size_t)k * blk, blk});
    rcv.push_back(Seg{peer(p), from_nodes + (size_t)k * blk, blk});
  }
  mesh_.exchange(tag(step0 + 1), snd, rcv);
  std::vector<const void*> parts(N);
  for (int k = 0; k < N; ++k) parts[k] = k == my_node ? (const void*)(partial + (size_t)my_node * blk) : (const void*)(from_nodes + (size_t)k * blk);
  host_reduce(dtype, R, parts, n, rop, scale);
}

// Broadcast in two levels: the root's column (the members with its local index, one per node) gets the buffer over the wire -
// scatter + all-gather among them from three nodes on, so the root's link carries the message once - and every column member
// hands it to the other members of its node through shared memory.  One copy of the message enters each node.
bool NetBackend::hierarchical_bcast(CommRequest& r, const ProcessGroup& g, const std::function<uint64_t(int)>& tag) {
  const long hier_kb = ctx_->env.net_hier_kb;
  const CommDesc& d = r.desc;
  const size_t bytes = d.count * dtype_size(d.dtype);
  if (hier_kb < 0 || bytes < (size_t)hier_kb << 10 || bytes < 64) return false;
  NodeMap nm;
  if (!node_map(g, nm)) return false;
  const int N = nm.N, L = nm.L, my_node = nm.my_node, li = nm.li, root = (int)d.root;
  int root_node = 0, root_li = 0;
  for (int k = 0; k < N; ++k)
    for (int j = 0; j < L; ++j)
      if (nm.on_node[k][j] == root) {
        root_node = k;
        root_li = j;
      }
  char* R = (char*)r.recv;
  auto peer = [&](int p) { return g.members[p]; };
  std::vector<Seg> snd, rcv;
  // Two nodes, two pieces or more: one exchange on every rank - the root feeds the other node's column member and its own
  // node piece by piece, the column member passes every piece on to its node when it arrives (shared memory while the next
  // piece is on the wire), the others receive the pieces in place.                  (tags: wire 210 + c, inside a node 230 + c)
  const size_t piece_bytes = std::max<size_t>(4096, (size_t)std::max(0l, ctx_->env.net_chunk_kb) << 8);
  if (N == 2 && ctx_->env.net_hier_pipeline && bytes >= 2 * piece_bytes) {
    const int C = (int)std::min<size_t>(ceil_div(bytes, piece_bytes), 16);
    const size_t ce = (ceil_div(bytes, (size_t)C) + 63) & ~(size_t)63;
    auto plo = [&](int c) { return std::min(bytes, (size_t)c * ce); };
    auto plen = [&](int c) { return std::min(bytes, plo(c) + ce) - plo(c); };
    auto to_node = [&](int c, bool queued) {
      for (int j = 0; j < L; ++j) {
        if (j == li) continue;
        const Seg sg{peer(nm.on_node[my_node][j]), R + plo(c), plen(c), tag(230 + c)};
        if (queued) mesh_.add_send(sg);
        else snd.push_back(sg);
      }
    };
    for (int c = 0; c < C; ++c) {
      if (!plen(c)) continue;
      if (li != root_li) {
        rcv.push_back(Seg{peer(nm.on_node[my_node][root_li]), R + plo(c), plen(c), tag(230 + c)});
      } else if (my_node == root_node) {
        snd.push_back(Seg{peer(nm.on_node[1 - my_node][li]), R + plo(c), plen(c), tag(210 + c)});
        to_node(c, false);
      } else {
        rcv.push_back(Seg{peer(root), R + plo(c), plen(c), tag(210 + c)});
      }
    }
    std::vector<int> piece_of;                             // receive index -> piece (empty pieces are not posted)
    for (int c = 0; c < C; ++c)
      if (plen(c)) piece_of.push_back(c);
    Mesh::RecvFn on_recv = [&](size_t i) {
      if (li == root_li && my_node != root_node) to_node(piece_of[i], true);
    };
    mesh_.exchange(tag(206), snd, rcv, &on_recv);
    return true;
  }
  // Three nodes and more: the same scatter + all-gather among the root's column, slice by slice inside one exchange - a
  // column member passes its own slice on to the other column members the moment it has it, and every slice (its own, the
  // root's, the others') goes to its node when it arrives.                    (tags: wire 210 + slice, inside a node 230 + slice)
  if (N >= 3 && N <= 16 && ctx_->env.net_hier_pipeline && bytes >= 2 * piece_bytes) {
    const size_t per = (ceil_div(bytes, (size_t)N) + 63) & ~(size_t)63;
    auto lo = [&](int k) { return std::min(bytes, (size_t)k * per); };
    auto len = [&](int k) { return std::min(bytes, lo(k) + per) - lo(k); };
    const bool column = li == root_li, is_root = column && my_node == root_node;
    std::vector<int> slice_of;                             // receive index -> slice
    auto to_node = [&](int sl, bool queued) {
      for (int j = 0; j < L; ++j) {
        if (j == li) continue;
        const Seg sg{peer(nm.on_node[my_node][j]), R + lo(sl), len(sl), tag(230 + sl)};
        if (queued) mesh_.add_send(sg);
        else snd.push_back(sg);
      }
    };
    auto expect = [&](int from, int sl, int step) {
      if (!len(sl)) return;
      rcv.push_back(Seg{peer(from), R + lo(sl), len(sl), tag(step + sl)});
      slice_of.push_back(sl);
    };
    if (!column) {
      for (int sl = 0; sl < N; ++sl) expect(nm.on_node[my_node][root_li], sl, 230);
    } else if (is_root) {
      for (int k = 0; k < N; ++k)
        if (k != my_node && len(k)) snd.push_back(Seg{peer(nm.on_node[k][li]), R + lo(k), len(k), tag(210 + k)});
      for (int k = 0; k < N; ++k)
        if (k != my_node && len(my_node)) snd.push_back(Seg{peer(nm.on_node[k][li]), R + lo(my_node), len(my_node), tag(210 + my_node)});
      for (int sl = 0; sl < N; ++sl)
        if (len(sl)) to_node(sl, false);
    } else {
      expect(root, my_node, 210);
      expect(root, root_node, 210);
      for (int k = 0; k < N; ++k)
        if (k != my_node && k != root_node) expect(nm.on_node[k][li], k, 210);
    }
    Mesh::RecvFn on_recv = [&](size_t i) {
      if (!column || is_root) return;
      const int sl = slice_of[i];
      if (sl == my_node)
        for (int k = 0; k < N; ++k)
          if (k != my_node && k != root_node) mesh_.add_send(Seg{peer(nm.on_node[k][li]), R + lo(sl), len(sl), tag(210 + sl)});
      to_node(sl, true);
    };
    mesh_.exchange(tag(206), snd, rcv, &on_recv);
    return true;
  }
  if (li == root_li) {                                   // the column of the root: between nodes
    if (N == 2) {
      if (my_node == root_node) snd.push_back(Seg{peer(nm.on_node[1 - my_node][li]), R, bytes});
      else rcv.push_back(Seg{peer(root), R, bytes});
      mesh_.exchange(tag(206), snd, rcv);
    } else {
      const size_t per = ceil_div(bytes, (size_t)N);
      auto lo = [&](int k) { return std::min(bytes, (size_t)k * per); };
      auto len = [&](int k) { return std::min(bytes, lo(k) + per) - lo(k); };
      if (my_node == root_node) {
        for (int k = 0; k < N; ++k)
          if (k != my_node && len(k)) snd.push_back(Seg{peer(nm.on_node[k][li]), R + lo(k), len(k)});
      } else if (len(my_node)) {
        rcv.push_back(Seg{peer(root), R + lo(my_node), len(my_node)});
      }
      mesh_.exchange(tag(206), snd, rcv);
      snd.clear();
      rcv.clear();
      for (int k = 0; k < N; ++k) {
        if (k == my_node) continue;
        const int p = nm.on_node[k][li];
        if (k != root_node && len(my_node)) snd.push_back(Seg{peer(p), R + lo(my_node), len(my_node)});
        if (my_node != root_node && len(k)) rcv.push_back(Seg{peer(p), R + lo(k), len(k)});
      }
      mesh_.exchange(tag(207), snd, rcv);
    }
    snd.clear();
    rcv.clear();
    for (int j = 0; j < L; ++j)                           // inside the node
      if (j != li) snd.push_back(Seg{peer(nm.on_node[my_node][j]), R, bytes});
    mesh_.exchange(tag(208), snd, rcv);
  } else {
    rcv.push_back(Seg{peer(nm.on_node[my_node][root_li]), R, bytes});
    mesh_.exchange(tag(208), snd, rcv);
  }
  return true;
}

// Quantised all-reduce (CT_QUANTIZATION) between nodes - where the reference's gradient compression matters most: the same
// block-scaled FP8 format and the same three steps as the host and device editions (quant.hpp: 128-element blocks, one
// fp32 scale each, error feedback into a per-request residual), with the two data movements as mesh exchanges.  Every rank
// sends (P-1)/P * n * 1.03 bytes per exchange instead of (P-1)/P * 4n, and all ranks end up with bitwise identical values
// because every block is dequantised from the one requantised copy its owner produced.
void NetBackend::quantized_allreduce(CommRequest& r, const ProcessGroup& g) {
  const CommDesc& d = r.desc;
  if (!r.backend_state) r.backend_state = new NetReqState();
  NetReqState* st = (NetReqState*)r.backend_state;
  auto tag = [&](int step) {
    return ((uint64_t)(uint8_t)g.row << 56) | ((uint64_t)(r.lane & 0xff) << 48) | ((r.group_seq & 0xffffffffffull) << 8) |
           (uint64_t)(step & 0xff);
  };
  quantized_exchange((const float*)r.send, (float*)r.recv, d.count, g.members, g.idx, st->residual, d.scale, tag, 0);
}

// x (+ the residual of earlier rounds) -> fp8 blocks -> every member's range summed by its owner and requantised once ->
// y = dequantised sum * scale, among `ranks` (global ranks, this one at position `me`); two exchanges under tag(step0), tag(step0 + 1)
void NetBackend::quantized_exchange(const float* x, float* y, size_t n, const std::vector<int>& ranks, int me, std::vector<float>& residual,
                                    float scale, const std::function<uint64_t(int)>& tag, int step0) {
  const int P = (int)ranks.size();
  if (residual.size() != n) residual.assign(n, 0.f);
  const size_t nblk = ceil_div(n, (size_t)kQuantBlock), blk_per = ceil_div(nblk, (size_t)P);
  auto blo = [&](int p) { return std::min(nblk, (size_t)p * blk_per); };
  auto bcnt = [&](int p) { return std::min(nblk, blo(p) + blk_per) - blo(p); };
  const size_t chunk = blk_per * kQuantBlockBytes;                       // room for one owner's blocks: [fp8 bytes | scales]
  auto qof = [&](char* base, int p) { return (uint8_t*)(base + (size_t)p * chunk); };
  auto sof = [&](char* base, int p) { return (float*)(base + (size_t)p * chunk + bcnt(p) * kQuantBlock); };
  // scratch: outgoing chunks by owner | incoming chunks of my range by source | requantised ranges by owner
  char* out = net_scratch_q(3 * (size_t)P * chunk);
  char* in = out + (size_t)P * chunk;
  char* red = in + (size_t)P * chunk;
  // step 1: x + residual -> fp8 blocks laid out per owner, residual update
  for (int p = 0; p < P; ++p)
    for (size_t k = 0; k < bcnt(p); ++k) {
      const size_t b = blo(p) + k, lo = b * kQuantBlock, hi = std::min(n, lo + kQuantBlock);
      float v[kQuantBlock];
      for (size_t i = lo; i < hi; ++i) v[i - lo] = x[i] + residual[i];
      for (size_t i = hi - lo; i < (size_t)kQuantBlock; ++i) v[i] = 0.f;
      uint8_t* q = qof(out, p) + k * kQuantBlock;
      const float sc = quant_block(v, q);
      sof(out, p)[k] = sc;
      for (size_t i = lo; i < hi; ++i) residual[i] = v[i - lo] - e4m3_to_f32(q[i - lo]) * sc;
    }
  std::vector<Seg> snd, rcv;
  const size_t mine_bytes = bcnt(me) * kQuantBlockBytes;
  for (int p = 0; p < P; ++p)
    if (p != me) {
      if (bcnt(p)) snd.push_back(Seg{ranks[p], out + (size_t)p * chunk, bcnt(p) * kQuantBlockBytes});
      if (mine_bytes) rcv.push_back(Seg{ranks[p], in + (size_t)p * chunk, mine_bytes});
    }
  mesh_.exchange(tag(step0), snd, rcv);
  snd.clear();
  rcv.clear();
  // step 2: my blocks - dequantise and add in member order, requantise once
  for (size_t k = 0; k < bcnt(me); ++k) {
    float acc[kQuantBlock];
    for (int i = 0; i < kQuantBlock; ++i) acc[i] = 0.f;
    for (int p = 0; p < P; ++p) {
      const char* base = p == me ? out + (size_t)me * chunk : in + (size_t)p * chunk;
      const uint8_t* q = (const uint8_t*)base + k * kQuantBlock;
      const float sc = ((const float*)(base + bcnt(me) * kQuantBlock))[k];
      for (int i = 0; i < kQuantBlock; ++i) acc[i] += e4m3_to_f32(q[i]) * sc;
    }
    sof(red, me)[k] = quant_block(acc, qof(red, me) + k * kQuantBlock);
  }
  for (int p = 0; p < P; ++p)
    if (p != me) {
      if (mine_bytes) snd.push_back(Seg{ranks[p], red + (size_t)me * chunk, mine_bytes});
      if (bcnt(p)) rcv.push_back(Seg{ranks[p], red + (size_t)p * chunk, bcnt(p) * kQuantBlockBytes});
    }
  mesh_.exchange(tag(step0 + 1), snd, rcv);
  // step 3: every range from its owner's requantised copy, output scale fused
  for (int p = 0; p < P; ++p)
    for (size_t k = 0; k < bcnt(p); ++k) {
      const size_t b = blo(p) + k, lo = b * kQuantBlock, hi = std::min(n, lo + kQuantBlock);
      const uint8_t* q = qof(red, p) + k * kQuantBlock;
      const float sc = sof(red, p)[k];
      for (size_t i = lo; i < hi; ++i) y[i] = e4m3_to_f32(q[i - lo]) * sc * scale;   // same order as the host edition
    }
}

}  // namespace

std::unique_ptr<Backend> make_net_backend(RankContext* ctx) { return std::unique_ptr<Backend>(new NetBackend(ctx)); }

}  // namespace mlslb
