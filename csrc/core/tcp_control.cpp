#include "tcp_control.hpp"

#include <algorithm>

#include <arpa/inet.h>
#include <ifaddrs.h>
#include <net/if.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <tuple>

#include "common.hpp"
#include "log.hpp"

namespace mlslb {
constexpr uint64_t kMaxCtlPayload = (uint64_t)1 << 20;   // gather payloads are kBootSlotBytes-sized; 1 MiB is generous


// ---- socket helpers ------------------------------------------------------------------------------------------------
static bool resolve(const std::string& addr, int port, sockaddr_in* out) {
  memset(out, 0, sizeof(*out));
  out->sin_family = AF_INET;
  out->sin_port = htons((uint16_t)port);
  if (addr.empty() || addr == "*") {
    out->sin_addr.s_addr = htonl(INADDR_ANY);
    return true;
  }
  if (inet_pton(AF_INET, addr.c_str(), &out->sin_addr) == 1) return true;
  addrinfo hints, *res = nullptr;
  memset(&hints, 0, sizeof(hints));
  hints.ai_family = AF_INET;
  hints.ai_socktype = SOCK_STREAM;
  if (getaddrinfo(addr.c_str(), nullptr, &hints, &res) != 0 || !res) return false;
  out->sin_addr = ((sockaddr_in*)res->ai_addr)->sin_addr;
  freeaddrinfo(res);
  return true;
}

void tcp_tune(int fd, long sockbuf_kb) {
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
  // 0: the kernel sizes the buffers itself (tcp_rmem / tcp_wmem auto-tuning) - a fixed size switches that off and is capped by
  // net.core.[rw]mem_max; measured on loop-back, the auto-tuned sockets move 64 MiB collectives 7 - 10 % faster than 4 MiB ones
  int buf = (int)std::min<long>(sockbuf_kb, 1 << 20) << 10;
  if (buf > 0) {
    setsockopt(fd, SOL_SOCKET, SO_SNDBUF, &buf, sizeof(buf));
    setsockopt(fd, SOL_SOCKET, SO_RCVBUF, &buf, sizeof(buf));
  }
}

int tcp_listen(const std::string& addr, int port, int backlog, int* bound_port) {
  int fd = socket(AF_INET, SOCK_STREAM | SOCK_CLOEXEC, 0);
  MLSLB_ASSERT(fd >= 0, "socket(): %s", strerror(errno));
  int one = 1;
  setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  sockaddr_in sa;
  MLSLB_ASSERT(resolve(addr, port, &sa), "cannot resolve listen address '%s'", addr.c_str());
  MLSLB_ASSERT(bind(fd, (sockaddr*)&sa, sizeof(sa)) == 0, "bind(%s:%d): %s", addr.c_str(), port, strerror(errno));
  MLSLB_ASSERT(listen(fd, backlog) == 0, "listen(): %s", strerror(errno));
  if (bound_port) {
    socklen_t len = sizeof(sa);
    getsockname(fd, (sockaddr*)&sa, &len);
    *bound_port = ntohs(sa.sin_port);
  }
  return fd;
}

int tcp_connect_retry(const std::string& addr, int port, int timeout_sec) {
  sockaddr_in sa;
  MLSLB_ASSERT(resolve(addr, port, &sa), "cannot resolve '%s'", addr.c_str());
  const uint64_t t0 = now_ns();
  for (;;) {
    int fd = socket(AF_INET, SOCK_STREAM | SOCK_CLOEXEC, 0);
    MLSLB_ASSERT(fd >= 0, "socket(): %s", strerror(errno));
    if (connect(fd, (sockaddr*)&sa, sizeof(sa)) == 0) {
      tcp_tune(fd);
      return fd;
    }
    int e = errno;
    close(fd);
    MLSLB_ASSERT(now_ns() - t0 < (uint64_t)timeout_sec * 1000000000ull, "connect(%s:%d) kept failing for %d s: %s",
                 addr.c_str(), port, timeout_sec, strerror(e));
    usleep(20000);
  }
}

void tcp_send_all(int fd, const void* buf, size_t bytes) {
  const char* p = (const char*)buf;
  while (bytes) {
    ssize_t n = send(fd, p, bytes, MSG_NOSIGNAL);
    if (n < 0 && errno == EINTR) continue;
    // an exception in every assert mode: callers (reader threads, teardown paths) decide what a lost connection means
    if (n <= 0) throw Error(std::string("send(): ") + (n == 0 ? "connection closed" : strerror(errno)));
    p += n;
    bytes -= (size_t)n;
  }
}

void tcp_recv_all(int fd, void* buf, size_t bytes) {
  char* p = (char*)buf;
  while (bytes) {
    ssize_t n = recv(fd, p, bytes, 0);
    if (n < 0 && errno == EINTR) continue;
    if (n <= 0) throw Error(std::string("recv(): ") + (n == 0 ? "connection closed by the peer" : strerror(errno)));
    p += n;
    bytes -= (size_t)n;
  }
}

std::string tcp_local_address_towards(const std::string& addr, int port) {
  // a connected UDP socket never sends anything but tells which local address routes to the peer
  sockaddr_in sa;
  if (!resolve(addr, port, &sa)) return "127.0.0.1";
  int fd = socket(AF_INET, SOCK_DGRAM | SOCK_CLOEXEC, 0);
  if (fd < 0) return "127.0.0.1";
  std::string out = "127.0.0.1";
  if (connect(fd, (sockaddr*)&sa, sizeof(sa)) == 0) {
    sockaddr_in me;
    socklen_t len = sizeof(me);
    char buf[64];
    if (getsockname(fd, (sockaddr*)&me, &len) == 0 && inet_ntop(AF_INET, &me.sin_addr, buf, sizeof(buf))) out = buf;
  }
  close(fd);
  return out;
}

std::string tcp_address_of_interface(const std::string& prefix, int idx) {
  ifaddrs* list = nullptr;
  if (getifaddrs(&list) != 0) return "";
  std::string out;
  int seen = 0;
  for (ifaddrs* ifa = list; ifa; ifa = ifa->ifa_next) {
    if (!ifa->ifa_addr || ifa->ifa_addr->sa_family != AF_INET || (ifa->ifa_flags & IFF_LOOPBACK) || !(ifa->ifa_flags & IFF_UP)) continue;
    const bool hit = !prefix.empty() ? strncmp(ifa->ifa_name, prefix.c_str(), prefix.size()) == 0 : seen == idx;
    ++seen;
    if (!hit) continue;
    char buf[64];
    if (inet_ntop(AF_INET, &((sockaddr_in*)ifa->ifa_addr)->sin_addr, buf, sizeof(buf))) out = buf;
    break;
  }
  freeifaddrs(list);
  return out;
}

uint64_t tcp_job_token() {
  static const uint64_t token = [] {
    const char* v = getenv("MLSL_JOB_TOKEN");
    uint64_t h = 1469598103934665603ull;                       // FNV-1a
    for (const char* c = (v && *v) ? v : "mlsl-b200"; *c; ++c) h = (h ^ (uint8_t)*c) * 1099511628211ull;
    return h ? h : 1;
  }();
  return token;
}

std::string tcp_resolve_to_ip(const std::string& host) {
  sockaddr_in sa;
  char buf[64];
  if (!resolve(host, 0, &sa) || !inet_ntop(AF_INET, &sa.sin_addr, buf, sizeof(buf))) return "";
  return buf;
}

// ---- wire format -----------------------------------------------------------------------------------------------------
namespace {
enum : uint32_t { MSG_HELLO = 1, MSG_GATHER = 2, MSG_POISON = 3, MSG_REPLY = 4, MSG_NOTIFY = 5, MSG_BYE = 6 };
struct CtlMsg {
  uint32_t type, rank;
  uint64_t req, key, seq;
  uint32_t nmembers, idx, first_member, bytes;   // bytes: payload that follows
};
}  // namespace

struct TcpControl::Server {
  int listen_fd = -1;
  std::thread acceptor;
  std::vector<std::thread> clients;
  std::mutex mu;
  std::vector<int> fds;                 // by rank (-1 until that rank says hello)
  std::vector<std::unique_ptr<std::mutex>> wmu;
  std::vector<bool> said_bye;
  std::atomic<int> live_clients{0};
  struct Op {
    uint32_t nmembers = 0, bytes = 0, arrived = 0;
    std::vector<char> data;             // nmembers * bytes
    std::vector<std::pair<int, uint64_t>> waiters;   // (rank, request id) by arrival
  };
  std::map<std::tuple<uint64_t, uint64_t, uint32_t>, Op> ops;
  uint64_t poison = 0;
  void send_to(int rank, const CtlMsg& h, const void* payload) {
    if (rank < 0 || rank >= (int)fds.size() || fds[rank] < 0) return;
    std::lock_guard<std::mutex> g(*wmu[rank]);
    try {
      tcp_send_all(fds[rank], &h, sizeof(h));
      if (h.bytes) tcp_send_all(fds[rank], payload, h.bytes);
    } catch (const std::exception&) {   // that rank is gone; its own reader notices
    }
  }
  void broadcast_poison(uint64_t code) {
    CtlMsg h;
    memset(&h, 0, sizeof(h));
    h.type = MSG_NOTIFY;
    h.key = code;
    for (int r = 0; r < (int)fds.size(); ++r) send_to(r, h, nullptr);
  }
};

TcpControl::TcpControl(const std::string& master_addr, int master_port, int rank, int world) : rank_(rank), world_(world) {
  if (rank == 0) {
    srv_.reset(new Server());
    srv_->fds.assign(world, -1);
    srv_->said_bye.assign(world, false);
    for (int i = 0; i < world; ++i) srv_->wmu.emplace_back(new std::mutex());
    srv_->listen_fd = tcp_listen("*", master_port, world + 8, nullptr);
    srv_->acceptor = std::thread([this] { server_accept_loop(); });
  }
  sock_ = tcp_connect_retry(rank == 0 ? std::string("127.0.0.1") : master_addr, master_port, 120);
  CtlMsg h;
  memset(&h, 0, sizeof(h));
  h.type = MSG_HELLO;
  h.rank = (uint32_t)rank;
  h.nmembers = (uint32_t)world;
  h.key = tcp_job_token();
  tcp_send_all(sock_, &h, sizeof(h));
  rx_ = std::thread([this] { rx_loop(); });
}

TcpControl::~TcpControl() {
  stopping_.store(true);
  if (sock_ >= 0) shutdown(sock_, SHUT_RDWR);
  if (rx_.joinable()) rx_.join();
  if (sock_ >= 0) close(sock_);
  if (srv_) {
    // Peers say goodbye and close after their last collective; replies to them may still be in flight from one of the
    // reader threads, so give them a moment before the sockets are torn down (a peer that died is not waited for long).
    const uint64_t t0 = now_ns();
    while (srv_->live_clients.load() > 0 && now_ns() - t0 < 10ull * 1000000000ull) usleep(1000);
    shutdown(srv_->listen_fd, SHUT_RDWR);
    close(srv_->listen_fd);
    if (srv_->acceptor.joinable()) srv_->acceptor.join();
    {
      std::lock_guard<std::mutex> g(srv_->mu);
      for (int fd : srv_->fds)
        if (fd >= 0) shutdown(fd, SHUT_RDWR);
    }
    for (auto& t : srv_->clients)
      if (t.joinable()) t.join();
    for (int fd : srv_->fds)
      if (fd >= 0) close(fd);
  }
}

void TcpControl::goodbye() {
  CtlMsg h;
  memset(&h, 0, sizeof(h));
  h.type = MSG_BYE;
  h.rank = (uint32_t)rank_;
  std::lock_guard<std::mutex> g(tx_mu_);
  try {
    tcp_send_all(sock_, &h, sizeof(h));
  } catch (const std::exception&) {
  }
}

void TcpControl::poison(int code) {
  uint64_t expect = 0;
  poison_.compare_exchange_strong(expect, (uint64_t)code + 1);
  CtlMsg h;
  memset(&h, 0, sizeof(h));
  h.type = MSG_POISON;
  h.rank = (uint32_t)rank_;
  h.key = (uint64_t)code + 1;
  std::lock_guard<std::mutex> g(tx_mu_);
  try {
    tcp_send_all(sock_, &h, sizeof(h));
  } catch (const std::exception&) {
  }
  cv_.notify_all();
}

void TcpControl::gather(uint64_t key, uint64_t seq, const std::vector<int>& members, int idx, const void* in, void* out,
                        size_t bytes) {
  CtlMsg h;
  memset(&h, 0, sizeof(h));
  h.type = MSG_GATHER;
  h.rank = (uint32_t)rank_;
  h.key = key;
  h.seq = seq;
  h.nmembers = (uint32_t)members.size();
  h.idx = (uint32_t)idx;
  h.first_member = (uint32_t)members[0];
  h.bytes = (uint32_t)bytes;
  uint64_t req;
  {
    std::lock_guard<std::mutex> g(mu_);
    req = next_req_++;
    pending_[req];
  }
  h.req = req;
  try {
    std::lock_guard<std::mutex> g(tx_mu_);
    tcp_send_all(sock_, &h, sizeof(h));
    if (bytes) tcp_send_all(sock_, in, bytes);
  } catch (const Error& e) {
    MLSLB_ASSERT(false, "lost the connection to the control server (rank 0): %s", e.what());
  }
  std::unique_lock<std::mutex> lk(mu_);
  const uint64_t t0 = now_ns();
  while (!pending_[req].done) {
    cv_.wait_for(lk, std::chrono::milliseconds(50));
    if (pending_[req].done) break;
    if (poison_.load() != 0) {
      pending_.erase(req);
      MLSLB_ASSERT(false, "job poisoned by rank %d while waiting in a control collective", (int)poison_.load() - 1);
    }
    MLSLB_ASSERT(now_ns() - t0 < 300ull * 1000000000ull, "control collective (key %llu seq %llu) never completed",
                 (unsigned long long)key, (unsigned long long)seq);
  }
  Pending p = std::move(pending_[req]);
  pending_.erase(req);
  lk.unlock();
  MLSLB_ASSERT(p.payload.size() == bytes * members.size(), "control collective returned %zu bytes, expected %zu",
               p.payload.size(), bytes * members.size());
  if (bytes && out) memcpy(out, p.payload.data(), p.payload.size());
}

void TcpControl::rx_loop() {
  for (;;) {
    CtlMsg h;
    try {
      tcp_recv_all(sock_, &h, sizeof(h));
      std::vector<char> payload(h.bytes);
      if (h.bytes) tcp_recv_all(sock_, payload.data(), h.bytes);
      if (h.type == MSG_REPLY) {
        std::lock_guard<std::mutex> g(mu_);
        auto it = pending_.find(h.req);
        if (it != pending_.end()) {
          it->second.payload = std::move(payload);
          it->second.done = true;
        }
        cv_.notify_all();
      } else if (h.type == MSG_NOTIFY) {
        uint64_t expect = 0;
        poison_.compare_exchange_strong(expect, h.key);
        cv_.notify_all();
      }
    } catch (const std::exception&) {
      if (!stopping_.load()) {   // the control server vanished: nobody can make progress any more
        uint64_t expect = 0;
        poison_.compare_exchange_strong(expect, 1);
        cv_.notify_all();
      }
      return;
    }
  }
}

// ---- server side (rank 0) ------------------------------------------------------------------------------------------------
void TcpControl::server_accept_loop() {
  Server* s = srv_.get();
  for (;;) {
    int fd = accept4(s->listen_fd, nullptr, nullptr, SOCK_CLOEXEC);
    if (fd < 0) {
      if (errno == EINTR) continue;
      return;   // listener closed
    }
    tcp_tune(fd);
    CtlMsg h;
    try {
      tcp_recv_all(fd, &h, sizeof(h));
    } catch (const std::exception&) {
      close(fd);
      continue;
    }
    // (compared as unsigned: a rank >= 2^31 must not turn into a negative index)
    if (h.type != MSG_HELLO || (uint64_t)h.rank >= (uint64_t)world_ || (uint64_t)h.nmembers != (uint64_t)world_ ||
        h.key != tcp_job_token()) {
      if (h.type == MSG_HELLO && h.key != tcp_job_token())
        MLSLB_LOG(LOG_ERROR, "control server: turned away a connection that claims rank %u with another job token (MLSL_JOB_TOKEN)", h.rank);
      close(fd);
      continue;
    }
    std::lock_guard<std::mutex> g(s->mu);
    s->fds[h.rank] = fd;
    s->live_clients.fetch_add(1);
    s->clients.emplace_back([this, fd, h] {
      server_client_loop(fd, (int)h.rank);
      srv_->live_clients.fetch_sub(1);
    });
  }
}

void TcpControl::server_client_loop(int fd, int peer_rank) {
  Server* s = srv_.get();
  for (;;) {
    CtlMsg h;
    std::vector<char> payload;
    try {
      tcp_recv_all(fd, &h, sizeof(h));
      if ((uint64_t)h.bytes > kMaxCtlPayload) throw std::runtime_error("control message too large");   // never trust a length from the wire
      payload.resize(h.bytes);
      if (h.bytes) tcp_recv_all(fd, payload.data(), h.bytes);
    } catch (const std::exception&) {
      bool bye;
      {
        std::lock_guard<std::mutex> g(s->mu);
        bye = s->said_bye[peer_rank];
      }
      if (!bye && !stopping_.load()) {   // a rank died without saying goodbye: fail the whole job fast
        uint64_t code = (uint64_t)peer_rank + 1;
        {
          std::lock_guard<std::mutex> g(s->mu);
          if (s->poison == 0) s->poison = code;
          code = s->poison;
        }
        s->broadcast_poison(code);
      }
      return;
    }
    if (h.type == MSG_BYE) {
      std::lock_guard<std::mutex> g(s->mu);
      s->said_bye[peer_rank] = true;
      continue;
    }
    if (h.type == MSG_POISON) {
      uint64_t code;
      {
        std::lock_guard<std::mutex> g(s->mu);
        if (s->poison == 0) s->poison = h.key;
        code = s->poison;
      }
      s->broadcast_poison(code);
      continue;
    }
    if (h.type != MSG_GATHER) continue;
    std::vector<std::pair<int, uint64_t>> waiters;
    std::vector<char> result;
    {
      std::lock_guard<std::mutex> g(s->mu);
      auto key = std::make_tuple(h.key, h.seq, h.first_member);
      Server::Op& op = s->ops[key];
      if (op.nmembers == 0) {
        op.nmembers = h.nmembers;
        op.bytes = h.bytes;
        op.data.assign((size_t)h.nmembers * h.bytes, 0);
      }
      if (op.nmembers == h.nmembers && op.bytes == h.bytes && h.idx < h.nmembers) {
        if (h.bytes) memcpy(op.data.data() + (size_t)h.idx * h.bytes, payload.data(), h.bytes);
        op.waiters.emplace_back(peer_rank, h.req);
        if (++op.arrived == op.nmembers) {
          waiters = std::move(op.waiters);
          result = std::move(op.data);
          s->ops.erase(key);
        }
      }
    }
    for (auto& w : waiters) {
      CtlMsg r;
      memset(&r, 0, sizeof(r));
      r.type = MSG_REPLY;
      r.req = w.second;
      r.bytes = (uint32_t)result.size();
      s->send_to(w.first, r, result.data());
    }
  }
}

}  // namespace mlslb
