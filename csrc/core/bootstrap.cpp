#include "bootstrap.hpp"

#include <fcntl.h>
#include <sched.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "log.hpp"
#include "tcp_control.hpp"

namespace mlslb {

static constexpr uint64_t kMagic = 0x4d4c534c42323030ull;  // "MLSLB200"

struct InprocWorld {
  BootCtl* ctl = nullptr;
  std::mutex mu;
  std::map<std::string, void*> regions;   // "<owner>/<name>" -> ptr
  ~InprocWorld() { free(ctl); }
};

static inline void cpu_relax() {
#if defined(__x86_64__)
  __builtin_ia32_pause();
#endif
}

static uint64_t proc_start_time(int pid) {
  char path[64];
  snprintf(path, sizeof(path), "/proc/%d/stat", pid);
  FILE* f = fopen(path, "r");
  if (!f) return 0;
  char buf[1024];
  size_t n = fread(buf, 1, sizeof(buf) - 1, f);
  fclose(f);
  buf[n] = 0;
  char* p = strrchr(buf, ')');   // comm may contain spaces; fields resume after the last ')'
  if (!p) return 0;
  p++;
  uint64_t val = 0;
  for (int field = 3; field <= 22; ++field) {
    while (*p == ' ') p++;
    if (field == 22) val = strtoull(p, nullptr, 10);
    while (*p && *p != ' ') p++;
  }
  return val;
}

Bootstrap::~Bootstrap() {
  if (tcp_) {
    tcp_->goodbye();
    tcp_.reset();
    return;
  }
  if (uds_fd_ >= 0) close(uds_fd_);
  if (!inproc_ && ctl_) {
    seal_regions();
    munmap(ctl_, ctl_bytes_);
  }
}

std::vector<std::unique_ptr<Bootstrap>> Bootstrap::create_inproc(int world) {
  MLSLB_ASSERT(world >= 1 && world <= kMaxHostRanks, "in-process world size %d out of range", world);
  auto w = std::make_shared<InprocWorld>();
  void* mem = nullptr;
  MLSLB_ASSERT(posix_memalign(&mem, 64, sizeof(BootCtl)) == 0, "oom");
  memset(mem, 0, sizeof(BootCtl));
  w->ctl = (BootCtl*)mem;
  w->ctl->magic = kMagic;
  w->ctl->world = world;
  std::vector<std::unique_ptr<Bootstrap>> out;
  for (int r = 0; r < world; ++r) {
    std::unique_ptr<Bootstrap> b(new Bootstrap());
    b->rank_ = r;
    b->world_ = world;
    b->ctl_ = w->ctl;
    b->inproc_ = w;
    b->key_ = "inproc";
    out.push_back(std::move(b));
  }
  return out;
}

std::string Bootstrap::shm_name(int owner, const std::string& name) const {
  char buf[256];
  snprintf(buf, sizeof(buf), "/mlslb_%d_%s_%s_%d", (int)getuid(), key_.c_str(), name.c_str(), owner);
  return buf;
}

std::unique_ptr<Bootstrap> Bootstrap::create_shm(const std::string& job_key, int rank, int world) {
  MLSLB_ASSERT(world >= 1 && world <= kMaxHostRanks, "world size %d out of range (max %d)", world, kMaxHostRanks);
  MLSLB_ASSERT(rank >= 0 && rank < world, "rank %d out of range", rank);
  std::unique_ptr<Bootstrap> b(new Bootstrap());
  b->rank_ = rank;
  b->world_ = world;
  b->key_ = job_key;
  b->ctl_bytes_ = round_up(sizeof(BootCtl), 4096);
  std::string name = b->shm_name(0, "ctl");
  if (rank == 0) {
    // Build the block under a private name, then publish atomically (a stale block of a crashed job with the
    // same key is replaced, never reused).
    std::string tmp = name + ".tmp";
    shm_unlink(tmp.c_str());
    int fd = shm_open(tmp.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    MLSLB_ASSERT(fd >= 0, "shm_open(%s) failed: %s", tmp.c_str(), strerror(errno));
    MLSLB_ASSERT(ftruncate(fd, (off_t)b->ctl_bytes_) == 0, "ftruncate failed: %s", strerror(errno));
    void* p = mmap(nullptr, b->ctl_bytes_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    MLSLB_ASSERT(p != MAP_FAILED, "mmap failed: %s", strerror(errno));
    close(fd);
    BootCtl* c = (BootCtl*)p;
    c->world = world;
    c->creator_pid = (int)getpid();
    c->creator_start = proc_start_time((int)getpid());
    c->attached.store(1);
    std::atomic_thread_fence(std::memory_order_release);
    c->magic = kMagic;
    std::string src = "/dev/shm" + tmp, dst = "/dev/shm" + name;
    MLSLB_ASSERT(rename(src.c_str(), dst.c_str()) == 0, "rename(%s) failed: %s", src.c_str(), strerror(errno));
    b->ctl_ = c;
  } else {
    uint64_t t0 = now_ns();
    for (;;) {
      int fd = shm_open(name.c_str(), O_RDWR, 0600);
      if (fd >= 0) {
        struct stat st;
        if (fstat(fd, &st) == 0 && (size_t)st.st_size >= b->ctl_bytes_) {
          void* p = mmap(nullptr, b->ctl_bytes_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
          close(fd);
          if (p != MAP_FAILED) {
            BootCtl* c = (BootCtl*)p;
            bool ok = c->magic == kMagic && c->world == world && c->creator_pid > 0 &&
                      (kill(c->creator_pid, 0) == 0 || errno == EPERM) &&
                      proc_start_time(c->creator_pid) == c->creator_start;
            if (ok) {
              b->ctl_ = c;
              c->attached.fetch_add(1);
              break;
            }
            munmap(p, b->ctl_bytes_);
          }
        } else {
          close(fd);
        }
      }
      MLSLB_ASSERT(now_ns() - t0 < 120ull * 1000000000ull,
                   "rank %d: timed out waiting for rank 0's control block %s", rank, name.c_str());
      usleep(2000);
    }
  }
  // Everyone is in: rank 0 removes the name so nothing is left behind whatever happens later.
  if (rank == 0) {
    uint64_t t0 = now_ns();
    while (b->ctl_->attached.load() < (uint32_t)world) {
      MLSLB_ASSERT(now_ns() - t0 < 120ull * 1000000000ull, "rank 0: only %u of %d ranks attached",
                   b->ctl_->attached.load(), world);
      usleep(1000);
    }
    shm_unlink(name.c_str());
  }
  // Datagram socket for fd passing (abstract namespace: vanishes with the process).
  b->uds_fd_ = socket(AF_UNIX, SOCK_DGRAM | SOCK_CLOEXEC, 0);
  if (b->uds_fd_ >= 0) {
    sockaddr_un a;
    memset(&a, 0, sizeof(a));
    a.sun_family = AF_UNIX;
    int n = snprintf(a.sun_path + 1, sizeof(a.sun_path) - 1, "mlslb_%d_%s_%d", (int)getuid(), job_key.c_str(), rank);
    if (bind(b->uds_fd_, (sockaddr*)&a, (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n)) != 0) {
      MLSLB_LOG(LOG_INFO, "uds bind failed (%s): fd passing unavailable", strerror(errno));
      close(b->uds_fd_);
      b->uds_fd_ = -1;
    }
  }
  b->barrier();
  return b;
}

void Bootstrap::wait_slots(uint64_t round) {
  int par = (int)(round & 1);
  uint64_t t0 = now_ns();
  uint64_t spins = 0;
  for (int r = 0; r < world_; ++r) {
    while (ctl_->slots[r].seq[par].load(std::memory_order_acquire) < round) {
      cpu_relax();
      if ((++spins & 0x3ff) == 0) {
        sched_yield();
        if (ctl_->poison.load(std::memory_order_relaxed) != 0)
          MLSLB_ASSERT(false, "job poisoned by rank %d while waiting in bootstrap", (int)ctl_->poison.load() - 1);
        if (now_ns() - t0 > 300ull * 1000000000ull)
          MLSLB_ASSERT(false, "bootstrap: rank %d never arrived at round %llu", r, (unsigned long long)round);
      }
    }
  }
}

std::unique_ptr<Bootstrap> Bootstrap::create_tcp(const std::string& master_addr, int master_port, int rank, int world) {
  MLSLB_ASSERT(world >= 1 && rank >= 0 && rank < world, "bad rank %d / world %d", rank, world);
  std::unique_ptr<Bootstrap> b(new Bootstrap());
  b->rank_ = rank;
  b->world_ = world;
  b->key_ = master_addr + ":" + std::to_string(master_port);
  b->tcp_ = std::make_shared<TcpControl>(master_addr, master_port, rank, world);
  b->barrier();   // everybody is connected before anyone proceeds
  return b;
}

void Bootstrap::allgather(const void* in, void* out, size_t bytes) {
  if (tcp_) {
    std::vector<int> all(world_);
    for (int r = 0; r < world_; ++r) all[r] = r;
    tcp_->gather(~0ull, ++round_, all, rank_, in, out, bytes);
    return;
  }
  size_t off = 0;
  do {
    size_t n = bytes - off < kBootSlotBytes ? bytes - off : kBootSlotBytes;
    uint64_t round = ++round_;
    int par = (int)(round & 1);
    if (n) memcpy(ctl_->slots[rank_].data[par], (const char*)in + off, n);
    ctl_->slots[rank_].seq[par].store(round, std::memory_order_release);
    wait_slots(round);
    for (int r = 0; r < world_; ++r)
      if (n) memcpy((char*)out + (size_t)r * bytes + off, ctl_->slots[r].data[par], n);
    off += n;
  } while (off < bytes);
}

void Bootstrap::barrier() { allgather(nullptr, nullptr, 0); }

void Bootstrap::group_allgather(const std::vector<int>& members, int row, uint64_t seq, const void* in, void* out,
                                size_t bytes) {
  MLSLB_ASSERT(bytes <= kGroupSlotBytes, "group_allgather payload too large");
  MLSLB_ASSERT(row >= 0 && row < kMaxGroupRows, "bad signal row %d", row);
  if (tcp_) {
    int idx = -1;
    for (size_t i = 0; i < members.size(); ++i)
      if (members[i] == rank_) idx = (int)i;
    MLSLB_ASSERT(idx >= 0, "group_allgather: this rank is not a member");
    tcp_->gather((uint64_t)row, seq, members, idx, in, out, bytes);
    return;
  }
  int par = (int)(seq & 1);
  GroupSlot* slots = ctl_->gslots[row];
  if (bytes) memcpy(slots[rank_].data[par], in, bytes);
  slots[rank_].seq[par].store(seq, std::memory_order_release);
  uint64_t t0 = now_ns(), spins = 0;
  for (size_t i = 0; i < members.size(); ++i) {
    int r = members[i];
    while (slots[r].seq[par].load(std::memory_order_acquire) < seq) {
      cpu_relax();
      if ((++spins & 0x3ff) == 0) {
        sched_yield();
        if (ctl_->poison.load(std::memory_order_relaxed) != 0)
          MLSLB_ASSERT(false, "job poisoned by rank %d while waiting in a group collective", (int)ctl_->poison.load() - 1);
        if (now_ns() - t0 > 300ull * 1000000000ull)
          MLSLB_ASSERT(false, "group control collective: rank %d never arrived (row %d seq %llu)", r, row,
                       (unsigned long long)seq);
      }
    }
    if (bytes) memcpy((char*)out + i * bytes, slots[r].data[par], bytes);
  }
}

void* Bootstrap::create_region(const std::string& name, size_t bytes) {
  MLSLB_ASSERT(!tcp_, "shared regions do not exist across nodes");
  bytes = round_up(bytes, 4096);
  if (inproc_) {
    void* p = nullptr;
    MLSLB_ASSERT(posix_memalign(&p, 4096, bytes) == 0, "oom allocating %zu bytes", bytes);
    memset(p, 0, bytes);
    std::lock_guard<std::mutex> g(inproc_->mu);
    inproc_->regions[std::to_string(rank_) + "/" + name] = p;
    return p;
  }
  std::string n = shm_name(rank_, name);
  shm_unlink(n.c_str());
  int fd = shm_open(n.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
  MLSLB_ASSERT(fd >= 0, "shm_open(%s) failed: %s", n.c_str(), strerror(errno));
  MLSLB_ASSERT(ftruncate(fd, (off_t)bytes) == 0, "ftruncate(%zu) failed: %s (is /dev/shm large enough?)", bytes,
               strerror(errno));
  void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  MLSLB_ASSERT(p != MAP_FAILED, "mmap(%zu) failed: %s", bytes, strerror(errno));
  close(fd);
  created_names_.push_back(n);
  return p;
}

void* Bootstrap::attach_region(int owner, const std::string& name, size_t bytes) {
  MLSLB_ASSERT(!tcp_, "shared regions do not exist across nodes");
  bytes = round_up(bytes, 4096);
  if (inproc_) {
    std::string k = std::to_string(owner) + "/" + name;
    for (;;) {
      {
        std::lock_guard<std::mutex> g(inproc_->mu);
        auto it = inproc_->regions.find(k);
        if (it != inproc_->regions.end()) return it->second;
      }
      sched_yield();
    }
  }
  std::string n = shm_name(owner, name);
  int fd = shm_open(n.c_str(), O_RDWR, 0600);
  MLSLB_ASSERT(fd >= 0, "attach: shm_open(%s) failed: %s", n.c_str(), strerror(errno));
  void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  MLSLB_ASSERT(p != MAP_FAILED, "attach: mmap failed: %s", strerror(errno));
  close(fd);
  return p;
}

void Bootstrap::release_region(void* ptr, size_t bytes, bool owner, const std::string& name) {
  bytes = round_up(bytes, 4096);
  if (inproc_) {
    if (owner) {
      std::lock_guard<std::mutex> g(inproc_->mu);
      inproc_->regions.erase(std::to_string(rank_) + "/" + name);
      free(ptr);
    }
    return;
  }
  munmap(ptr, bytes);
}

void Bootstrap::seal_regions() {
  for (auto& n : created_names_) shm_unlink(n.c_str());
  created_names_.clear();
}

std::vector<int> Bootstrap::allgather_fd(int fd) {
  std::vector<int> out(world_, -1);
  if (inproc_) {
    // same process: descriptors are directly shareable
    std::vector<int> all(world_);
    allgather(&fd, all.data(), sizeof(int));
    for (int r = 0; r < world_; ++r) out[r] = dup(all[r]);
    barrier();
    return out;
  }
  MLSLB_ASSERT(uds_fd_ >= 0, "fd passing socket unavailable");
  barrier();   // every socket is bound
  for (int r = 0; r < world_; ++r) {
    if (r == rank_) continue;
    sockaddr_un a;
    memset(&a, 0, sizeof(a));
    a.sun_family = AF_UNIX;
    int n = snprintf(a.sun_path + 1, sizeof(a.sun_path) - 1, "mlslb_%d_%s_%d", (int)getuid(), key_.c_str(), r);
    int payload = rank_;
    iovec iov{&payload, sizeof(payload)};
    char cbuf[CMSG_SPACE(sizeof(int))];
    memset(cbuf, 0, sizeof(cbuf));
    msghdr m;
    memset(&m, 0, sizeof(m));
    m.msg_name = &a;
    m.msg_namelen = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
    m.msg_iov = &iov;
    m.msg_iovlen = 1;
    m.msg_control = cbuf;
    m.msg_controllen = sizeof(cbuf);
    cmsghdr* c = CMSG_FIRSTHDR(&m);
    c->cmsg_level = SOL_SOCKET;
    c->cmsg_type = SCM_RIGHTS;
    c->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(c), &fd, sizeof(int));
    ssize_t s;
    do { s = sendmsg(uds_fd_, &m, 0); } while (s < 0 && (errno == EINTR || errno == EAGAIN));
    MLSLB_ASSERT(s == (ssize_t)sizeof(payload), "sendmsg(fd) to rank %d failed: %s", r, strerror(errno));
  }
  out[rank_] = dup(fd);
  for (int k = 0; k < world_ - 1; ++k) {
    int payload = -1;
    iovec iov{&payload, sizeof(payload)};
    char cbuf[CMSG_SPACE(sizeof(int))];
    msghdr m;
    memset(&m, 0, sizeof(m));
    m.msg_iov = &iov;
    m.msg_iovlen = 1;
    m.msg_control = cbuf;
    m.msg_controllen = sizeof(cbuf);
    ssize_t s;
    do { s = recvmsg(uds_fd_, &m, 0); } while (s < 0 && errno == EINTR);
    MLSLB_ASSERT(s == (ssize_t)sizeof(payload), "recvmsg(fd) failed: %s", strerror(errno));
    cmsghdr* c = CMSG_FIRSTHDR(&m);
    MLSLB_ASSERT(c && c->cmsg_type == SCM_RIGHTS && payload >= 0 && payload < world_, "malformed fd message");
    int got;
    memcpy(&got, CMSG_DATA(c), sizeof(int));
    out[payload] = got;
  }
  barrier();
  return out;
}

bool Bootstrap::send_fd_to(int r, int fd, const uint64_t payload[3]) {
  if (inproc_ || uds_fd_ < 0) return false;
  sockaddr_un a;
  memset(&a, 0, sizeof(a));
  a.sun_family = AF_UNIX;
  int n = snprintf(a.sun_path + 1, sizeof(a.sun_path) - 1, "mlslb_%d_%s_%d", (int)getuid(), key_.c_str(), r);
  uint64_t buf[3] = {payload[0], payload[1], payload[2]};
  iovec iov{buf, sizeof(buf)};
  char cbuf[CMSG_SPACE(sizeof(int))];
  memset(cbuf, 0, sizeof(cbuf));
  msghdr m;
  memset(&m, 0, sizeof(m));
  m.msg_name = &a;
  m.msg_namelen = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
  m.msg_iov = &iov;
  m.msg_iovlen = 1;
  m.msg_control = cbuf;
  m.msg_controllen = sizeof(cbuf);
  cmsghdr* c = CMSG_FIRSTHDR(&m);
  c->cmsg_level = SOL_SOCKET;
  c->cmsg_type = SCM_RIGHTS;
  c->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(c), &fd, sizeof(int));
  ssize_t s;
  do { s = sendmsg(uds_fd_, &m, 0); } while (s < 0 && (errno == EINTR || errno == EAGAIN));
  return s == (ssize_t)sizeof(buf);
}

bool Bootstrap::try_recv_fd(int* fd, uint64_t payload[3]) {
  if (inproc_ || uds_fd_ < 0) return false;
  uint64_t buf[3] = {0, 0, 0};
  iovec iov{buf, sizeof(buf)};
  char cbuf[CMSG_SPACE(sizeof(int))];
  msghdr m;
  memset(&m, 0, sizeof(m));
  m.msg_iov = &iov;
  m.msg_iovlen = 1;
  m.msg_control = cbuf;
  m.msg_controllen = sizeof(cbuf);
  ssize_t s = recvmsg(uds_fd_, &m, MSG_DONTWAIT);
  if (s != (ssize_t)sizeof(buf)) return false;
  cmsghdr* c = CMSG_FIRSTHDR(&m);
  if (!c || c->cmsg_type != SCM_RIGHTS) return false;
  memcpy(fd, CMSG_DATA(c), sizeof(int));
  payload[0] = buf[0];
  payload[1] = buf[1];
  payload[2] = buf[2];
  return true;
}

void Bootstrap::poison(int code) {
  if (tcp_) {
    tcp_->poison(code);
    return;
  }
  uint64_t expect = 0;
  ctl_->poison.compare_exchange_strong(expect, (uint64_t)code + 1);
}
uint64_t Bootstrap::poisoned() const { return tcp_ ? tcp_->poisoned() : ctl_->poison.load(std::memory_order_relaxed); }
void Bootstrap::heartbeat() {
  if (!tcp_) ctl_->heartbeat[rank_].fetch_add(1, std::memory_order_relaxed);
}
uint64_t Bootstrap::peer_heartbeat(int r) const { return tcp_ ? 0 : ctl_->heartbeat[r].load(std::memory_order_relaxed); }

}  // namespace mlslb
