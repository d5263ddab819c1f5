// Rendezvous / out-of-band control plane.
//
// Replaces the reference's MPI bootstrap + hydra launcher (reference src/comm_ep.cpp:1496-1750,
// eplib/server.c:401-575) for a single NVSwitch node: ranks find each other through a POSIX shared-memory
// control block keyed by the job id (torchrun-compatible: RANK / WORLD_SIZE / MASTER_PORT), or - for tests and
// single-GPU loopback - live as N "virtual ranks" (threads) inside one process sharing a heap control block.
// The control block offers: a small all-gather mailbox, a barrier, named shared regions (host heaps, signal
// pads), file-descriptor passing (CUDA VMM / multicast handles), a poison word (fail-fast across ranks, the
// reference only has per-process _exit) and heartbeats for the watchdog.
#pragma once
#include <atomic>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "common.hpp"

namespace mlslb {

constexpr size_t kBootSlotBytes = 1024;

struct BootSlot {
  std::atomic<uint64_t> seq[2];
  char data[2][kBootSlotBytes];
};

constexpr size_t kGroupSlotBytes = 64;
constexpr size_t kOrderLogDepth = 256;
constexpr int kMaxGrowChunks = 24;

// Device-heap expansion (reference eplib/memory.c:396-410 registers a further shared-memory region with every server):
// a rank that outgrows its slab creates another physical chunk, maps it behind the slab in its own reserved address range,
// and publishes (pid, fd, offset, size) here; a watcher thread of every peer imports the descriptor through
// /proc/<pid>/fd/<fd>, maps the chunk at the same offset of ITS reservation for that rank and acknowledges.
struct GrowRec {
  std::atomic<uint64_t> gen;          // chunks published so far
  int64_t pid;
  struct {
    int64_t fd;
    uint64_t off, bytes;
  } chunk[kMaxGrowChunks];
};
struct GroupSlot {
  std::atomic<uint64_t> seq[2];
  char data[2][kGroupSlotBytes];
};

struct alignas(64) BootCtl {
  uint64_t magic;
  int32_t world;
  int32_t creator_pid;
  uint64_t creator_start;
  std::atomic<uint32_t> attached;
  std::atomic<uint32_t> detached;
  std::atomic<uint64_t> poison;        // non-zero: some rank failed; value = 1 + failing rank
  alignas(64) std::atomic<uint64_t> heartbeat[kMaxHostRanks];
  alignas(64) BootSlot slots[kMaxHostRanks];
  alignas(64) GroupSlot gslots[kMaxGroupRows][kMaxHostRanks];   // per signal-row mailboxes (sub-group control plane)
  // collectives launched so far per (signal row, lane) and rank: loop-back ranks (several ranks on one GPU) use it to
  // launch a collective's kernels together instead of letting the first one spin on the device (cuda_backend.cu)
  alignas(64) std::atomic<uint64_t> launch_seq[kMaxGroupRows * 2][kMaxHostRanks];
  // Launch-order log of a (signal row, lane): with message prioritisation on, the group's first member decides in which
  // order queued collectives go out (newest big gradient first) and publishes the decision here; the other members
  // replay it, so every rank launches the row's collectives in the same order (runtime.cpp: ProgressEngine).
  alignas(64) GrowRec grow[kMaxHostRanks];
  alignas(64) std::atomic<uint64_t> grow_ack[kMaxHostRanks][kMaxHostRanks];   // [owner][peer]: chunks of `owner` that `peer` has mapped
  alignas(64) std::atomic<uint64_t> order_head[kMaxGroupRows * 2];
  alignas(64) std::atomic<uint64_t> order_pos[kMaxGroupRows * 2][kMaxHostRanks];
  alignas(64) std::atomic<uint64_t> order_log[kMaxGroupRows * 2][kOrderLogDepth];
};

struct InprocWorld;   // shared state of an in-process world
class TcpControl;     // control plane of a job that spans nodes (tcp_control.hpp)

class Bootstrap {
 public:
  ~Bootstrap();
  // N virtual ranks in this process; returns one Bootstrap per rank (each to be used by its own thread).
  static std::vector<std::unique_ptr<Bootstrap>> create_inproc(int world);
  // Multi-process rendezvous through /dev/shm.
  static std::unique_ptr<Bootstrap> create_shm(const std::string& job_key, int rank, int world);
  // Multi-node rendezvous: rank 0 serves the control plane on master_addr:master_port (net backend).  All-gather,
  // barrier, the sub-group mailbox and poison work as above; regions and fd passing do not exist across nodes.
  static std::unique_ptr<Bootstrap> create_tcp(const std::string& master_addr, int master_port, int rank, int world);
  bool is_tcp() const { return tcp_ != nullptr; }

  int rank() const { return rank_; }
  int size() const { return world_; }
  bool inproc() const { return inproc_ != nullptr; }
  const std::string& key() const { return key_; }

  // in: `bytes` from every rank; out: world*bytes ordered by rank.  Any size (chunked internally).
  void allgather(const void* in, void* out, size_t bytes);
  void barrier();
  // Sub-group all-gather (<= kGroupSlotBytes per rank) among `members` (global ranks) using signal row `row`;
  // `seq` must be the same strictly increasing ticket on every member.
  void group_allgather(const std::vector<int>& members, int row, uint64_t seq, const void* in, void* out,
                       size_t bytes);

  // Named regions.  create_region: this rank creates+maps a zero-filled region other ranks can attach to.
  // Usage pattern: create -> barrier -> attach peers -> barrier -> seal_regions() (unlinks the names).
  void* create_region(const std::string& name, size_t bytes);
  void* attach_region(int owner_rank, const std::string& name, size_t bytes);
  void release_region(void* ptr, size_t bytes, bool owner, const std::string& name);
  void seal_regions();   // unlink every name this rank created (mappings stay valid)

  // File-descriptor exchange (multi-process only): every rank contributes one fd, receives world fds
  // (its own slot is a dup).  Used for CUDA VMM shareable handles.
  std::vector<int> allgather_fd(int fd);
  // Point-to-point descriptor passing over the same sockets, any time after start-up (device-heap expansion): send one
  // descriptor + 24 bytes to `rank`; poll for one (non-blocking, false when nothing is waiting).  Multi-process only.
  bool send_fd_to(int rank, int fd, const uint64_t payload[3]);
  bool try_recv_fd(int* fd, uint64_t payload[3]);

  // Fail-fast: mark the job poisoned / query it / beat.
  void poison(int code);
  uint64_t poisoned() const;
  void heartbeat();
  uint64_t peer_heartbeat(int r) const;
  BootCtl* ctl() { return ctl_; }

 private:
  Bootstrap() = default;
  int rank_ = 0, world_ = 1;
  std::string key_;
  BootCtl* ctl_ = nullptr;
  size_t ctl_bytes_ = 0;
  uint64_t round_ = 0;
  std::shared_ptr<InprocWorld> inproc_;
  std::shared_ptr<TcpControl> tcp_;
  std::vector<std::string> created_names_;
  int uds_fd_ = -1;
  std::string shm_name(int owner, const std::string& name) const;
  void wait_slots(uint64_t round);
};

}  // namespace mlslb
