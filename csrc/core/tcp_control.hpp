// TCP control plane for jobs that span nodes (the role MPI + hydra play for the reference across a cluster).
//
// Star topology: rank 0 runs a small control server, every rank (rank 0 included) holds one connection to it.
//   gather   : the members of a group each contribute `bytes`; when all have arrived the server sends every member the
//              concatenation in member order.  World all-gather, barrier and the sub-group mailbox of the shared-memory
//              bootstrap are all this one operation.
//   poison   : a failing rank tells the server, the server tells everyone; a connection that drops without a goodbye
//              poisons the job on behalf of the rank that owned it (fail-fast across nodes).
// A receiver thread per rank turns replies into completed requests and poison notices into a flag every wait loop polls.
// Volumes are tiny (addresses, group agreements); the data plane is the peer mesh of the net backend.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace mlslb {

// ---- socket helpers (also used by the net backend) ---------------------------------------------------------------
int tcp_listen(const std::string& addr, int port, int backlog, int* bound_port);       // port 0 = ephemeral
int tcp_connect_retry(const std::string& addr, int port, int timeout_sec);              // retries until the peer listens
void tcp_send_all(int fd, const void* buf, size_t bytes);                               // throws mlslb::Error
void tcp_recv_all(int fd, void* buf, size_t bytes);
void tcp_tune(int fd, long sockbuf_kb = 0);                                                                  // TCP_NODELAY; fixed socket buffers of that size if > 0
std::string tcp_local_address_towards(const std::string& addr, int port);              // my address on the route to addr
// IPv4 address of the first non-loop-back interface whose name starts with `prefix`, or - empty prefix - of the idx-th one
// ("" if none); and of a host name / dotted address
std::string tcp_address_of_interface(const std::string& prefix, int idx);
std::string tcp_resolve_to_ip(const std::string& host);
// What every connection of a job has to present in its first frame: a hash of MLSL_JOB_TOKEN (a fixed value when the variable
// is not set).  `mlslrun --hosts` hands every node the same random token; it keeps strays - another job that was given the
// same port, a port scanner - from being taken for a member (it is not a defence against someone who can read the environment).
uint64_t tcp_job_token();

class TcpControl {
 public:
  TcpControl(const std::string& master_addr, int master_port, int rank, int world);
  ~TcpControl();
  // members: global ranks in group order; idx: my position; key/seq identify the operation (same on every member)
  void gather(uint64_t key, uint64_t seq, const std::vector<int>& members, int idx, const void* in, void* out, size_t bytes);
  void poison(int code);
  uint64_t poisoned() const { return poison_.load(std::memory_order_relaxed); }
  void goodbye();   // orderly shutdown: my connection closing is not a failure

 private:
  struct Pending {
    bool done = false;
    std::vector<char> payload;
  };
  void rx_loop();
  void server_accept_loop();
  void server_client_loop(int fd, int peer_rank);
  int rank_, world_;
  int sock_ = -1;
  std::mutex tx_mu_, mu_;
  std::condition_variable cv_;
  std::map<uint64_t, Pending> pending_;
  uint64_t next_req_ = 1;
  std::atomic<uint64_t> poison_{0};
  std::atomic<bool> stopping_{false};
  std::thread rx_;
  // ---- server (rank 0 only) ----
  struct Server;
  std::unique_ptr<Server> srv_;
};

}  // namespace mlslb
