// Per-rank runtime: process groups, communication requests, backend interface, progress engine.
//
// This is the layer the reference calls "internal comm interface" (reference src/comm.hpp:48-424:
// ProcessGroup, CommOp*, CommDesc, CommRequest, Comm* free functions) plus the endpoint-proxy runtime behind it
// (reference eplib/cqueue.c, eplib/server.c).  B200 re-design:
//   * a request is a POD descriptor + state word; the data plane is a Backend (host shared memory, or CUDA
//     peer-memory kernels) instead of MPI calls;
//   * the client<->ep_server shared-memory command ring becomes an in-process SPSC ring consumed by background
//     progress thread(s) ("servers") that drive CUDA streams (or execute host reductions);
//   * process groups are rank lists + a row in the symmetric signal pad - creating one allocates no transport.
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "bootstrap.hpp"
#include "common.hpp"
#include "env.hpp"
#include "heap.hpp"

namespace mlslb {

struct RankContext;
class Backend;
class CommRequest;

// ----------------------------------------------------------------------------------------------------------
struct ProcessGroup {
  RankContext* ctx = nullptr;
  std::vector<int> members;   // global ranks in group order
  int idx = 0;                // this rank's index inside members
  int row = -1;               // signal-pad row (-1: self group, never communicates)
  uint64_t seq[2] = {0, 0};   // collectives issued per lane (0 normal, 1 priority); identical on every member
  uint64_t ctl_seq = 0;       // control-plane (host mailbox) collectives issued on this group
  bool is_world = false, is_self = false;
  uint64_t hwm() const { return std::max(ctl_seq, std::max(seq[0], seq[1])); }
  int size() const { return (int)members.size(); }
};

// What a request does (reference src/comm.hpp:250-366 "CommDesc" with exactly one op).
struct CommDesc {
  enum CompType { FPROP = 0, BPROP, PARAM_GRAD, PARAM_INC, GENERIC };
  OpKind kind = OpKind::BARRIER;
  DType dtype = DType::F32;
  RedOp rop = RedOp::SUM;
  ProcessGroup* group = nullptr;
  size_t count = 0;   // elements: Bcast/Reduce/AllReduce total; AlltoAll/Gather/AllGather per-peer send count;
                      // Scatter/ReduceScatter per-rank recv count
  size_t root = 0;
  std::vector<size_t> send_counts, send_offsets, recv_counts, recv_offsets;   // *v variants (elements)
  CompType comp_type = GENERIC;
  int64_t op_uid = -1;
  // ---- B200 extensions (fused epilogues) ----
  DType out_dtype = DType::F32;   // only honoured when has_out_dtype
  bool has_out_dtype = false;
  float scale = 1.0f;             // result multiplier fused into the reduction epilogue (e.g. 1/N averaging)
  bool compress = false;          // block-scaled fp8 quantised transport with error feedback
  // fused distributed update (OpKind::FUSED_UPDATE)
  struct FusedUpdate {
    int optimizer = 0;            // 0 = SGD(momentum), 1 = AdamW
    float lr = 0.f, momentum = 0.f, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f, weight_decay = 0.f;
    int64_t step = 0;
    void* param = nullptr;        // full parameter buffer (all-gathered in place, dtype = out_dtype)
    void* master = nullptr;       // fp32 owned master shard (may be null: param is the master)
    void* state1 = nullptr;       // momentum / exp_avg shard (fp32)
    void* state2 = nullptr;       // exp_avg_sq shard (fp32)
  } fused;
  // fused GEMM + reduce-scatter (OpKind::GEMM_RS): out[M/P, N] = reduce_scatter_rows(A[M, K] * W[N, K]^T)
  struct GemmRs {
    int M = 0, N = 0, K = 0;
    const void* a = nullptr;      // [M, K] bf16, row-major (activations, K-slice of this rank)
    const void* w = nullptr;      // [N, K] bf16, row-major (nn.Linear weight layout, K-slice of this rank)
  } gemm;
  // fused all-gather + GEMM (OpKind::AG_GEMM): Y[M, N] = concat_rows(X_0..X_{P-1}) * W[N, K]^T; gemm.M/N/K, gemm.a = this
  // rank's shard X_r [M/P, K], gemm.w = W; `gathered` receives the full X [M, K] (kept for backward)
  void* gathered = nullptr;
  // strided all-to-all (Activation::StartCommFused): every member pulls, from each peer's UNPACKED tensor,
  // the rectangle that is meant for it - `rows` rows of `row_bytes`, `src_stride` apart, starting `src_off` into the peer's
  // tensor (the geometry is the same on every member) - and writes it either packed into slot p of the receive region or,
  // with `dst_direct`, straight into its own unpacked tensor at `dst_off[p]` with rows `dst_stride` apart.  No pack kernel,
  // no unpack kernel, no trip of the payload through a packed send region.
  struct Strided {
    bool on = false, dst_direct = false;
    size_t rows = 0, row_bytes = 0, src_off = 0, src_stride = 0, src_total = 0;
    size_t dst_stride = 0, dst_total = 0;
    std::vector<size_t> dst_off;
  } strided;
};

class CommRequest {
 public:
  enum State : int { IDLE = 0, QUEUED = 1, LAUNCHED = 2, DONE = 3, FAILED = 4 };   // FAILED: launch threw on a progress thread
  std::string error;     // what the launch said (valid once state == FAILED); re-thrown by wait() / test()
  // device-measured duration of completed runs that nobody has read yet (CUDA backend: an event pair around the kernel,
  // harvested when the run is known to be complete); statistics and the trace take it with take_device_ns()
  uint64_t device_ns_pending = 0;
  uint64_t device_ns_last = 0;
  uint64_t take_device_ns() {
    uint64_t v = device_ns_pending;
    device_ns_pending = 0;
    return v;
  }
  CommRequest(RankContext* ctx, DType dt, int64_t uid, CommDesc::CompType ct);
  ~CommRequest();
  CommDesc desc;
  RankContext* ctx;
  bool one_shot = false;           // freed by Environment::Wait/Test (Distribution collectives)
  // Setup(): sizes derived from the op (reference src/comm_ep.cpp:568-766)
  void setup();
  size_t buf_bytes() const { return buf_bytes_; }      // bytes a comm buffer must have for Start(buf, buf+send_bytes)
  size_t send_bytes() const { return send_bytes_; }    // bytes of the send region (== tmpBufOffset for activations)
  size_t recv_bytes() const { return recv_bytes_; }
  size_t msg_bytes() const { return msg_bytes_; }      // payload size used for statistics / priority
  bool out_of_place_default() const { return oop_default_; }

  void start(void* send, void* recv);
  void* wait();                     // returns recv pointer
  void* test(bool* done);           // returns recv pointer when done else nullptr
  bool active() const { return state.load(std::memory_order_acquire) != IDLE; }

  // runtime state
  std::atomic<int> state{IDLE};
  void* send = nullptr;
  void* recv = nullptr;
  uint64_t group_seq = 0;           // ticket in the (group, lane) collective order
  int lane = 0;                     // 0 = normal, 1 = priority lane (own signal row + high-priority stream)
  uint64_t start_ns = 0, done_ns = 0;
  void* backend_state = nullptr;    // owned by the backend (events, staging buffers)
  bool setup_done = false;

 private:
  size_t buf_bytes_ = 0, send_bytes_ = 0, recv_bytes_ = 0, msg_bytes_ = 0;
  bool oop_default_ = false;
};

struct BlockDesc {
  size_t mb_off, mb_cnt, fm_off, fm_cnt, fm_size, buf_off;
};

// ----------------------------------------------------------------------------------------------------------
class Backend {
 public:
  virtual ~Backend() {}
  virtual const char* name() const = 0;
  virtual bool supports_strided_alltoall() const { return false; }
  virtual bool peek_done(CommRequest&) { return true; }         // launched collective finished? (never consumes it)
  virtual int default_servers() const { return 0; }              // progress threads when MLSL_NUM_SERVERS is unset
  virtual bool stream_ordered_wait() const { return false; }   // Wait only orders a stream (the host never blocks)
  virtual void harvest_device_time(CommRequest&) {}             // fold finished device timings into the request
  virtual bool is_device() const { return false; }
  virtual void* alloc(size_t bytes, size_t align) = 0;
  virtual void free(void* p) = 0;
  virtual bool owns(const void* p, size_t len) const = 0;       // inside this rank's symmetric heap?
  virtual void group_created(ProcessGroup&) {}
  virtual void group_destroyed(ProcessGroup&) {}
  virtual void prepare(CommRequest&) {}
  virtual void release(CommRequest&) {}
  // Issue the collective.  Device backends return once the work is enqueued on a stream; the host backend
  // returns when the collective has completed.  Must move r.state to LAUNCHED (or DONE).
  virtual void launch(CommRequest& r) = 0;
  virtual void on_start(CommRequest&) {}      // called on the API thread inside Start(), before the hand-off
  virtual bool test(CommRequest& r) = 0;
  virtual void wait(CommRequest& r) = 0;
  virtual void set_user_stream(void*) {}
  virtual void* user_stream() { return nullptr; }
  virtual void set_wait_mode(bool /*stream_ordered*/) {}
  // Strided (mb, fm, fmSize) gather/scatter between a local activation tensor and the comm buffer
  // (reference: the user-side triple loop of tests/examples/mlsl_test/mlsl_test.cpp:214-254).
  virtual void pack_blocks(const BlockDesc* blocks, size_t nblocks, size_t local_fm_count, DType dt,
                           const void* src, void* dst, bool unpack);
  virtual bool is_device_pointer(const void*) const { return false; }
  virtual void copy_from_host(void* dst, const void* src, size_t bytes);   // blocking; default memcpy
  // One-sided access (RMA windows): heap pointers travel between ranks as offsets; every rank can address every
  // peer's heap.  rma_copy is ordered like any other work of the caller (stream order on the device, immediate on the host).
  virtual uint64_t heap_offset(const void* p) const = 0;
  virtual void* peer_heap_ptr(int global_rank, uint64_t offset) = 0;
  virtual void rma_copy(void* dst, const void* src, size_t bytes);         // default memcpy
  virtual void finalize() {}
  virtual std::string describe() const { return name(); }
};

std::unique_ptr<Backend> make_host_backend(RankContext* ctx);
std::unique_ptr<Backend> make_net_backend(RankContext* ctx);   // TCP mesh between nodes (net_backend.cpp)
// CPU building blocks shared by the host and net backends (host_backend.cpp)
void host_reduce(DType dt, void* dst, const std::vector<const void*>& srcs, size_t n, RedOp op, float scale);
void host_optimizer_step(const CommDesc::FusedUpdate& f, DType pdt, char* param_owned, const float* gsum, size_t n);
std::unique_ptr<Backend> make_cuda_backend(RankContext* ctx);   // null when no usable GPU
bool cuda_backend_available();

// ----------------------------------------------------------------------------------------------------------
// Bounded single-producer / single-consumer ring (the descendant of reference eplib/cqueue.h:95-183: client
// fills table[tail], server drains table[head]; ours has back-pressure and C++11 acquire/release instead of
// volatile + x86-TSO).
template <typename T, size_t N>
class SpscRing {
 public:
  bool push(const T& v) {
    size_t t = tail_.load(std::memory_order_relaxed);
    if (t - head_.load(std::memory_order_acquire) >= N) return false;
    slots_[t % N] = v;
    tail_.store(t + 1, std::memory_order_release);
    return true;
  }
  bool pop(T& v) {
    size_t h = head_.load(std::memory_order_relaxed);
    if (h == tail_.load(std::memory_order_acquire)) return false;
    v = slots_[h % N];
    head_.store(h + 1, std::memory_order_release);
    return true;
  }
  size_t size() const { return tail_.load(std::memory_order_acquire) - head_.load(std::memory_order_acquire); }

 private:
  alignas(64) std::atomic<size_t> head_{0};
  alignas(64) std::atomic<size_t> tail_{0};
  alignas(64) T slots_[N];
};

struct Command {
  enum Kind : int { EXEC = 0, SUSPEND, RESUME, STOP };
  int kind = EXEC;
  CommRequest* req = nullptr;
  uint64_t arg = 0;                   // SUSPEND: generation to wait for
};

// Background "endpoint server" threads.  One ring per server; requests of one process group always go to the
// same server so the per-group launch order is the program order on every rank.
class ProgressEngine {
 public:
  ProgressEngine(RankContext* ctx, int num_servers);
  ~ProgressEngine();
  int servers() const { return (int)servers_.size(); }
  void submit(CommRequest* r);        // inline launch when there are no servers
  void drain();                       // block until every submitted command has been launched
  void suspend();                     // park the servers (reference EPLIB_suspend/EPLIB_execute)
  void resume();
  uint64_t launched() const { return launched_.load(); }
  // op uids of the most recently launched collectives (oldest first, at most 256): what order did the engine choose?
  std::vector<int64_t> recent_launches();

 private:
  struct Server {
    SpscRing<Command, 1024> ring;
    std::thread th;
    std::atomic<uint64_t> submitted{0}, completed{0};
    std::atomic<bool> parked{false};
    std::mutex mu;                    // producers are API threads: serialise pushes (SPSC per ring)
    // message prioritisation: commands taken off the ring but not launched yet, and what is in flight per (row, lane)
    std::vector<CommRequest*> pending;
    std::vector<CommRequest*> inflight;
  };
  void run(Server* s, int idx);
  void exec(Server* s, CommRequest* r);
  bool prioritised(const CommRequest* r) const;
  CommRequest* choose(Server* s);     // next pending command to launch under the ordering rules (nullptr: none yet)
  RankContext* ctx_;
  std::vector<std::unique_ptr<Server>> servers_;
  std::atomic<uint64_t> launched_{0};
  std::mutex recent_mu_;
  std::vector<int64_t> recent_;
  std::atomic<uint64_t> suspend_gen_{0}, resume_gen_{0};   // a server stays parked while resume_gen_ < its SUSPEND's generation
};

// ----------------------------------------------------------------------------------------------------------
// Registry of CommAlloc'ed ranges (reference src/pointer_checker.cpp:29-105, opt-in there at build time,
// MLSL_POINTER_CHECK=1 here): every buffer handed to a collective must lie inside one registered range.
class PointerChecker {
 public:
  void add(const void* p, size_t len);
  void remove(const void* p);
  // returns true when [p, p+len) is inside a registered range
  bool check(const void* p, size_t len) const;
  size_t count() const;

 private:
  mutable std::mutex mu_;
  std::set<std::pair<uintptr_t, uintptr_t>> ranges_;
};

struct QuantConfig {               // deep copy of the public QuantParams (reference include/mlsl.hpp:150-158)
  bool set = false;
  std::string lib_path, quant_name, dequant_name, reduce_name;
  size_t block_size = 0, elem_in_block = 0;
};

struct RankContext {
  EnvConfig env;
  std::unique_ptr<Bootstrap> boot;
  int rank = 0, world = 1;
  std::unique_ptr<Backend> backend;
  std::unique_ptr<ProgressEngine> progress;
  ProcessGroup* world_group = nullptr;     // all ranks of the job
  ProcessGroup* global_group = nullptr;    // the "global" group (world, or the Configure("color=") subset)
  ProcessGroup* self_group = nullptr;
  uint64_t row_used = 0;                   // bitmap of signal rows in use
  uint64_t seq_hwm = 0;                    // highest ticket this rank has used on any row (see create_group_by_color)
  int session_ops_hint = 0;                // operations in the committed session (priority-lane rule)
  std::mutex req_mu;
  std::set<CommRequest*> inflight;         // RequestStorage (reference src/mlsl_impl.hpp:60-94)
  PointerChecker ptrcheck;
  QuantConfig quant;
  std::atomic<int64_t> next_op_uid{0};
  int init_pid = 0;
  bool initialized = false;
  void* io_service = nullptr;              // file-IO offload thread (fileio.cpp), created on first use
  // MLSL_TRACE_FILE=<prefix>: every completed request becomes one slice of a Chrome / Perfetto trace
  // (<prefix>.<rank>.json, written at Finalize): start = Start() on the API thread, end = Wait/Test saw it complete
  struct TraceEvent {
    uint64_t t0, t1;
    int kind, row, lane;
    size_t bytes;
    uint64_t device_ns;   // duration of the kernel on the device (0 = not measured)
  };
  std::string trace_prefix;
  std::mutex trace_mu;
  std::vector<TraceEvent> trace;
  void trace_request(const CommRequest& r);
  void trace_dump();
  void* api_env = nullptr;                 // MLSL::impl::EnvironmentImpl bound to this context
  void (*api_env_free)(void*) = nullptr;   // its deleter (the type lives in graph.cpp)
  ~RankContext() {
    if (api_env && api_env_free) api_env_free(api_env);
  }

  ProcessGroup* create_group_by_color(ProcessGroup* parent, int color);   // collective over parent
  // members only: `used` = OR of the members' row bitmaps, `base` = max of their ticket marks (exchanged by the caller)
  ProcessGroup* create_group_from_members(const std::vector<int>& members, uint64_t used, uint64_t base);
  void free_group(ProcessGroup* g);
  void group_barrier(ProcessGroup* g);      // host control-plane barrier among the members
  void register_request(CommRequest* r);
  void remove_request(CommRequest* r);      // also frees one-shot requests
  void check_pointer(const void* p, size_t len, const char* what);
};

// Context lifecycle.  `bind_inproc_rank` makes the calling thread a virtual rank of an in-process world.
RankContext* current_context();             // thread-bound context if any, else the process-wide one
RankContext* process_context();
int inproc_world_create(int nranks);        // returns world id
void inproc_world_destroy(int world_id);
void inproc_bind_thread(int world_id, int rank);
void inproc_unbind_thread();
bool thread_is_inproc_rank();
std::unique_ptr<Bootstrap> take_thread_bootstrap();   // bootstrap reserved for the calling thread (or null)

// file-IO offload (fileio.cpp)
struct IoFile;
struct IoRequest;
IoFile* io_open(RankContext* ctx, const char* path);
size_t io_size(IoFile* f);
void io_close(IoFile* f);
IoRequest* io_read_nb(IoFile* f, void* dst, size_t bytes, long long offset);
IoRequest* io_open_read_close_nb(RankContext* ctx, const char* path, void* dst, size_t bytes, long long offset);
bool io_test(IoRequest* r, size_t* bytes_read);
size_t io_wait(IoRequest* r);          // frees the request
void io_shutdown(RankContext* ctx);
void install_signal_handlers(RankContext* ctx);   // sig_handler.cpp
void remove_signal_handlers();
void context_init(RankContext* ctx);        // bootstrap + backend + progress engine + base groups
void context_finalize(RankContext* ctx);

}  // namespace mlslb
