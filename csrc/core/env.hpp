// Runtime configuration read once at Environment::Init.
//
// The variable names follow the reference where the concept survives on a GPU node
// (reference src/env.cpp:21-40, src/comm_ep.cpp:1546-1699, eplib/env.c:373-407); knobs that only made sense
// for MPI/endpoints are accepted and mapped onto the Blackwell equivalents:
//   MLSL_NUM_SERVERS    -> number of background progress threads per rank (0 = launch inline, the reference's
//                          own default on a single node); on the host backend also the reduction worker count
//   MLSL_NUM_CHANNELS   -> CTAs per collective kernel ("endpoints" of the device path); 0 = auto by size
//   MLSL_HEAP_SIZE_GB   -> symmetric heap (device slab / host shm) size
//   MLSL_LARGE_MSG_*    -> chunked pipelining of very large collectives
//   MLSL_MSG_PRIORITY*  -> newest-first launch order for big gradient messages
#pragma once
#include <cstddef>
#include <string>

namespace mlslb {

// Tuning knobs of the device path: read once from the environment, listed by print_env, and changeable at run time
// through Environment::SetTuning(key, value) (all ranks must apply the same change at the same point of the program -
// most of them select kernels or grids, which have to agree across a group).  Replaces ad-hoc getenv() calls in the
// launch paths.
struct Tunables {
  long ll = 1;                  // MLSL_LL: flag-in-data latency kernels for small all-reduces
  long mid_max_kb = 1024;       // MLSL_MID_MAX_KB: largest all-reduce that takes the multi-CTA flag-in-data kernel
  long mid_oneshot_kb = 512;    // MLSL_MID_ONESHOT_KB: one-shot while (P-1) * bytes <= this, two-shot above
  long nvls_min_ranks = 4;      // MLSL_NVLS_MIN_RANKS: multicast kernels from this group size up
  long ar_unroll = 0;           // MLSL_AR_UNROLL: 16-byte vectors per thread and pass of the large all-reduce (0 = by size)
  long ar_channels = 0;         // MLSL_AR_CHANNELS: CTAs of the large all-reduce (0 = MLSL_NUM_CHANNELS / auto)
  long ar_p2p_pct = 0;          // MLSL_AR_P2P_PCT: multicast all-reduce - this % of a large message goes peer-to-peer at the same time
  long ar_p2p_cta_pct = 33;     // MLSL_AR_P2P_CTA_PCT: ... on this % of the CTAs
  long nvls_chunk_mb = 0;       // MLSL_NVLS_CHUNK_MB: split giant multicast all-reduces into launches of this size (0 = one launch)
  long bulk_copy_kb = 256;      // MLSL_BULK_COPY_KB: gather-like collectives move segments >= this with cp.async.bulk rings
  long nvls_collectives = 3;    // MLSL_NVLS_COLLECTIVES: bit 0 multimem.ld_reduce reduce-scatter / reduce, bit 1 multimem.st bcast,
                                // bit 2 multimem.st all-gather - taken when the members' buffers are symmetric
  long host_pipeline = 1;       // MLSL_HOST_PIPELINE: chunked H2D / all-reduce / D2H pipeline for host-resident buffers
  long pipe_chunk_mb = 16;      // MLSL_PIPE_CHUNK_MB: chunk size of that pipeline
  long pipe_bufs = 4;           // MLSL_PIPE_BUFS: device buffers of that pipeline (2..8)
  long numa_bind = 1;           // MLSL_NUMA_BIND: bind the process to the NUMA node of its GPU at init (one rank per process)
  long gemm_2cta = 1;           // MLSL_GEMM_2CTA: cta_group::2 GEMM + reduce-scatter kernel when the shape allows
  long ag_gemm = 1;             // MLSL_AG_GEMM: fused all-gather + GEMM kernel when the shape allows
  long nvtx = 1;                // MLSL_NVTX: NVTX range per collective launch
  long trace_launch = 0;        // MLSL_TRACE_LAUNCH: one stderr line per kernel launch
  long force_kernel_solo = 0;   // MLSL_FORCE_KERNEL_SOLO: single-rank groups run the peer kernels against themselves
  long quant_mx = 0;            // MLSL_QUANT_MX: fp8 transport with one ue8m0 (power-of-two) scale per 32 elements (MX) instead of
                                // one fp32 scale per 128 - same wire size; host and device backends implement the same format
  long dev_timestamps = 1;      // MLSL_DEV_TIMESTAMPS: statistics / trace use device event timestamps
  long loopback_rendezvous_ms = 20;   // MLSL_LOOPBACK_RENDEZVOUS_MS: ranks sharing a GPU wait this long on the HOST for their
                                      // peers before launching a collective (0 = launch at once and spin on the device)
};
struct TuneDesc {
  const char* key;
  const char* env;
  long Tunables::*field;
  const char* help;
};
const TuneDesc* tune_table(size_t* n);
bool tune_set(Tunables& t, const char* key, long value);   // false: unknown key
bool tune_get(const Tunables& t, const char* key, long* value);
void parse_tunables(Tunables& t);

struct EnvConfig {
  int log_level = 0;
  bool stats = false;            // MLSL_STATS
  bool dup_group = true;         // MLSL_DUP_GROUP: full-world data group gets its own signal row
  int auto_config = 1;           // MLSL_AUTO_CONFIG_TYPE
  int num_servers = -1;          // MLSL_NUM_SERVERS (-1 = unset)
  std::string server_affinity;   // MLSL_SERVER_AFFINITY: cpu list for progress threads
  int num_channels = 0;          // MLSL_NUM_CHANNELS (0=auto)
  double heap_size_gb = 4.0;     // MLSL_HEAP_SIZE_GB
  double heap_max_gb = 0.0;      // MLSL_HEAP_MAX_GB: address range reserved per rank for heap growth (0 = 8 x MLSL_HEAP_SIZE_GB,
                                 // at least 32 GiB); the device heap grows into it chunk by chunk when it runs full
  bool check_mem_size = false;   // MLSL_CHECK_MEM_SIZE
  size_t max_short_msg = 0;      // MLSL_MAX_SHORT_MSG_SIZE (elements): <= this -> single channel
  size_t large_msg_mb = 128;     // MLSL_LARGE_MSG_SIZE_MB
  int large_msg_chunks = 4;      // MLSL_LARGE_MSG_CHUNKS
  // MLSL_ALLTOALL_SPLIT / MLSL_ALLTOALLV_SPLIT (reference src/comm_ep.cpp:1192,1270: split every pair message across all
  // endpoints): >= 1 (and the default, -1) every pair message is spread over ALL channels, one pair after the other;
  // 0 = the channels are dealt to the pairs, each pair message moves on its own channels, all pairs at once
  int alltoall_split = -1;
  int alltoallv_split = -1;
  bool msg_priority = false;     // MLSL_MSG_PRIORITY
  size_t msg_priority_threshold = 10000;  // MLSL_MSG_PRIORITY_THRESHOLD (bytes)
  int msg_priority_mode = 1;     // MLSL_MSG_PRIORITY_MODE (1 = newest first)
  bool check_single_node = true; // MLSL_CHECK_SINGLE_NODE
  bool pointer_check = false;    // MLSL_POINTER_CHECK (reference: build-time ENABLE_CHKP_INT)
  std::string backend = "auto";  // MLSL_BACKEND: auto | host | cuda
  bool use_nvls = true;          // MLSL_NVLS: allow multicast path when available
  int watchdog_sec = 120;        // MLSL_WATCHDOG_SEC: flag-wait timeout before poison+abort (0=off)
  std::string wait_mode = "host";// MLSL_WAIT_MODE: host (block the CPU) | stream (order the user stream)
  std::string job_id;            // MLSL_JOB_ID (else derived from MASTER_PORT / TORCHELASTIC_RUN_ID)
  int rank = -1, world = -1, local_rank = -1;  // MLSL_RANK/RANK, MLSL_WORLD_SIZE/WORLD_SIZE, LOCAL_RANK
  std::string master_addr = "127.0.0.1";   // MLSL_MASTER_ADDR / MASTER_ADDR: where rank 0's control server listens (net backend)
  int master_port = 0;                      // MLSL_MASTER_PORT, else MASTER_PORT + 1, else 29571
  int inproc_ranks = 0;          // MLSL_INPROC_RANKS: >0 -> N virtual ranks inside this process (tests/loopback)
  int stats_iters = 10, stats_skip = 4;  // isolation statistics iterations (reference: 10 / skip 4)
  // MLSL_DYNAMIC_SERVER (reference src/comm_ep.cpp:1509-1531, eplib/env.c:62-81): process | thread | asyncthread | hybrid all
  // mean progress THREADS here (there are no server processes); "disable" = no servers, collectives run on the caller
  std::string dynamic_server;
  size_t thp_threshold_mb = 128; // MLSL_THP_THRESHOLD_MB: host allocations from this size are 2 MiB aligned + MADV_HUGEPAGE
                                 // (reference eplib/common.h:78-93)
  // which address the rank's data connections use across nodes (reference eplib/server.c:228-330: MLSL_HOSTNAME,
  // MLSL_HOSTNAME_TYPE 0 = as given by the launcher / 1 = name / 2 = IP, MLSL_IFACE_NAME prefix, MLSL_IFACE_IDX)
  std::string hostname, iface_name;
  int hostname_type = 0, iface_idx = -1;
  std::string not_applicable;    // reference knobs that were set but have nothing to act on here (listed at INFO)
  // ---- net backend (csrc/core/net_backend.cpp) ----
  std::string net_addr;          // MLSL_NET_ADDR: explicit address of this rank's data connections
  long net_eager_kb = 32;        // MLSL_NET_EAGER_KB: an early message up to this size is parked, a larger one stays in the socket
  long net_oneshot_kb = 32;      // MLSL_NET_ONESHOT_KB: all-reduce up to this size = one exchange of whole vectors
  long net_chunk_kb = 512;       // MLSL_NET_CHUNK_KB: reductions of slices >= 2x this travel in pieces of this size
  long net_hier_kb = 1024;       // MLSL_NET_HIER_KB: two-level all-reduce / all-gather / reduce-scatter from this size (-1 never)
  long net_sockbuf_kb = 0;       // MLSL_NET_SOCKBUF_KB: fixed send / receive buffers of the data sockets (0: kernel auto-tuning)
  bool net_hier_pipeline = true; // MLSL_NET_HIER_PIPELINE: the two-level collectives run piece by piece inside one exchange (MLSL_NET_CHUNK_KB)
  bool net_shm = true;           // MLSL_NET_SHM: same-node ranks exchange through shared-memory rings
  long net_shm_ring_kb = 1024;   // MLSL_NET_SHM_RING_KB: ring size per direction of a same-node pair
  double net_emulate_gbit = 0;   // MLSL_NET_EMULATE_GBIT: pace every rank's TCP egress (bench / test knob)
  std::string node_rank;         // MLSL_NODE_RANK / GROUP_RANK: which launcher ("node") on this machine the rank belongs to
  Tunables tune;                 // device-path knobs (see below)
};

EnvConfig parse_env();
void print_env(const EnvConfig& c);

}  // namespace mlslb
