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

struct EnvConfig {
  int log_level = 0;
  bool stats = false;            // MLSL_STATS
  bool dup_group = true;         // MLSL_DUP_GROUP: full-world data group gets its own signal row
  int auto_config = 1;           // MLSL_AUTO_CONFIG_TYPE
  int num_servers = -1;          // MLSL_NUM_SERVERS (-1 = unset)
  std::string server_affinity;   // MLSL_SERVER_AFFINITY: cpu list for progress threads
  int num_channels = 0;          // MLSL_NUM_CHANNELS (0=auto)
  double heap_size_gb = 4.0;     // MLSL_HEAP_SIZE_GB
  bool check_mem_size = false;   // MLSL_CHECK_MEM_SIZE
  size_t max_short_msg = 0;      // MLSL_MAX_SHORT_MSG_SIZE (elements): <= this -> single channel
  size_t large_msg_mb = 128;     // MLSL_LARGE_MSG_SIZE_MB
  int large_msg_chunks = 4;      // MLSL_LARGE_MSG_CHUNKS
  int alltoall_split = 0;        // MLSL_ALLTOALL_SPLIT
  int alltoallv_split = 0;       // MLSL_ALLTOALLV_SPLIT
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
};

EnvConfig parse_env();
void print_env(const EnvConfig& c);

}  // namespace mlslb
