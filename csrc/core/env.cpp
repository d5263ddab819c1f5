#include "env.hpp"

#include <cstdlib>
#include <cstring>

#include "log.hpp"

namespace mlslb {

static const char* ev(const char* a, const char* b = nullptr) {
  const char* v = getenv(a);
  if (v && *v) return v;
  if (b) {
    v = getenv(b);
    if (v && *v) return v;
  }
  return nullptr;
}
static void geti(int& dst, const char* a, const char* b = nullptr) {
  if (const char* v = ev(a, b)) dst = atoi(v);
}
static void getb(bool& dst, const char* a) {
  if (const char* v = ev(a)) dst = atoi(v) != 0;
}
static void getz(size_t& dst, const char* a) {
  if (const char* v = ev(a)) dst = (size_t)strtoull(v, nullptr, 10);
}
static void gets(std::string& dst, const char* a, const char* b = nullptr) {
  if (const char* v = ev(a, b)) dst = v;
}

static const TuneDesc kTune[] = {
    {"ll", "MLSL_LL", &Tunables::ll, "flag-in-data latency kernels for small all-reduces"},
    {"mid_max_kb", "MLSL_MID_MAX_KB", &Tunables::mid_max_kb, "largest all-reduce on the multi-CTA flag-in-data kernel"},
    {"mid_oneshot_kb", "MLSL_MID_ONESHOT_KB", &Tunables::mid_oneshot_kb, "one-shot while (P-1)*bytes <= this"},
    {"nvls_min_ranks", "MLSL_NVLS_MIN_RANKS", &Tunables::nvls_min_ranks, "multicast kernels from this group size up"},
    {"ar_unroll", "MLSL_AR_UNROLL", &Tunables::ar_unroll, "vectors per thread and pass of the large all-reduce"},
    {"ar_channels", "MLSL_AR_CHANNELS", &Tunables::ar_channels, "CTAs of the large all-reduce"},
    {"ar_p2p_pct", "MLSL_AR_P2P_PCT", &Tunables::ar_p2p_pct, "multicast all-reduce: % of the message moved peer-to-peer concurrently"},
    {"ar_p2p_cta_pct", "MLSL_AR_P2P_CTA_PCT", &Tunables::ar_p2p_cta_pct, "... on this % of the CTAs"},
    {"nvls_chunk_mb", "MLSL_NVLS_CHUNK_MB", &Tunables::nvls_chunk_mb, "split giant multicast all-reduces"},
    {"bulk_copy_kb", "MLSL_BULK_COPY_KB", &Tunables::bulk_copy_kb, "cp.async.bulk rings for segments >= this"},
    {"nvls_collectives", "MLSL_NVLS_COLLECTIVES", &Tunables::nvls_collectives, "multimem reduce-scatter / bcast"},
    {"host_pipeline", "MLSL_HOST_PIPELINE", &Tunables::host_pipeline, "H2D / all-reduce / D2H pipeline for host buffers"},
    {"pipe_chunk_mb", "MLSL_PIPE_CHUNK_MB", &Tunables::pipe_chunk_mb, "chunk size of the host pipeline"},
    {"pipe_bufs", "MLSL_PIPE_BUFS", &Tunables::pipe_bufs, "device buffers of the host pipeline"},
    {"numa_bind", "MLSL_NUMA_BIND", &Tunables::numa_bind, "bind the process to the GPU's NUMA node"},
    {"gemm_2cta", "MLSL_GEMM_2CTA", &Tunables::gemm_2cta, "cta_group::2 GEMM + reduce-scatter"},
    {"ag_gemm", "MLSL_AG_GEMM", &Tunables::ag_gemm, "fused all-gather + GEMM"},
    {"nvtx", "MLSL_NVTX", &Tunables::nvtx, "NVTX ranges"},
    {"trace_launch", "MLSL_TRACE_LAUNCH", &Tunables::trace_launch, "stderr line per launch"},
    {"force_kernel_solo", "MLSL_FORCE_KERNEL_SOLO", &Tunables::force_kernel_solo, "1-rank groups run the peer kernels"},
    {"quant_mx", "MLSL_QUANT_MX", &Tunables::quant_mx, "fp8 transport: ue8m0 scale per 32 elements (MX) instead of fp32 per 128"},
    {"dev_timestamps", "MLSL_DEV_TIMESTAMPS", &Tunables::dev_timestamps, "device timestamps in statistics / trace"},
    {"loopback_rendezvous_ms", "MLSL_LOOPBACK_RENDEZVOUS_MS", &Tunables::loopback_rendezvous_ms,
     "ranks sharing a GPU wait this long on the host for their peers before launching"},
};
const TuneDesc* tune_table(size_t* n) {
  *n = sizeof(kTune) / sizeof(kTune[0]);
  return kTune;
}
bool tune_set(Tunables& t, const char* key, long value) {
  for (const TuneDesc& d : kTune)
    if (!strcmp(d.key, key) || !strcmp(d.env, key)) {
      t.*(d.field) = value;
      return true;
    }
  return false;
}
bool tune_get(const Tunables& t, const char* key, long* value) {
  for (const TuneDesc& d : kTune)
    if (!strcmp(d.key, key) || !strcmp(d.env, key)) {
      *value = t.*(d.field);
      return true;
    }
  return false;
}
void parse_tunables(Tunables& t) {
  for (const TuneDesc& d : kTune)
    if (const char* v = ev(d.env)) t.*(d.field) = strtol(v, nullptr, 10);
}

EnvConfig parse_env() {
  EnvConfig c;
  parse_tunables(c.tune);
  geti(c.log_level, "MLSL_LOG_LEVEL");
  getb(c.stats, "MLSL_STATS");
  getb(c.dup_group, "MLSL_DUP_GROUP");
  geti(c.auto_config, "MLSL_AUTO_CONFIG_TYPE");
  geti(c.num_servers, "MLSL_NUM_SERVERS", "EPLIB_MAX_EP_PER_TASK");
  gets(c.server_affinity, "MLSL_SERVER_AFFINITY", "EPLIB_SERVER_AFFINITY");
  geti(c.num_channels, "MLSL_NUM_CHANNELS");
  if (const char* v = ev("MLSL_HEAP_SIZE_GB", "EPLIB_SHM_SIZE_GB")) c.heap_size_gb = atof(v);
  if (const char* v = ev("MLSL_HEAP_MAX_GB")) c.heap_max_gb = atof(v);
  getb(c.check_mem_size, "MLSL_CHECK_MEM_SIZE");
  getz(c.max_short_msg, "MLSL_MAX_SHORT_MSG_SIZE");
  getz(c.large_msg_mb, "MLSL_LARGE_MSG_SIZE_MB");
  geti(c.large_msg_chunks, "MLSL_LARGE_MSG_CHUNKS");
  geti(c.alltoall_split, "MLSL_ALLTOALL_SPLIT");
  geti(c.alltoallv_split, "MLSL_ALLTOALLV_SPLIT");
  getb(c.msg_priority, "MLSL_MSG_PRIORITY");
  getz(c.msg_priority_threshold, "MLSL_MSG_PRIORITY_THRESHOLD");
  geti(c.msg_priority_mode, "MLSL_MSG_PRIORITY_MODE");
  getb(c.check_single_node, "MLSL_CHECK_SINGLE_NODE");
  getb(c.pointer_check, "MLSL_POINTER_CHECK");
  gets(c.backend, "MLSL_BACKEND");
  getb(c.use_nvls, "MLSL_NVLS");
  geti(c.watchdog_sec, "MLSL_WATCHDOG_SEC");
  gets(c.wait_mode, "MLSL_WAIT_MODE");
  gets(c.job_id, "MLSL_JOB_ID");
  geti(c.rank, "MLSL_RANK", "RANK");
  geti(c.world, "MLSL_WORLD_SIZE", "WORLD_SIZE");
  geti(c.local_rank, "MLSL_LOCAL_RANK", "LOCAL_RANK");
  geti(c.inproc_ranks, "MLSL_INPROC_RANKS");
  gets(c.master_addr, "MLSL_MASTER_ADDR", "MASTER_ADDR");
  if (const char* v = ev("MLSL_MASTER_PORT")) c.master_port = atoi(v);
  else if (const char* t = ev("MASTER_PORT")) c.master_port = atoi(t) + 1;   // torch's own store owns MASTER_PORT itself
  else c.master_port = 29571;
  gets(c.dynamic_server, "MLSL_DYNAMIC_SERVER", "EPLIB_DYNAMIC_SERVER");
  if (c.dynamic_server == "disable") c.num_servers = 0;
  getz(c.thp_threshold_mb, "MLSL_THP_THRESHOLD_MB");
  if (const char* v = ev("EPLIB_THP_THRESHOLD_MB")) if (!ev("MLSL_THP_THRESHOLD_MB")) c.thp_threshold_mb = (size_t)strtoull(v, nullptr, 10);
  gets(c.hostname, "MLSL_HOSTNAME", "EPLIB_HOSTNAME");
  geti(c.hostname_type, "MLSL_HOSTNAME_TYPE", "EPLIB_HOSTNAME_TYPE");
  gets(c.iface_name, "MLSL_IFACE_NAME", "EPLIB_IFACE_NAME");
  geti(c.iface_idx, "MLSL_IFACE_IDX", "EPLIB_IFACE_IDX");
  if (c.job_id.empty()) gets(c.job_id, "EPLIB_UUID");     // the reference's shared-memory name key (eplib/env.c:373-407)
  // knobs of the reference's server processes / MPI glue: accepted, nothing to configure
  for (const char* name : {"MLSL_SERVER_CREATION_TYPE", "MLSL_SERVER_PREFIX", "MLSL_USE_COPY_THREADS", "MLSL_COPY_THREADS",
                           "MLSL_COPY_THRESHOLD", "MLSL_MPI_VERSION_CHECK", "EPLIB_USE_ALLOCATOR", "EPLIB_USE_MEM_HOOKS",
                           "EPLIB_STD_MPI_MODE", "EPLIB_MPI_THREAD_MULTIPLE", "EPLIB_ROOT"})
    if (ev(name)) c.not_applicable += std::string(c.not_applicable.empty() ? "" : " ") + name;
  gets(c.net_addr, "MLSL_NET_ADDR");
  if (const char* v = ev("MLSL_NET_EAGER_KB")) c.net_eager_kb = atol(v);
  if (const char* v = ev("MLSL_NET_ONESHOT_KB")) c.net_oneshot_kb = atol(v);
  if (const char* v = ev("MLSL_NET_CHUNK_KB")) c.net_chunk_kb = atol(v);
  if (const char* v = ev("MLSL_NET_HIER_KB")) c.net_hier_kb = atol(v);
  getb(c.net_shm, "MLSL_NET_SHM");
  getb(c.net_hier_pipeline, "MLSL_NET_HIER_PIPELINE");
  if (const char* v = ev("MLSL_NET_SOCKBUF_KB")) c.net_sockbuf_kb = atol(v);
  if (const char* v = ev("MLSL_NET_SHM_RING_KB")) c.net_shm_ring_kb = atol(v);
  if (const char* v = ev("MLSL_NET_EMULATE_GBIT")) c.net_emulate_gbit = atof(v);
  gets(c.node_rank, "MLSL_NODE_RANK", "GROUP_RANK");
  geti(c.stats_iters, "MLSL_STATS_ITERS");
  geti(c.stats_skip, "MLSL_STATS_SKIP");
  if (c.num_servers > 16) c.num_servers = 16;
  if (c.large_msg_chunks < 1) c.large_msg_chunks = 1;
  return c;
}

void print_env(const EnvConfig& c) {
  MLSLB_LOG(LOG_INFO, "MLSL_LOG_LEVEL=%d MLSL_STATS=%d MLSL_DUP_GROUP=%d MLSL_AUTO_CONFIG_TYPE=%d", c.log_level,
            (int)c.stats, (int)c.dup_group, c.auto_config);
  MLSLB_LOG(LOG_INFO, "MLSL_BACKEND=%s MLSL_NUM_SERVERS=%d MLSL_NUM_CHANNELS=%d MLSL_HEAP_SIZE_GB=%.2f MLSL_HEAP_MAX_GB=%.2f",
            c.backend.c_str(), c.num_servers, c.num_channels, c.heap_size_gb, c.heap_max_gb);
  MLSLB_LOG(LOG_INFO, "MLSL_MAX_SHORT_MSG_SIZE=%zu MLSL_LARGE_MSG_SIZE_MB=%zu MLSL_LARGE_MSG_CHUNKS=%d",
            c.max_short_msg, c.large_msg_mb, c.large_msg_chunks);
  MLSLB_LOG(LOG_INFO, "MLSL_MSG_PRIORITY=%d MLSL_MSG_PRIORITY_THRESHOLD=%zu MLSL_MSG_PRIORITY_MODE=%d",
            (int)c.msg_priority, c.msg_priority_threshold, c.msg_priority_mode);
  MLSLB_LOG(LOG_INFO, "MLSL_NVLS=%d MLSL_WAIT_MODE=%s MLSL_WATCHDOG_SEC=%d MLSL_CHECK_SINGLE_NODE=%d",
            (int)c.use_nvls, c.wait_mode.c_str(), c.watchdog_sec, (int)c.check_single_node);
  MLSLB_LOG(LOG_INFO, "MLSL_ALLTOALL_SPLIT=%d MLSL_ALLTOALLV_SPLIT=%d", c.alltoall_split, c.alltoallv_split);
  MLSLB_LOG(LOG_INFO, "MLSL_DYNAMIC_SERVER=%s (servers are progress threads) MLSL_SERVER_AFFINITY=%s MLSL_THP_THRESHOLD_MB=%zu",
            c.dynamic_server.empty() ? "thread" : c.dynamic_server.c_str(), c.server_affinity.c_str(), c.thp_threshold_mb);
  MLSLB_LOG(LOG_INFO, "MLSL_HOSTNAME=%s MLSL_HOSTNAME_TYPE=%d MLSL_IFACE_NAME=%s MLSL_IFACE_IDX=%d", c.hostname.c_str(), c.hostname_type,
            c.iface_name.c_str(), c.iface_idx);
  MLSLB_LOG(LOG_INFO, "MLSL_NET_ADDR=%s MLSL_NET_EAGER_KB=%ld MLSL_NET_ONESHOT_KB=%ld MLSL_NET_CHUNK_KB=%ld MLSL_NET_HIER_KB=%ld MLSL_NET_HIER_PIPELINE=%d MLSL_NET_SHM=%d "
            "MLSL_NET_SHM_RING_KB=%ld MLSL_NET_EMULATE_GBIT=%g MLSL_NODE_RANK=%s MLSL_NET_SOCKBUF_KB=%ld", c.net_addr.c_str(), c.net_eager_kb, c.net_oneshot_kb,
            c.net_chunk_kb, c.net_hier_kb, (int)c.net_hier_pipeline, (int)c.net_shm, c.net_shm_ring_kb, c.net_emulate_gbit, c.node_rank.c_str(),
            c.net_sockbuf_kb);
  MLSLB_LOG(LOG_INFO, "MLSL_JOB_TOKEN=%s", getenv("MLSL_JOB_TOKEN") ? "(set)" : "(not set)");
  // read where they are used, once, at start-up (device selection and slab / stream set-up of the CUDA backend, logging, tracing)
  auto shown = [](const char* name) { const char* v = getenv(name); return v ? v : "(unset)"; };
  MLSLB_LOG(LOG_INFO, "MLSL_DEVICE=%s MLSL_SLAB=%s MLSL_STREAM_MODE=%s MLSL_RANKS_PER_DEVICE=%s MLSL_ASSERT_MODE=%s MLSL_SIG_HANDLERS=%s "
            "MLSL_TRACE_FILE=%s", shown("MLSL_DEVICE"), shown("MLSL_SLAB"), shown("MLSL_STREAM_MODE"), shown("MLSL_RANKS_PER_DEVICE"),
            shown("MLSL_ASSERT_MODE"), shown("MLSL_SIG_HANDLERS"), shown("MLSL_TRACE_FILE"));
  if (!c.not_applicable.empty())
    MLSLB_LOG(LOG_INFO, "set but not applicable (no server processes, no MPI underneath): %s", c.not_applicable.c_str());
  for (const TuneDesc& d : kTune) MLSLB_LOG(LOG_INFO, "%s=%ld  (%s)", d.env, c.tune.*(d.field), d.help);
}

}  // namespace mlslb
