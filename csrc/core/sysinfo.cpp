#include "sysinfo.hpp"

#include <dirent.h>

#include <cstdio>
#include <cstring>
#include <set>

#include "log.hpp"
#include "runtime.hpp"

namespace mlslb {

const char* cpu_kind_name(CpuKind k) {
  switch (k) {
    case CpuKind::XEON: return "Xeon";
    case CpuKind::XEON_PHI: return "Xeon Phi";
    case CpuKind::EPYC: return "EPYC";
    case CpuKind::OTHER: return "other";
    default: return "unknown";
  }
}
const char* net_kind_name(NetKind k) {
  switch (k) {
    case NetKind::ETHERNET: return "Ethernet";
    case NetKind::INFINIBAND: return "InfiniBand";
    case NetKind::OMNIPATH: return "Omni-Path";
    default: return "none";
  }
}

void cuda_fill_sysinfo(SysInfo& s);   // csrc/cuda: no-op when no device is usable

SysInfo probe_system() {
  SysInfo s;
  if (FILE* f = fopen("/proc/cpuinfo", "r")) {
    char line[512];
    std::set<std::pair<int, int>> cores;
    int phys = 0, core = 0;
    while (fgets(line, sizeof(line), f)) {
      if (!strncmp(line, "model name", 10) && s.cpu_model.empty()) {
        const char* c = strchr(line, ':');
        if (c) {
          s.cpu_model = c + 2;
          while (!s.cpu_model.empty() && (s.cpu_model.back() == '\n' || s.cpu_model.back() == ' ')) s.cpu_model.pop_back();
        }
      } else if (!strncmp(line, "processor", 9)) {
        s.threads++;
      } else if (!strncmp(line, "physical id", 11)) {
        const char* c = strchr(line, ':');
        phys = c ? atoi(c + 1) : 0;
      } else if (!strncmp(line, "core id", 7)) {
        const char* c = strchr(line, ':');
        core = c ? atoi(c + 1) : 0;
        cores.insert({phys, core});
      }
    }
    fclose(f);
    s.cores = cores.empty() ? s.threads : (int)cores.size();
    if (s.cpu_model.find("Phi") != std::string::npos) s.cpu = CpuKind::XEON_PHI;
    else if (s.cpu_model.find("Xeon") != std::string::npos) s.cpu = CpuKind::XEON;
    else if (s.cpu_model.find("EPYC") != std::string::npos) s.cpu = CpuKind::EPYC;
    else s.cpu = s.cpu_model.empty() ? CpuKind::UNKNOWN : CpuKind::OTHER;
  }
  s.net = NetKind::ETHERNET;
  if (DIR* d = opendir("/sys/class/infiniband")) {
    while (dirent* e = readdir(d)) {
      if (e->d_name[0] == '.') continue;
      s.net_device = e->d_name;
      if (!strncmp(e->d_name, "hfi", 3)) s.net = NetKind::OMNIPATH;
      else s.net = NetKind::INFINIBAND;
      break;
    }
    closedir(d);
  }
  cuda_fill_sysinfo(s);
  return s;
}

void auto_config(RankContext* ctx) {
  SysInfo s = probe_system();
  int type = ctx->env.auto_config;
  if (ctx->rank == 0) {
    if (type == 1 || type == 3)
      MLSLB_LOG(LOG_DEBUG, "net type: %s %s", net_kind_name(s.net), s.net_device.c_str());
    if (type == 2 || type == 3)
      MLSLB_LOG(LOG_DEBUG, "cpu type: %s (%s), cores: %d, threads: %d", cpu_kind_name(s.cpu), s.cpu_model.c_str(),
                s.cores, s.threads);
    if (s.gpus)
      MLSLB_LOG(LOG_INFO, "gpu: %d x %s (sm_%d%d, %d SMs) peer_access=%d multicast=%d", s.gpus, s.gpu_name.c_str(),
                s.cc_major, s.cc_minor, s.sms, (int)s.peer_access, (int)s.multicast);
  }
  // Tuning decisions.  The reference bumps MLSL_LARGE_MSG_CHUNKS to 128 on Ethernet; the device path instead keys
  // on the GPU interconnect: without peer access big messages are chunked finely so staging copies pipeline.
  if (ctx->backend && ctx->backend->is_device()) {
    if (!s.peer_access && !getenv("MLSL_LARGE_MSG_CHUNKS")) ctx->env.large_msg_chunks = 16;
    if (!s.multicast) ctx->env.use_nvls = false;
  } else if (s.net == NetKind::ETHERNET && !getenv("MLSL_LARGE_MSG_CHUNKS") && ctx->world > 1 && !ctx->boot->inproc()) {
    ctx->env.large_msg_chunks = 4;   // single-node shm transport: nothing to gain from the Ethernet setting
  }
}

}  // namespace mlslb
