// System discovery + auto-configuration (reference src/sysinfo.cpp:57-201, src/mlsl.cpp:649-682: CPU family/cores
// from /proc/cpuinfo, fabric type from /sys/class/infiniband, and a tuning decision derived from them).  The B200
// edition adds what matters on an NVSwitch node: GPU count/model, SM count, peer access and multicast (NVLS)
// support, reported by the CUDA backend.
#pragma once
#include <string>

namespace mlslb {

struct RankContext;

enum class CpuKind { UNKNOWN = 0, XEON, XEON_PHI, EPYC, OTHER };
enum class NetKind { NONE = 0, ETHERNET, INFINIBAND, OMNIPATH };

struct SysInfo {
  CpuKind cpu = CpuKind::UNKNOWN;
  std::string cpu_model;
  int cores = 0, threads = 0;
  NetKind net = NetKind::NONE;
  std::string net_device;
  // device side (filled by the CUDA backend when active)
  int gpus = 0, sms = 0;
  std::string gpu_name;
  int cc_major = 0, cc_minor = 0;
  bool peer_access = false, multicast = false;
};

SysInfo probe_system();
const char* cpu_kind_name(CpuKind k);
const char* net_kind_name(NetKind k);
void auto_config(RankContext* ctx);    // MLSL_AUTO_CONFIG_TYPE: 0 off, 1 net, 2 cpu, 3 both

}  // namespace mlslb
