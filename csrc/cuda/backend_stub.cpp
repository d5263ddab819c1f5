// Host-only build (make NO_CUDA=1 / TSAN=1): the device backend is absent.
#include "core/runtime.hpp"
#include "core/sysinfo.hpp"

namespace mlslb {
std::unique_ptr<Backend> make_cuda_backend(RankContext*) { return nullptr; }
bool cuda_backend_available() { return false; }
void cuda_fill_sysinfo(SysInfo&) {}
}  // namespace mlslb
