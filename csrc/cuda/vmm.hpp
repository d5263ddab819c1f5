// Symmetric heap built on the CUDA virtual-memory-management API, with an NVLS multicast mapping on top.
//
// This is the device-side descendant of the reference's shared heap (eplib/memory.c:147-263: shm_open + mmap on the
// client, the same name opened by every server).  Every rank cuMemCreate()s its slab, exports it as a POSIX file
// descriptor, passes the descriptor to its peers over the bootstrap's Unix socket (SCM_RIGHTS) and maps everybody
// else's slab; rank 0 additionally creates a multicast object spanning all devices, every rank binds its slab to it
// and maps it: a store to the multicast address lands in ALL slabs at that offset, a `multimem.ld_reduce` returns
// the sum over ALL slabs - the reduction happens inside the NVSwitch (NVLS).  The driver API is reached through
// cudaGetDriverEntryPoint, so the library has no link-time dependency on libcuda (it must load on CPU-only hosts).
#pragma once
#include <cstddef>
#include <string>
#include <vector>

namespace mlslb {

class Bootstrap;

struct VmmSlab {
  bool ok = false;
  char* local = nullptr;             // my slab
  std::vector<char*> peers;          // every rank's slab in my address space (peers[me] == local)
  char* mc = nullptr;                // multicast mapping (nullptr when NVLS is unavailable)
  size_t bytes = 0;
  std::string why;                   // reason when !ok / mc == nullptr
  // opaque driver handles (unsigned long long to keep cuda.h out of this header)
  unsigned long long h_local = 0, h_mc = 0;
  std::vector<unsigned long long> h_peers;
};

// Collective over the whole bootstrap world.  On failure on ANY rank every rank returns ok == false (and has
// released whatever it had set up) so the caller can fall back to cudaMalloc + CUDA IPC consistently.
VmmSlab vmm_slab_create(Bootstrap* boot, int device, size_t bytes, bool want_multicast);
void vmm_slab_destroy(VmmSlab& s);

}  // namespace mlslb
