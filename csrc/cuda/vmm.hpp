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
  size_t bytes = 0;                  // size of the initial slab (= what the multicast object covers)
  size_t reserved = 0;               // address range reserved per rank (>= bytes): room to grow
  size_t local_mapped = 0;           // bytes of my slab that are backed by memory now
  std::vector<size_t> peer_mapped;   // the same for every peer's slab in my address space
  size_t gran = 0;
  int device = 0;
  std::vector<unsigned long long> h_grown;   // extra chunks (mine and imported), released at destroy
  std::string why;                   // reason when !ok / mc == nullptr
  // opaque driver handles (unsigned long long to keep cuda.h out of this header)
  unsigned long long h_local = 0, h_mc = 0;
  std::vector<unsigned long long> h_peers;
};

// Collective over the whole bootstrap world.  On failure on ANY rank every rank returns ok == false (and has
// released whatever it had set up) so the caller can fall back to cudaMalloc + CUDA IPC consistently.
// `reserve_bytes` of address space per rank (0 = just `bytes`): vmm_slab_grow / vmm_slab_map_peer_chunk fill it later.
VmmSlab vmm_slab_create(Bootstrap* boot, int device, size_t bytes, bool want_multicast, size_t reserve_bytes = 0);
// Back the next `add_bytes` (rounded up to the granularity) of MY reservation with a fresh chunk; returns the descriptor
// to publish (kept open: peers read it through /proc) or -1.  *off / *got receive the chunk's place and size.
int vmm_slab_grow(VmmSlab& s, size_t add_bytes, size_t* off, size_t* got);
// Map a chunk a peer published at `off` of its slab (fd already opened in this process).
bool vmm_slab_map_peer_chunk(VmmSlab& s, int peer, int fd, size_t off, size_t bytes);
void vmm_slab_destroy(VmmSlab& s);

}  // namespace mlslb
