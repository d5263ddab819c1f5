#include "cuda/vmm.hpp"

#include <cuda.h>
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>

#include "core/bootstrap.hpp"
#include "core/log.hpp"

namespace mlslb {

namespace {

struct Driver {
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  bool loaded = false;

  template <typename F>
  bool sym(const char* name, F& fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &qr) != cudaSuccess || !p) {
      cudaGetLastError();
      return false;
    }
    fn = reinterpret_cast<F>(p);
    return true;
  }
  bool load() {
    if (loaded) return true;
    bool ok = sym("cuMemCreate", MemCreate) && sym("cuMemRelease", MemRelease) &&
              sym("cuMemAddressReserve", MemAddressReserve) && sym("cuMemAddressFree", MemAddressFree) &&
              sym("cuMemMap", MemMap) && sym("cuMemUnmap", MemUnmap) && sym("cuMemSetAccess", MemSetAccess) &&
              sym("cuMemExportToShareableHandle", MemExportToShareableHandle) &&
              sym("cuMemImportFromShareableHandle", MemImportFromShareableHandle) &&
              sym("cuMemGetAllocationGranularity", MemGetAllocationGranularity) && sym("cuDeviceGet", DeviceGet) &&
              sym("cuDeviceGetAttribute", DeviceGetAttribute) && sym("cuGetErrorString", GetErrorString);
    // multicast entry points are optional (older drivers)
    sym("cuMulticastCreate", MulticastCreate);
    sym("cuMulticastAddDevice", MulticastAddDevice);
    sym("cuMulticastBindMem", MulticastBindMem);
    sym("cuMulticastUnbind", MulticastUnbind);
    sym("cuMulticastGetGranularity", MulticastGetGranularity);
    loaded = ok;
    return ok;
  }
  std::string err(CUresult r) {
    const char* s = nullptr;
    if (GetErrorString && GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
    return "CUresult " + std::to_string((int)r);
  }
};

Driver g_drv;

#define DRV_TRY(call, what)                                        \
  do {                                                             \
    CUresult _r = (call);                                          \
    if (_r != CUDA_SUCCESS) {                                      \
      s.why = std::string(what) + ": " + g_drv.err(_r);            \
      goto done;                                                   \
    }                                                              \
  } while (0)

// every rank reports its local status; the step succeeded only if it did everywhere
bool all_ok(Bootstrap* b, bool mine) {
  int v = mine ? 1 : 0;
  std::vector<int> all(b->size());
  b->allgather(&v, all.data(), sizeof(int));
  for (int x : all)
    if (!x) return false;
  return true;
}

}  // namespace

void vmm_slab_destroy(VmmSlab& s) {
  if (!g_drv.loaded) return;
  if (s.mc) {
    g_drv.MemUnmap((CUdeviceptr)s.mc, s.bytes);
    g_drv.MemAddressFree((CUdeviceptr)s.mc, s.bytes);
    s.mc = nullptr;
  }
  if (s.h_mc) {
    g_drv.MemRelease((CUmemGenericAllocationHandle)s.h_mc);
    s.h_mc = 0;
  }
  for (size_t p = 0; p < s.peers.size(); ++p) {
    if (!s.peers[p]) continue;
    const size_t mapped = p < s.peer_mapped.size() && s.peer_mapped[p] ? s.peer_mapped[p] : s.bytes;
    const size_t mine = s.peers[p] == s.local ? std::max(mapped, s.local_mapped) : mapped;
    g_drv.MemUnmap((CUdeviceptr)s.peers[p], mine);
    g_drv.MemAddressFree((CUdeviceptr)s.peers[p], s.reserved ? s.reserved : s.bytes);
    if (p < s.h_peers.size() && s.h_peers[p]) g_drv.MemRelease((CUmemGenericAllocationHandle)s.h_peers[p]);
  }
  for (unsigned long long h : s.h_grown) g_drv.MemRelease((CUmemGenericAllocationHandle)h);
  s.h_grown.clear();
  s.peers.clear();
  s.h_peers.clear();
  s.local = nullptr;
  s.h_local = 0;
  s.ok = false;
}

int vmm_slab_grow(VmmSlab& s, size_t add_bytes, size_t* off, size_t* got) {
  if (!s.ok || !g_drv.loaded) return -1;
  add_bytes = (add_bytes + s.gran - 1) / s.gran * s.gran;
  if (s.local_mapped + add_bytes > s.reserved) return -1;
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = s.device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemGenericAllocationHandle h = 0;
  if (g_drv.MemCreate(&h, add_bytes, &prop, 0) != CUDA_SUCCESS) return -1;
  const CUdeviceptr va = (CUdeviceptr)(s.local + s.local_mapped);
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = s.device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  int fd = -1;
  if (g_drv.MemMap(va, add_bytes, 0, h, 0) != CUDA_SUCCESS || g_drv.MemSetAccess(va, add_bytes, &acc, 1) != CUDA_SUCCESS ||
      g_drv.MemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS) {
    g_drv.MemRelease(h);
    return -1;
  }
  s.h_grown.push_back(h);
  *off = s.local_mapped;
  *got = add_bytes;
  s.local_mapped += add_bytes;
  return fd;
}

bool vmm_slab_map_peer_chunk(VmmSlab& s, int peer, int fd, size_t off, size_t bytes) {
  if (!s.ok || !g_drv.loaded || peer < 0 || peer >= (int)s.peers.size() || !s.peers[peer]) return false;
  if (off + bytes > s.reserved) return false;
  CUmemGenericAllocationHandle h = 0;
  if (g_drv.MemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) != CUDA_SUCCESS) return false;
  const CUdeviceptr va = (CUdeviceptr)(s.peers[peer] + off);
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = s.device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  if (g_drv.MemMap(va, bytes, 0, h, 0) != CUDA_SUCCESS || g_drv.MemSetAccess(va, bytes, &acc, 1) != CUDA_SUCCESS) {
    g_drv.MemRelease(h);
    return false;
  }
  s.h_grown.push_back(h);
  s.peer_mapped[peer] = std::max(s.peer_mapped[peer], off + bytes);
  return true;
}

VmmSlab vmm_slab_create(Bootstrap* boot, int device, size_t bytes, bool want_multicast, size_t reserve_bytes) {
  VmmSlab s;
  s.device = device;
  const int W = boot->size(), me = boot->rank();
  s.peers.assign(W, nullptr);
  s.h_peers.assign(W, 0);
  bool local_ok = false;
  int fd = -1;
  CUmemAllocationProp prop;
  CUmemAccessDesc acc;
  CUdevice cudev = 0;
  size_t gran = 0;
  std::vector<int> fds;
  memset(&prop, 0, sizeof(prop));
  memset(&acc, 0, sizeof(acc));
  if (!g_drv.load()) {
    s.why = "driver entry points unavailable";
    goto done;
  }
  cudaSetDevice(device);
  cudaFree(0);   // make sure the primary context exists
  DRV_TRY(g_drv.DeviceGet(&cudev, device), "cuDeviceGet");
  {
    int vmm = 0, fdok = 0;
    g_drv.DeviceGetAttribute(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, cudev);
    g_drv.DeviceGetAttribute(&fdok, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, cudev);
    if (!vmm || !fdok) {
      s.why = "device lacks VMM / POSIX fd handle support";
      goto done;
    }
  }
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  DRV_TRY(g_drv.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
  if (want_multicast && g_drv.MulticastGetGranularity) {
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof(mp));
    mp.numDevices = (unsigned)W;
    mp.size = bytes;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mg = 0;
    if (g_drv.MulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > gran) gran = mg;
  }
  if (gran == 0) gran = (size_t)2 << 20;
  bytes = (bytes + gran - 1) / gran * gran;
  s.bytes = bytes;
  s.gran = gran;
  s.reserved = std::max(bytes, (reserve_bytes + gran - 1) / gran * gran);
  s.local_mapped = bytes;
  s.peer_mapped.assign(W, 0);
  {
    CUmemGenericAllocationHandle h = 0;
    CUdeviceptr va = 0;
    DRV_TRY(g_drv.MemCreate(&h, bytes, &prop, 0), "cuMemCreate");
    s.h_local = h;
    s.h_peers[me] = h;
    DRV_TRY(g_drv.MemAddressReserve(&va, s.reserved, gran, 0, 0), "cuMemAddressReserve");
    s.peers[me] = (char*)va;
    s.peer_mapped[me] = bytes;
    DRV_TRY(g_drv.MemMap(va, bytes, 0, h, 0), "cuMemMap");
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    DRV_TRY(g_drv.MemSetAccess(va, bytes, &acc, 1), "cuMemSetAccess");
    s.local = (char*)va;
    DRV_TRY(g_drv.MemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle");
    local_ok = true;
  }
done:
  if (!all_ok(boot, local_ok)) {
    if (fd >= 0) close(fd);
    if (s.why.empty()) s.why = "a peer could not create its VMM slab";
    vmm_slab_destroy(s);
    return s;
  }
  // ---- map every peer's slab --------------------------------------------------------------------------------
  fds = boot->allgather_fd(fd);
  close(fd);
  bool map_ok = true;
  for (int p = 0; p < W && map_ok; ++p) {
    if (p == me) continue;
    CUmemGenericAllocationHandle h = 0;
    CUdeviceptr va = 0;
    CUresult r = g_drv.MemImportFromShareableHandle(&h, (void*)(uintptr_t)fds[p], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    if (r == CUDA_SUCCESS) {
      s.h_peers[p] = h;
      r = g_drv.MemAddressReserve(&va, s.reserved, gran, 0, 0);
    }
    if (r == CUDA_SUCCESS) {
      s.peers[p] = (char*)va;
      s.peer_mapped[p] = bytes;
      r = g_drv.MemMap(va, bytes, 0, h, 0);
    }
    if (r == CUDA_SUCCESS) r = g_drv.MemSetAccess(va, bytes, &acc, 1);
    if (r != CUDA_SUCCESS) {
      s.why = "mapping the slab of rank " + std::to_string(p) + ": " + g_drv.err(r);
      map_ok = false;
    }
  }
  for (int f : fds)
    if (f >= 0) close(f);
  if (!all_ok(boot, map_ok)) {
    if (s.why.empty()) s.why = "a peer could not map the slabs";
    vmm_slab_destroy(s);
    return s;
  }
  s.ok = true;

  // ---- NVLS: one multicast object over all devices, every slab bound at offset 0 ---------------------------------
  if (!want_multicast) {
    s.why = "multicast not requested";
    return s;
  }
  {
    int mcsup = 0;
    g_drv.DeviceGetAttribute(&mcsup, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cudev);
    bool have = mcsup && g_drv.MulticastCreate && g_drv.MulticastAddDevice && g_drv.MulticastBindMem;
    if (!all_ok(boot, have)) {
      s.why = "multicast (NVLS) is not supported on every device";
      return s;
    }
    CUmemGenericAllocationHandle mc = 0;
    int mfd = -1;
    bool ok = true;
    std::string why;
    if (me == 0) {
      CUmulticastObjectProp mp;
      memset(&mp, 0, sizeof(mp));
      mp.numDevices = (unsigned)W;
      mp.size = bytes;
      mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
      CUresult r = g_drv.MulticastCreate(&mc, &mp);
      if (r == CUDA_SUCCESS) r = g_drv.MemExportToShareableHandle(&mfd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
      if (r != CUDA_SUCCESS) {
        ok = false;
        why = "cuMulticastCreate: " + g_drv.err(r);
      }
    }
    if (!all_ok(boot, ok)) {
      s.why = why.empty() ? "rank 0 could not create the multicast object" : why;
      if (mc) g_drv.MemRelease(mc);
      return s;
    }
    int dummy = dup(0);
    std::vector<int> mfds = boot->allgather_fd(me == 0 ? mfd : dummy);
    close(dummy);
    if (mfd >= 0) close(mfd);
    if (me != 0) {
      CUresult r = g_drv.MemImportFromShareableHandle(&mc, (void*)(uintptr_t)mfds[0], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
      if (r != CUDA_SUCCESS) {
        ok = false;
        why = "importing the multicast object: " + g_drv.err(r);
      }
    }
    for (int f : mfds)
      if (f >= 0) close(f);
    if (ok) {
      CUresult r = g_drv.MulticastAddDevice(mc, cudev);
      if (r != CUDA_SUCCESS) {
        ok = false;
        why = "cuMulticastAddDevice: " + g_drv.err(r);
      }
    }
    if (!all_ok(boot, ok)) {   // doubles as the "all devices added" barrier required before binding
      s.why = why.empty() ? "a peer could not join the multicast object" : why;
      if (mc) g_drv.MemRelease(mc);
      return s;
    }
    CUdeviceptr mva = 0;
    CUresult r = g_drv.MulticastBindMem(mc, 0, (CUmemGenericAllocationHandle)s.h_local, 0, bytes, 0);
    if (r == CUDA_SUCCESS) r = g_drv.MemAddressReserve(&mva, bytes, gran, 0, 0);
    if (r == CUDA_SUCCESS) r = g_drv.MemMap(mva, bytes, 0, mc, 0);
    if (r == CUDA_SUCCESS) r = g_drv.MemSetAccess(mva, bytes, &acc, 1);
    if (r != CUDA_SUCCESS) {
      ok = false;
      why = "binding / mapping the multicast object: " + g_drv.err(r);
    }
    if (!all_ok(boot, ok)) {
      s.why = why.empty() ? "a peer could not bind the multicast object" : why;
      if (mva) {
        g_drv.MemUnmap(mva, bytes);
        g_drv.MemAddressFree(mva, bytes);
      }
      g_drv.MemRelease(mc);
      return s;
    }
    s.h_mc = mc;
    s.mc = (char*)mva;
  }
  return s;
}

}  // namespace mlslb
