// EPLIB-style entry points (include/eplib.h) on top of the runtime: see the header for what maps to what.
// Reference: eplib/eplib.h (the public header), eplib/wrapper.c:880-1100 (init / memory / file wrappers).
#include "eplib.h"

#include <cstring>
#include <map>
#include <mutex>

#include "log.hpp"
#include "mlsl.h"
#include "mlsl.hpp"
#include "runtime.hpp"

namespace {

std::mutex g_mu;
bool g_inited_here = false;                  // EPLIB_init did the Environment::Init
std::map<void*, size_t> g_blocks;            // blocks handed out by EPLIB_malloc & co. (realloc needs the old size)
struct ReqInfo {
  size_t item;                               // bytes per item of the read behind an EPLIB_Request
  bool close_after;
};
std::map<EPLIB_Request, ReqInfo> g_reqs;

}  // namespace

struct eplib_file_s {
  mlslb::IoFile* f;
  long long pos;
};

extern "C" {

int EPLIB_init(void) {
  try {
    MLSL::Environment& e = MLSL::Environment::GetEnv();
    std::lock_guard<std::mutex> g(g_mu);
    if (!e.IsInitialized()) {
      e.Init(nullptr, nullptr);
      g_inited_here = true;
    }
    return 0;
  } catch (const std::exception& ex) {
    MLSLB_LOG(mlslb::LOG_ERROR, "EPLIB_init: %s", ex.what());
    return -1;
  }
}

int EPLIB_finalize(void) {
  try {
    std::lock_guard<std::mutex> g(g_mu);
    if (g_inited_here) {                     // (the reference: "already finalized, skip" - eplib/wrapper.c:931-936)
      MLSL::Environment& e = MLSL::Environment::GetEnv();
      if (e.IsInitialized()) e.Finalize();
      g_inited_here = false;
      g_blocks.clear();
    }
    return 0;
  } catch (const std::exception& ex) {
    MLSLB_LOG(mlslb::LOG_ERROR, "EPLIB_finalize: %s", ex.what());
    return -1;
  }
}

void* EPLIB_memalign(size_t alignment, size_t bytes) {
  try {
    void* p = MLSL::Environment::GetEnv().Alloc(bytes ? bytes : 1, alignment ? alignment : 64);
    std::lock_guard<std::mutex> g(g_mu);
    g_blocks[p] = bytes;
    return p;
  } catch (const std::exception&) {
    return nullptr;
  }
}
void* EPLIB_malloc(size_t bytes) { return EPLIB_memalign(64, bytes); }

void* EPLIB_calloc(size_t count, size_t size) {
  if (size && count > (size_t)-1 / size) return nullptr;
  void* p = EPLIB_memalign(64, count * size);
  // host-addressable backends only: device heaps are cleared by the caller's own kernels
  if (p && !MLSL::Environment::GetEnv().IsDeviceBackend()) memset(p, 0, count * size);
  return p;
}

void EPLIB_free(void* ptr) {
  if (!ptr) return;
  {
    std::lock_guard<std::mutex> g(g_mu);
    g_blocks.erase(ptr);
  }
  try {
    MLSL::Environment::GetEnv().Free(ptr);
  } catch (const std::exception&) {
  }
}

void* EPLIB_realloc(void* ptr, size_t bytes) {
  if (!ptr) return EPLIB_malloc(bytes);
  if (!bytes) {
    EPLIB_free(ptr);
    return nullptr;
  }
  size_t old = 0;
  {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_blocks.find(ptr);
    if (it == g_blocks.end()) return nullptr;            // not one of ours
    old = it->second;
  }
  void* q = EPLIB_malloc(bytes);
  if (!q) return nullptr;
  if (!MLSL::Environment::GetEnv().IsDeviceBackend()) memcpy(q, ptr, old < bytes ? old : bytes);
  EPLIB_free(ptr);
  return q;
}

int EPLIB_memory_is_shmem(void* ptr) {
  mlslb::RankContext* c = mlslb::current_context();
  return (ptr && c->initialized && c->backend->owns(ptr, 1)) ? 1 : 0;
}

void EPLIB_set_mem_hooks(void) {}

void* EPLIB_quant_params_submit(void* params) {
  if (params) {
    mlsl_environment env = 0;
    if (mlsl_environment_get_env(&env) == 0) mlsl_environment_set_quantization_params(env, (mlsl_quant_params*)params);
  }
  return params;
}

void EPLIB_execute(void) {
  try {
    MLSL::Environment::GetEnv().ResumeServers();
  } catch (const std::exception&) {
  }
}
void EPLIB_suspend(void) {
  try {
    MLSL::Environment::GetEnv().SuspendServers();
  } catch (const std::exception&) {
  }
}

EPLIB_FILE EPLIB_fopen(int, const char* filename, const char* mode) {
  if (!filename || !mode || mode[0] != 'r') return nullptr;
  try {
    mlslb::RankContext* c = mlslb::current_context();
    if (!c->initialized) return nullptr;
    eplib_file_s* s = new eplib_file_s{mlslb::io_open(c, filename), 0};
    return s;
  } catch (const std::exception&) {
    return nullptr;
  }
}

size_t EPLIB_fread_nb(int, void* buffer, size_t size, size_t count, EPLIB_FILE stream, EPLIB_Request* request) {
  if (!stream || !request || !size) return 0;
  try {
    const size_t total = mlslb::io_size(stream->f);
    size_t bytes = size * count;
    const size_t left = (size_t)stream->pos < total ? total - (size_t)stream->pos : 0;
    if (bytes > left) bytes = left / size * size;          // whole items only, like fread
    mlslb::IoRequest* r = mlslb::io_read_nb(stream->f, buffer, bytes, stream->pos);
    stream->pos += (long long)bytes;
    *request = (EPLIB_Request)(uintptr_t)r;
    std::lock_guard<std::mutex> g(g_mu);
    g_reqs[*request] = ReqInfo{size, false};
    return 0;                                               // the reference returns 0 from the non-blocking forms too
  } catch (const std::exception&) {
    *request = 0;
    return 0;
  }
}

int EPLIB_fwait(EPLIB_Request* request, size_t* readcount) {
  if (readcount) *readcount = 0;
  if (!request || !*request) return -1;
  size_t item = 1;
  {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_reqs.find(*request);
    if (it == g_reqs.end()) return -1;
    item = it->second.item;
    g_reqs.erase(it);
  }
  try {
    const size_t bytes = mlslb::io_wait((mlslb::IoRequest*)(uintptr_t)*request);
    if (readcount) *readcount = bytes / item;
    *request = 0;
    return 0;
  } catch (const std::exception&) {
    *request = 0;
    return -1;
  }
}

int EPLIB_fwaitall(int count, EPLIB_Request* requests, size_t* readcounts) {
  int rc = 0;
  for (int i = 0; i < count; ++i)
    if (EPLIB_fwait(&requests[i], readcounts ? &readcounts[i] : nullptr) != 0) rc = -1;
  return rc;
}

size_t EPLIB_fread(int epid, void* buffer, size_t size, size_t count, EPLIB_FILE stream) {
  EPLIB_Request r = 0;
  size_t n = 0;
  EPLIB_fread_nb(epid, buffer, size, count, stream, &r);
  if (!r || EPLIB_fwait(&r, &n) != 0) return 0;
  return n;
}

size_t EPLIB_forc_nb(int, const char* filename, const char* mode, void* buffer, size_t size, size_t count, EPLIB_Request* request) {
  if (!request) return 0;
  *request = 0;
  if (!filename || !mode || mode[0] != 'r' || !size) return 0;
  try {
    mlslb::RankContext* c = mlslb::current_context();
    if (!c->initialized) return 0;
    mlslb::IoRequest* r = mlslb::io_open_read_close_nb(c, filename, buffer, size * count, 0);
    *request = (EPLIB_Request)(uintptr_t)r;
    std::lock_guard<std::mutex> g(g_mu);
    g_reqs[*request] = ReqInfo{size, true};
  } catch (const std::exception&) {
  }
  return 0;
}

int EPLIB_fclose(int, EPLIB_FILE stream) {
  if (!stream) return -1;
  int rc = 0;
  try {
    mlslb::io_close(stream->f);
  } catch (const std::exception&) {
    rc = -1;
  }
  delete stream;
  return rc;
}

}  // extern "C"
