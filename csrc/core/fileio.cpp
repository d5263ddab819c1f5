// File-IO offload: non-blocking reads executed by a background thread, completion through request handles.
//
// The reference's endpoint library offers EPLIB_fopen / fread / fread_nb / forc_nb ("open-read-close") / fwait /
// fwaitall / fclose so that data loading runs on the endpoint servers instead of the compute threads (reference
// eplib/wrapper.c:1009-1097, compiled in with FILEIO=1).  Same service here: requests are queued to one IO thread per
// rank context; the destination may be host memory or - on the CUDA backend - device memory (the chunk is read into a
// bounce buffer and handed to Backend::copy_from_host, i.e. cudaMemcpy on the device path).
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "log.hpp"
#include "runtime.hpp"

namespace mlslb {

struct IoFile {
  int fd = -1;
  RankContext* ctx = nullptr;
  std::string path;
};

struct IoRequest {
  IoFile* file = nullptr;
  bool owns_file = false;         // forc_nb: close after the read
  void* dst = nullptr;
  size_t bytes = 0;
  long long offset = 0;
  std::atomic<int> done{0};
  size_t result = 0;
  int err = 0;
};

class IoService {
 public:
  explicit IoService(RankContext* ctx) : ctx_(ctx), th_([this] { run(); }) {}
  ~IoService() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    if (th_.joinable()) th_.join();
  }
  void submit(IoRequest* r) {
    {
      std::lock_guard<std::mutex> g(mu_);
      q_.push_back(r);
    }
    cv_.notify_one();
  }

 private:
  void run() {
    set_log_rank(ctx_->rank);
    std::vector<char> bounce;
    for (;;) {
      IoRequest* r = nullptr;
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [&] { return stop_ || !q_.empty(); });
        if (q_.empty()) return;
        r = q_.front();
        q_.pop_front();
      }
      const bool device_dst = ctx_->backend && ctx_->backend->is_device() && ctx_->backend->is_device_pointer(r->dst);
      size_t got = 0;
      const size_t chunk = (size_t)8 << 20;
      if (device_dst && bounce.size() < chunk) bounce.resize(chunk);
      while (got < r->bytes) {
        size_t want = std::min(chunk, r->bytes - got);
        char* into = device_dst ? bounce.data() : (char*)r->dst + got;
        ssize_t n = pread(r->file->fd, into, want, (off_t)(r->offset + (long long)got));
        if (n < 0) {
          if (errno == EINTR) continue;
          r->err = errno;
          break;
        }
        if (n == 0) break;   // end of file
        if (device_dst) ctx_->backend->copy_from_host((char*)r->dst + got, into, (size_t)n);
        got += (size_t)n;
      }
      r->result = got;
      if (r->owns_file) {
        close(r->file->fd);
        delete r->file;
        r->file = nullptr;
      }
      r->done.store(1, std::memory_order_release);
    }
  }
  RankContext* ctx_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<IoRequest*> q_;
  bool stop_ = false;
  std::thread th_;
};

static IoService* service_of(RankContext* ctx) {
  if (!ctx->io_service) ctx->io_service = new IoService(ctx);
  return (IoService*)ctx->io_service;
}

void io_shutdown(RankContext* ctx) {
  delete (IoService*)ctx->io_service;
  ctx->io_service = nullptr;
}

IoFile* io_open(RankContext* ctx, const char* path) {
  MLSLB_ASSERT(path != nullptr, "io_open: null path");
  int fd = open(path, O_RDONLY | O_CLOEXEC);
  MLSLB_ASSERT(fd >= 0, "io_open(%s): %s", path, strerror(errno));
  IoFile* f = new IoFile();
  f->fd = fd;
  f->ctx = ctx;
  f->path = path;
  return f;
}

size_t io_size(IoFile* f) {
  struct stat st;
  MLSLB_ASSERT(fstat(f->fd, &st) == 0, "fstat(%s): %s", f->path.c_str(), strerror(errno));
  return (size_t)st.st_size;
}

void io_close(IoFile* f) {
  if (!f) return;
  close(f->fd);
  delete f;
}

IoRequest* io_read_nb(IoFile* f, void* dst, size_t bytes, long long offset) {
  MLSLB_ASSERT(f && dst, "io_read_nb: null file or destination");
  IoRequest* r = new IoRequest();
  r->file = f;
  r->dst = dst;
  r->bytes = bytes;
  r->offset = offset;
  service_of(f->ctx)->submit(r);
  return r;
}

// open + read + close as one non-blocking request (the reference's EPLIB_forc_nb)
IoRequest* io_open_read_close_nb(RankContext* ctx, const char* path, void* dst, size_t bytes, long long offset) {
  IoFile* f = io_open(ctx, path);
  IoRequest* r = new IoRequest();
  r->file = f;
  r->owns_file = true;
  r->dst = dst;
  r->bytes = bytes;
  r->offset = offset;
  service_of(ctx)->submit(r);
  return r;
}

bool io_test(IoRequest* r, size_t* bytes_read) {
  if (!r->done.load(std::memory_order_acquire)) return false;
  if (bytes_read) *bytes_read = r->result;
  return true;
}

size_t io_wait(IoRequest* r) {
  while (!r->done.load(std::memory_order_acquire)) usleep(50);
  size_t n = r->result;
  int err = r->err;
  delete r;
  MLSLB_ASSERT(err == 0, "file read failed: %s", strerror(err));
  return n;
}

}  // namespace mlslb
