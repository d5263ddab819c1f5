#include "log.hpp"

#include <execinfo.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>

#include "common.hpp"

namespace mlslb {

static std::atomic<int> g_level{-1};
static std::atomic<int> g_throw{-1};
static std::atomic<int> g_rank{-1};
static std::atomic<void (*)()> g_fail_hook{nullptr};

int log_level() {
  int l = g_level.load(std::memory_order_relaxed);
  if (l < 0) {
    const char* e = getenv("MLSL_LOG_LEVEL");
    l = e ? atoi(e) : 0;
    if (l < 0) l = 0;
    if (l > 3) l = 3;
    g_level.store(l);
  }
  return l;
}
void set_log_level(int lvl) { g_level.store(lvl); }
void set_log_rank(int r) { g_rank.store(r); }

bool assert_throws() {
  int t = g_throw.load(std::memory_order_relaxed);
  if (t < 0) {
    const char* e = getenv("MLSL_ASSERT_MODE");
    t = (e && strcmp(e, "throw") == 0) ? 1 : 0;
    g_throw.store(t);
  }
  return t == 1;
}
void set_assert_throws(bool on) { g_throw.store(on ? 1 : 0); }
void set_fail_hook(void (*hook)()) { g_fail_hook.store(hook); }

static const char* lvl_name(int l) {
  switch (l) {
    case LOG_ERROR: return "ERROR";
    case LOG_INFO: return "INFO";
    case LOG_DEBUG: return "DEBUG";
    default: return "TRACE";
  }
}

void log_emit(int lvl, const char* file, int line, const char* func, const char* fmt, ...) {
  char msg[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(msg, sizeof(msg), fmt, ap);
  va_end(ap);
  const char* base = strrchr(file, '/');
  base = base ? base + 1 : file;
  long tid = syscall(SYS_gettid);
  int rank = g_rank.load();
  if (log_level() >= LOG_DEBUG) {
    timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    fprintf(stdout, "[%ld.%06ld] (r%d t%ld) %s %s:%d %s: %s\n", (long)ts.tv_sec, ts.tv_nsec / 1000, rank, tid,
            lvl_name(lvl), base, line, func, msg);
  } else {
    fprintf(stdout, "(r%d t%ld) %s %s:%d: %s\n", rank, tid, lvl_name(lvl), base, line, msg);
  }
  fflush(stdout);
}

void fail(const char* file, int line, const char* func, const char* cond, const char* fmt, ...) {
  char msg[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(msg, sizeof(msg), fmt, ap);
  va_end(ap);
  const char* base = strrchr(file, '/');
  base = base ? base + 1 : file;
  char full[2600];
  snprintf(full, sizeof(full), "MLSL assertion failed: (%s) at %s:%d %s: %s", cond, base, line, func, msg);
  if (assert_throws()) {
    if (getenv("MLSL_DEBUG_ERRORS")) {
      fprintf(stderr, "(r%d) %s\n", g_rank.load(), full);
      fflush(stderr);
    }
    throw Error(full);
  }
  fprintf(stderr, "(r%d) %s\n", g_rank.load(), full);
  void* bt[32];
  int n = backtrace(bt, 32);
  backtrace_symbols_fd(bt, n, 2);
  fflush(stderr);
  static std::atomic<int> once{0};
  auto hook = g_fail_hook.load();
  if (hook && once.fetch_add(1) == 0) hook();
  _exit(1);
}

uint64_t now_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000000000ull + ts.tv_nsec;
}

uint64_t cycles_now() {
#if defined(__x86_64__)
  unsigned lo, hi;
  __asm__ __volatile__("rdtsc" : "=a"(lo), "=d"(hi));
  return ((uint64_t)hi << 32) | lo;
#else
  return now_ns();
#endif
}

}  // namespace mlslb
