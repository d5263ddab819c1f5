// Fatal-signal hygiene (reference eplib/sig_handler.c:36-81: the client catches SIGSEGV/BUS/ILL/ABRT/INT/TERM, tears
// the endpoint servers down and unlinks /dev/shm files before exiting, so a crashed rank leaves nothing behind).
// Ours has no server processes and unlinks its shared-memory names right after start-up, so what remains to do on a
// fatal signal is to tell the PEERS: the handler stores the poison word in the shared control block (one atomic
// store - async-signal-safe); every spin loop of the other ranks (host waits, bootstrap, and through the host-mapped
// error word the device kernels) sees it and fails fast instead of waiting for the watchdog.  The previous
// disposition is then restored and the signal re-raised, so core dumps / Python's own handlers behave as before.
#include <dirent.h>
#include <execinfo.h>
#include <pthread.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <atomic>

#include "bootstrap.hpp"
#include "runtime.hpp"

namespace mlslb {

namespace {
const int kSignals[] = {SIGSEGV, SIGBUS, SIGILL, SIGABRT, SIGTERM, SIGINT};
constexpr int kNumSignals = (int)(sizeof(kSignals) / sizeof(kSignals[0]));
struct sigaction g_prev[kNumSignals];
bool g_installed[kNumSignals];
std::atomic<BootCtl*> g_ctl{nullptr};
std::atomic<int> g_rank{0};

void on_fatal(int sig) {
  BootCtl* c = g_ctl.load(std::memory_order_relaxed);
  if (c) {
    uint64_t expect = 0;
    c->poison.compare_exchange_strong(expect, (uint64_t)g_rank.load(std::memory_order_relaxed) + 1);
  }
  for (int i = 0; i < kNumSignals; ++i)
    if (kSignals[i] == sig && g_installed[i]) {
      sigaction(sig, &g_prev[i], nullptr);
      g_installed[i] = false;
    }
  raise(sig);
}
}  // namespace

void install_signal_handlers(RankContext* ctx) {
  if (!ctx->boot || ctx->boot->inproc()) return;          // virtual ranks share one process: nothing to protect
  const char* v = getenv("MLSL_SIG_HANDLERS");
  if (v && atoi(v) == 0) return;
  g_ctl.store(ctx->boot->ctl());
  g_rank.store(ctx->rank);
  for (int i = 0; i < kNumSignals; ++i) {
    struct sigaction cur;
    if (sigaction(kSignals[i], nullptr, &cur) != 0) continue;
    // leave SIGINT/SIGTERM alone when the host application (e.g. the Python interpreter) already handles them
    if ((kSignals[i] == SIGINT || kSignals[i] == SIGTERM) && cur.sa_handler != SIG_DFL) continue;
    struct sigaction sa;
    memset(&sa, 0, sizeof(sa));
    sa.sa_handler = on_fatal;
    sigemptyset(&sa.sa_mask);
    if (sigaction(kSignals[i], &sa, &g_prev[i]) == 0) g_installed[i] = true;
  }
}

void remove_signal_handlers() {
  for (int i = 0; i < kNumSignals; ++i)
    if (g_installed[i]) {
      sigaction(kSignals[i], &g_prev[i], nullptr);
      g_installed[i] = false;
    }
  g_ctl.store(nullptr);
}

}  // namespace mlslb

// ---- hang diagnosis: native stacks of every thread of this process ------------------------------------------------
// A collective that never completes usually means one rank's HOST thread is stuck in a call that waits for the device
// while its peers' kernels spin for it (see csrc/tools/probe_blocking.cu).  mlsl_debug_dump_stacks() interrupts every
// thread with SIGUSR2; each prints its own backtrace (glibc backtrace_symbols_fd: async-signal-tolerant, no malloc) to
// stderr, one thread at a time.  Python twin: faulthandler.dump_traceback (same pthread ids in the headers).
namespace mlslb {
namespace dbg {
std::atomic<int> g_dump_turn{0};
void on_dump(int) {
  int expect = 0;
  while (!g_dump_turn.compare_exchange_weak(expect, 1, std::memory_order_acquire)) {
    expect = 0;
    usleep(1000);
  }
  char head[128];
  int n = snprintf(head, sizeof(head), "\n== native stack of thread tid %ld (pthread 0x%016lx) ==\n", (long)syscall(SYS_gettid),
                   (unsigned long)pthread_self());
  if (write(2, head, (size_t)n) < 0) {}
  void* frames[48];
  int depth = backtrace(frames, 48);
  backtrace_symbols_fd(frames, depth, 2);
  g_dump_turn.store(0, std::memory_order_release);
}
}  // namespace dbg
}  // namespace mlslb

extern "C" void mlsl_debug_dump_stacks(void) {
  static bool installed = false;
  if (!installed) {
    void* warm[4];
    backtrace(warm, 4);                    // loads libgcc outside the handler
    struct sigaction sa;
    memset(&sa, 0, sizeof(sa));
    sa.sa_handler = mlslb::dbg::on_dump;
    sa.sa_flags = SA_RESTART;
    sigemptyset(&sa.sa_mask);
    sigaction(SIGUSR2, &sa, nullptr);
    installed = true;
  }
  const long self = (long)syscall(SYS_gettid);
  DIR* d = opendir("/proc/self/task");
  if (!d) return;
  while (struct dirent* e = readdir(d)) {
    long tid = atol(e->d_name);
    if (tid <= 0 || tid == self) continue;
    syscall(SYS_tgkill, (long)getpid(), tid, SIGUSR2);
  }
  closedir(d);
  usleep(300000);                          // let them print before the caller goes on
}

