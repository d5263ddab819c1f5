// mlslrun: single-node process launcher (the role mpiexec.hydra plays for the reference:
// `mpiexec.hydra -n 4 -ppn 1 ./mlsl_test ...`, reference tests/examples/mlsl_test/Makefile:58-106).
//
//   mlslrun -n N [-g] [-e NAME=VALUE]... [--timeout SEC] [--bind core|none] program [args...]
//   mlslrun -n N --nnodes M --node-rank I --master-addr HOST [--master-port P] program [args...]     (run on every node)
//
// Starts N copies of `program` with MLSL_RANK / MLSL_WORLD_SIZE / MLSL_LOCAL_RANK and a fresh MLSL_JOB_ID
// (plus the torchrun-style RANK / WORLD_SIZE / LOCAL_RANK), -g additionally pins rank r to GPU r through
// CUDA_VISIBLE_DEVICES.  Like hydra, the launcher binds every rank to its own slice of the cores it may use (rank r
// gets cpus [r*C/N, (r+1)*C/N) when C >= N; --bind none or MLSL_BIND=0 switches that off): ranks spin on each other
// through shared memory, and two of them time-sharing one core cost ~100 us per synchronisation.
// With --nnodes the job spans M nodes: this launcher starts the N local ranks I*N .. I*N+N-1 of a world of M*N, points
// them at rank 0's control server (MLSL_MASTER_ADDR / MLSL_MASTER_PORT) and selects the net (TCP) backend unless
// MLSL_BACKEND says otherwise.
// The first non-zero exit (or the timeout) terminates the whole process group and
// becomes the launcher's exit code - fail-fast like the reference's abort-on-assert.
#include <sched.h>
#include <signal.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static std::vector<pid_t> g_kids;

static void kill_all(int sig) {
  for (pid_t p : g_kids)
    if (p > 0) kill(p, sig);
}

static void on_signal(int) {
  kill_all(SIGTERM);
  _exit(130);
}

int main(int argc, char** argv) {
  int n = 1, timeout = 0;
  bool gpus = false;
  int nnodes = 1, node_rank = 0;
  std::string master_addr, master_port;
  bool bind = !(getenv("MLSL_BIND") && atoi(getenv("MLSL_BIND")) == 0);
  std::vector<std::string> envs;
  int i = 1;
  for (; i < argc; ++i) {
    if (!strcmp(argv[i], "-n") && i + 1 < argc) n = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-g")) gpus = true;
    else if (!strcmp(argv[i], "-e") && i + 1 < argc) envs.push_back(argv[++i]);
    else if (!strcmp(argv[i], "--timeout") && i + 1 < argc) timeout = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--nnodes") && i + 1 < argc) nnodes = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--node-rank") && i + 1 < argc) node_rank = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--master-addr") && i + 1 < argc) master_addr = argv[++i];
    else if (!strcmp(argv[i], "--master-port") && i + 1 < argc) master_port = argv[++i];
    else if (!strcmp(argv[i], "--bind") && i + 1 < argc) bind = strcmp(argv[++i], "none") != 0;
    else if (!strcmp(argv[i], "--")) { ++i; break; }
    else break;
  }
  if (i >= argc || n < 1 || nnodes < 1 || node_rank < 0 || node_rank >= nnodes) {
    fprintf(stderr, "usage: mlslrun -n N [-g] [-e NAME=VALUE]... [--timeout SEC] [--bind core|none] program [args...]\n");
    return 2;
  }
  timeval tv;
  gettimeofday(&tv, nullptr);
  char job[64];
  snprintf(job, sizeof(job), "j%d_%ld%06ld", (int)getpid(), (long)tv.tv_sec, (long)tv.tv_usec);
  signal(SIGINT, on_signal);
  signal(SIGTERM, on_signal);
  g_kids.assign(n, -1);
  std::vector<int> cpus;   // what the launcher itself may run on (cgroup / taskset aware)
  {
    cpu_set_t set;
    CPU_ZERO(&set);
    if (bind && sched_getaffinity(0, sizeof(set), &set) == 0)
      for (int c = 0; c < CPU_SETSIZE; ++c)
        if (CPU_ISSET(c, &set)) cpus.push_back(c);
    if ((int)cpus.size() < n) cpus.clear();   // fewer cores than ranks: let the kernel place them
  }
  for (int r = 0; r < n; ++r) {
    pid_t p = fork();
    if (p < 0) {
      perror("fork");
      kill_all(SIGKILL);
      return 1;
    }
    if (p == 0) {
      if (!cpus.empty()) {
        cpu_set_t set;
        CPU_ZERO(&set);
        const size_t lo = (size_t)r * cpus.size() / n, hi = (size_t)(r + 1) * cpus.size() / n;
        for (size_t k = lo; k < hi; ++k) CPU_SET(cpus[k], &set);
        sched_setaffinity(0, sizeof(set), &set);
      }
      char buf[32];
      snprintf(buf, sizeof(buf), "%d", node_rank * n + r);
      setenv("MLSL_RANK", buf, 1);
      setenv("RANK", buf, 1);
      snprintf(buf, sizeof(buf), "%d", r);
      setenv("MLSL_LOCAL_RANK", buf, 1);
      setenv("LOCAL_RANK", buf, 1);
      if (gpus) setenv("CUDA_VISIBLE_DEVICES", buf, 1);
      snprintf(buf, sizeof(buf), "%d", nnodes * n);
      setenv("MLSL_WORLD_SIZE", buf, 1);
      setenv("WORLD_SIZE", buf, 1);
      setenv("MLSL_JOB_ID", job, 1);
      if (nnodes > 1) {
        setenv("MLSL_BACKEND", "net", 0);
        if (!master_addr.empty()) setenv("MLSL_MASTER_ADDR", master_addr.c_str(), 1);
        if (!master_port.empty()) setenv("MLSL_MASTER_PORT", master_port.c_str(), 1);
      }
      for (auto& e : envs) {
        size_t eq = e.find('=');
        if (eq != std::string::npos) setenv(e.substr(0, eq).c_str(), e.substr(eq + 1).c_str(), 1);
      }
      execvp(argv[i], argv + i);
      fprintf(stderr, "mlslrun: cannot exec %s: %s\n", argv[i], strerror(errno));
      _exit(127);
    }
    g_kids[r] = p;
  }
  if (timeout > 0) alarm((unsigned)timeout);
  struct sigaction sa;
  memset(&sa, 0, sizeof(sa));
  sa.sa_handler = [](int) {
    fprintf(stderr, "mlslrun: timeout, killing ranks\n");
    kill_all(SIGKILL);
    _exit(124);
  };
  sigaction(SIGALRM, &sa, nullptr);
  int rc = 0, left = n;
  while (left > 0) {
    int st = 0;
    pid_t p = wait(&st);
    if (p < 0) {
      if (errno == EINTR) continue;
      break;
    }
    int code = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + WTERMSIG(st);
    for (int r = 0; r < n; ++r)
      if (g_kids[r] == p) {
        g_kids[r] = -1;
        if (code != 0 && rc == 0) {
          fprintf(stderr, "mlslrun: rank %d exited with code %d, stopping the job\n", r, code);
          rc = code;
          kill_all(SIGTERM);
        }
      }
    --left;
  }
  return rc;
}
