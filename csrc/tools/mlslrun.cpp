// mlslrun: single-node process launcher (the role mpiexec.hydra plays for the reference:
// `mpiexec.hydra -n 4 -ppn 1 ./mlsl_test ...`, reference tests/examples/mlsl_test/Makefile:58-106).
//
//   mlslrun -n N [-g] [-e NAME=VALUE]... [--timeout SEC] [--bind core|none] program [args...]
//   mlslrun -n N --nnodes M --node-rank I --master-addr HOST [--master-port P] program [args...]     (run on every node)
//   mlslrun -n N --hosts H0,H1,... [--rsh "ssh -o BatchMode=yes"] program [args...]                   (run once, anywhere)
//
// Starts N copies of `program` with MLSL_RANK / MLSL_WORLD_SIZE / MLSL_LOCAL_RANK and a fresh MLSL_JOB_ID
// (plus the torchrun-style RANK / WORLD_SIZE / LOCAL_RANK), -g additionally pins rank r to GPU r through
// CUDA_VISIBLE_DEVICES.  Like hydra, the launcher binds every rank to its own slice of the cores it may use (rank r
// gets cpus [r*C/N, (r+1)*C/N) when C >= N; --bind none or MLSL_BIND=0 switches that off): ranks spin on each other
// through shared memory, and two of them time-sharing one core cost ~100 us per synchronisation.
// With --nnodes the job spans M nodes: this launcher starts the N local ranks I*N .. I*N+N-1 of a world of M*N, points
// them at rank 0's control server (MLSL_MASTER_ADDR / MLSL_MASTER_PORT) and selects the net (TCP) backend unless
// MLSL_BACKEND says otherwise.
// With --hosts the launcher is the head of the job (hydra's -hosts): it starts `mlslrun --nnodes M --node-rank I ...` on
// every listed host through the remote shell (--rsh / MLSL_RSH, default ssh; the command runs in the current directory,
// MLSL_* variables and every -e are forwarded), H0 hosts the control server, and the first node that fails stops the rest.
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

static std::string sh_quote(const std::string& a) {
  std::string q = "'";
  for (char c : a) {
    if (c == '\'') q += "'\\''";
    else q += c;
  }
  return q + "'";
}

extern char** environ;

// head of a multi-host job: one remote mlslrun per host
static int run_hosts(const std::vector<std::string>& hosts, const std::string& rsh, int n, bool gpus, bool bind, int timeout,
                     std::string master_port, const std::vector<std::string>& envs, char** prog) {
  char self[4096];
  ssize_t len = readlink("/proc/self/exe", self, sizeof(self) - 1);
  if (len <= 0) {
    perror("readlink(/proc/self/exe)");
    return 1;
  }
  self[len] = 0;
  char cwd[4096];
  if (!getcwd(cwd, sizeof(cwd))) cwd[0] = 0;
  if (master_port.empty()) {
    timeval tv;
    gettimeofday(&tv, nullptr);
    master_port = std::to_string(20000 + (int)((tv.tv_usec ^ getpid()) % 12000));     // below the ephemeral range
  }
  if (!getenv("MLSL_JOB_TOKEN")) {          // every node of this job presents the same token when it connects (forwarded below)
    unsigned long long r = 0;
    if (FILE* f = fopen("/dev/urandom", "rb")) {
      if (fread(&r, sizeof(r), 1, f) != 1) r = 0;
      fclose(f);
    }
    timeval tv;
    gettimeofday(&tv, nullptr);
    char tok[40];
    snprintf(tok, sizeof(tok), "%016llx", r ^ ((unsigned long long)tv.tv_usec << 20) ^ (unsigned long long)getpid());
    setenv("MLSL_JOB_TOKEN", tok, 1);
  }
  std::vector<std::string> rsh_argv;   // the remote shell command may carry its own options
  {
    size_t b = 0;
    while (b < rsh.size()) {
      size_t e = rsh.find(' ', b);
      if (e == std::string::npos) e = rsh.size();
      if (e > b) rsh_argv.push_back(rsh.substr(b, e - b));
      b = e + 1;
    }
  }
  g_kids.assign(hosts.size(), -1);
  for (size_t h = 0; h < hosts.size(); ++h) {
    std::string cmd = cwd[0] ? "cd " + sh_quote(cwd) + " && " : std::string();
    cmd += "exec " + sh_quote(self) + " -n " + std::to_string(n) + " --nnodes " + std::to_string(hosts.size()) + " --node-rank " +
           std::to_string(h) + " --master-addr " + sh_quote(hosts[0]) + " --master-port " + master_port;
    if (gpus) cmd += " -g";
    if (!bind) cmd += " --bind none";
    if (timeout > 0) cmd += " --timeout " + std::to_string(timeout);
    for (char** e = environ; *e; ++e)
      if (!strncmp(*e, "MLSL_", 5) && strncmp(*e, "MLSL_RANK=", 10) && strncmp(*e, "MLSL_JOB_ID=", 12)) cmd += " -e " + sh_quote(*e);
    for (const std::string& e : envs) cmd += " -e " + sh_quote(e);
    cmd += " --";
    for (char** a = prog; *a; ++a) cmd += " " + sh_quote(*a);
    pid_t p = fork();
    if (p < 0) {
      perror("fork");
      kill_all(SIGKILL);
      return 1;
    }
    if (p == 0) {
      std::vector<char*> av;
      for (std::string& a : rsh_argv) av.push_back(&a[0]);
      av.push_back(const_cast<char*>(hosts[h].c_str()));
      av.push_back(&cmd[0]);
      av.push_back(nullptr);
      execvp(av[0], av.data());
      fprintf(stderr, "mlslrun: cannot start the remote shell '%s': %s\n", av[0], strerror(errno));
      _exit(127);
    }
    g_kids[h] = p;
  }
  int rc = 0;
  size_t left = hosts.size();
  while (left) {
    int st = 0;
    pid_t p = wait(&st);
    if (p < 0) {
      if (errno == EINTR) continue;
      break;
    }
    for (size_t h = 0; h < g_kids.size(); ++h)
      if (g_kids[h] == p) {
        g_kids[h] = -1;
        --left;
        const int code = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + WTERMSIG(st);
        if (code != 0 && rc == 0) {
          rc = code;
          fprintf(stderr, "mlslrun: node %zu (%s) exited with code %d, stopping the job\n", h, hosts[h].c_str(), code);
          kill_all(SIGTERM);
        }
      }
  }
  return rc;
}

int main(int argc, char** argv) {
  int n = 1, timeout = 0;
  bool gpus = false;
  int nnodes = 1, node_rank = 0;
  std::string master_addr, master_port, hosts_arg, rsh = getenv("MLSL_RSH") ? getenv("MLSL_RSH") : "ssh";
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
    else if (!strcmp(argv[i], "--hosts") && i + 1 < argc) hosts_arg = argv[++i];
    else if (!strcmp(argv[i], "--rsh") && i + 1 < argc) rsh = argv[++i];
    else if (!strcmp(argv[i], "--")) { ++i; break; }
    else break;
  }
  if (i >= argc || n < 1 || nnodes < 1 || node_rank < 0 || node_rank >= nnodes) {
    fprintf(stderr, "usage: mlslrun -n N [-g] [-e NAME=VALUE]... [--timeout SEC] [--bind core|none] program [args...]\n"
                    "       mlslrun -n N --nnodes M --node-rank I --master-addr HOST [--master-port P] program [args...]\n"
                    "       mlslrun -n N --hosts H0,H1,... [--rsh CMD] program [args...]\n");
    return 2;
  }
  if (!hosts_arg.empty()) {
    std::vector<std::string> hosts;
    size_t b = 0;
    while (b <= hosts_arg.size()) {
      size_t e = hosts_arg.find(',', b);
      if (e == std::string::npos) e = hosts_arg.size();
      if (e > b) hosts.push_back(hosts_arg.substr(b, e - b));
      b = e + 1;
    }
    if (hosts.empty()) {
      fprintf(stderr, "mlslrun: --hosts needs a comma separated list of host names\n");
      return 2;
    }
    signal(SIGINT, on_signal);
    signal(SIGTERM, on_signal);
    return run_hosts(hosts, rsh, n, gpus, bind, timeout, master_port, envs, argv + i);
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
        snprintf(buf, sizeof(buf), "%d", node_rank);
        setenv("MLSL_NODE_RANK", buf, 1);          // the net backend: ranks of one node talk through shared memory
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
