#!/usr/bin/env python
"""Net backend (TCP between "nodes") vs the unmodified reference over Intel MPI's TCP fabric, same machine, same harness
source: 4 ranks, ours as 4 single-rank nodes on loop-back (bin/mlslrun --nnodes 4), the reference with I_MPI_FABRICS=tcp.
Best of RUNS runs per size (loop-back TCP timings are noisy).
    python scripts/net_vs_reference.py > profiles/net_backend_vs_reference_tcp_cpu.txt"""
import json
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import ref_bench  # noqa: E402

N, RUNS = 4, 3
ARGS = ["4096", str(64 << 20), "20", "3", "8"]
OPS = ["allreduce", "allgather", "reducescatter", "alltoall", "bcast"]
EXE = os.path.join(ROOT, "bin", "mlsl_allreduce_bench")
RUN = os.path.join(ROOT, "bin", "mlslrun")


def rows(out):
    return [json.loads(l) for l in out.splitlines() if l.strip().startswith("{")]


def ours(op):
    port = str(random.randrange(20000, 32000))       # below the ephemeral range
    env = dict(os.environ, MLSL_BENCH_OP=op, MLSL_BENCH_OUT_OF_PLACE="0")
    env.pop("MLSL_BACKEND", None)
    cmd = lambda i: [RUN, "-n", "1", "--bind", "none", "--nnodes", str(N), "--node-rank", str(i), "--master-addr", "127.0.0.1",
                     "--master-port", port, "--timeout", "600", EXE] + ARGS
    others = [subprocess.Popen(cmd(i), env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i in range(1, N)]
    r = subprocess.run(cmd(0), env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    for p in others:
        p.wait(timeout=60)
    return rows(r.stdout)


def reference(op):
    env = ref_bench._env()
    env.update(MLSL_BENCH_OP=op, MLSL_BENCH_OUT_OF_PLACE="0", I_MPI_FABRICS="tcp")
    hydra = os.path.join(ref_bench.REF, "mpirt", "bin", "mpiexec.hydra")
    exe = os.path.join(ref_bench.REF, "bin", "ref_allreduce_bench")
    for extra in ([], ["-hosts", "127.0.0.1", "-localhost", "127.0.0.1"]):
        r = subprocess.run([hydra] + extra + ["-n", str(N), exe] + ARGS, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           text=True, timeout=1800)
        if r.returncode == 0 and "{" in r.stdout:
            return rows(r.stdout)
    raise RuntimeError(r.stderr[-300:])


def best(fn, op):
    acc = {}
    for _ in range(RUNS):
        for r in fn(op):
            acc[r["bytes"]] = min(acc.get(r["bytes"], 1e30), r["us"])
    return acc


print("Net backend vs unmodified intel/MLSL over Intel MPI's TCP fabric (I_MPI_FABRICS=tcp): 4 ranks on this machine, loop-back,")
print("fp32, `bytes` = the larger buffer, microseconds per call (max over ranks, best of %d runs); scripts/net_vs_reference.py" % RUNS)
for op in OPS:
    a, b = best(ours, op), best(reference, op)
    print("\n%s\n%12s %12s %14s %9s" % (op, "bytes", "ours us", "reference us", "speed-up"))
    for k in sorted(a):
        print("%12d %12.1f %14.1f %8.2fx" % (k, a[k], b[k], b[k] / a[k]))
