#!/usr/bin/env python
"""Host backend vs the unmodified reference on the same CPUs, collective by collective: the SAME benchmark source
(csrc/tests/mlsl_allreduce_bench.cpp, reference API only) linked against either library.
    python scripts/cpu_vs_reference.py [ranks=4] > profiles/host_backend_vs_reference_ops_cpu.txt"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import ref_bench  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ARGS = ["4096", str(64 << 20), "10", "3", "8"]
OPS = ["allreduce", "allgather", "reducescatter", "alltoall", "bcast"]


def rows(out):
    return [json.loads(l) for l in out.splitlines() if l.strip().startswith("{")]


def ours(op):
    env = dict(os.environ, MLSL_BACKEND="host", MLSL_BENCH_OP=op, MLSL_BENCH_OUT_OF_PLACE="0")
    r = subprocess.run([os.path.join(ROOT, "bin", "mlslrun"), "-n", str(N), os.path.join(ROOT, "bin", "mlsl_allreduce_bench")] + ARGS,
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    return rows(r.stdout)


def reference(op):
    env = ref_bench._env()
    env["MLSL_BENCH_OP"], env["MLSL_BENCH_OUT_OF_PLACE"] = op, "0"
    hydra = os.path.join(ref_bench.REF, "mpirt", "bin", "mpiexec.hydra")
    exe = os.path.join(ref_bench.REF, "bin", "ref_allreduce_bench")
    for extra in ([], ["-hosts", "127.0.0.1", "-localhost", "127.0.0.1"]):
        r = subprocess.run([hydra] + extra + ["-n", str(N), exe] + ARGS, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           text=True, timeout=1800)
        if r.returncode == 0 and "{" in r.stdout:
            return rows(r.stdout)
    raise RuntimeError(r.stderr[-300:])


print("Host backend vs unmodified intel/MLSL (process mode, Intel MPI shm), %d ranks on this machine, fp32, in place where the"
      % N)
print("operation allows it; `bytes` = the larger buffer; microseconds per call, max over ranks; scripts/cpu_vs_reference.py")
for op in OPS:
    a, b = ours(op), reference(op)
    print("\n%s\n%12s %12s %14s %9s" % (op, "bytes", "ours us", "reference us", "speed-up"))
    for x, y in zip(a, b):
        assert x["bytes"] == y["bytes"]
        print("%12d %12.2f %14.2f %8.1fx" % (x["bytes"], x["us"], y["us"], y["us"] / x["us"]))
