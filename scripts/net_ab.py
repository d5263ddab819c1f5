#!/usr/bin/env python
"""A/B runs of the net backend on one machine: the same collective under several values of one environment variable, runs
interleaved (loop-back timings drift), best of R per cell.  Nodes = launchers (bin/mlslrun --nnodes), so NODES x PER_NODE ranks.

    python scripts/net_ab.py MLSL_NET_CHUNK_KB 1048576,512 --op allreduce --nodes 4 --per-node 1
    MLSL_NET_EMULATE_GBIT=10 python scripts/net_ab.py MLSL_NET_CHUNK_KB 1048576,512,2048
    python scripts/net_ab.py MLSL_NET_SHM 0,1 --nodes 2 --per-node 2 --lo 4096
(the sources of profiles/r2/net_chunked_reductions_cpu.txt and net_same_node_shared_memory_cpu.txt)"""
import argparse
import json
import os
import random
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "bin", "mlsl_allreduce_bench")
RUN = os.path.join(ROOT, "bin", "mlslrun")


def run(env, a):
    port = str(random.randrange(20000, 32000))          # below the ephemeral range
    cmd = lambda i: [RUN, "-n", str(a.per_node), "--bind", "none", "--nnodes", str(a.nodes), "--node-rank", str(i), "--master-addr",
                     "127.0.0.1", "--master-port", port, "--timeout", "600", EXE, str(a.lo), str(a.hi), str(a.iters), "2", str(a.factor)]
    others = [subprocess.Popen(cmd(i), env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i in range(1, a.nodes)]
    r = subprocess.run(cmd(0), env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    for p in others:
        p.wait(timeout=120)
    return [json.loads(l) for l in r.stdout.splitlines() if l.strip().startswith("{")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("var")
    ap.add_argument("values")
    ap.add_argument("--op", default="allreduce")
    ap.add_argument("--nodes", type=int, default=4)
    ap.add_argument("--per-node", type=int, default=1)
    ap.add_argument("--lo", type=int, default=4 << 20)
    ap.add_argument("--hi", type=int, default=64 << 20)
    ap.add_argument("--factor", type=int, default=4)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--rounds", type=int, default=3)
    a = ap.parse_args()
    values = a.values.split(",")
    acc = {v: {} for v in values}
    for _ in range(a.rounds):
        for v in values:
            env = dict(os.environ, MLSL_BENCH_OP=a.op, MLSL_BENCH_OUT_OF_PLACE="0")
            env.pop("MLSL_BACKEND", None)
            env[a.var] = v
            for row in run(env, a):
                acc[v][row["bytes"]] = min(acc[v].get(row["bytes"], 1e30), row["us"])
    sizes = sorted(acc[values[0]])
    print("%s, %d x %d ranks, us per call (best of %d)\n%-22s %s" % (a.op, a.nodes, a.per_node, a.rounds, a.var, " ".join("%10d" % b for b in sizes)))
    for v in values:
        print("%-22s %s" % (v, " ".join("%10.0f" % acc[v].get(b, float("nan")) for b in sizes)))


if __name__ == "__main__":
    main()
