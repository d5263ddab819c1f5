"""Randomised sequences of collectives over random data x model factorizations against a plain reference: every op of
the Distribution API (incl. the *v variants, rooted ops and in-place forms), random sizes, dtypes and groups, heap and
foreign buffers mixed.  Deterministic per seed; the same program runs on every rank."""
import random

import pytest
import torch

from conftest import run_ranks

DT = {torch.float32: 0, torch.float64: 1, torch.uint8: 2, torch.int32: 5}


def _data(seed, rank, n, dtype):
    g = torch.Generator().manual_seed(seed * 1000003 + rank * 7919 + n)
    if dtype in (torch.int32, torch.uint8):
        return torch.randint(0, 5, (n,), generator=g, dtype=torch.int32).to(dtype)
    return (torch.rand(n, generator=g, dtype=torch.float64) * 4 - 2).to(dtype)


def _members(world, D, M, rank, group):
    """global ranks of `rank`'s data / model group under the reference's colour math (M consecutive ranks = model group)"""
    block = D * M
    base = rank // block * block
    lid = rank % block
    if group == 1:      # model
        start = base + lid // M * M
        return list(range(start, start + M))
    return [base + q * M + lid % M for q in range(D)]      # data


@pytest.mark.parametrize("world,seed", [(3, 1), (4, 2), (4, 3), (6, 4), (8, 5)])
def test_random_collective_sequences(world, seed):
    run_fuzz(world, seed, "host")


def make_program(world, seed):
    rng = random.Random(seed)
    facts = [(d, m) for d in range(1, world + 1) for m in range(1, world + 1) if d * m == world]
    D, M = rng.choice(facts)
    program = []
    for step in range(28):
        program.append(dict(op=rng.choice(["allreduce", "allreduce_inplace", "reduce_scatter", "allgather", "allgatherv",
                                           "alltoall", "alltoallv", "bcast", "reduce", "gather", "scatter", "barrier"]),
                            n=rng.choice([1, 3, 17, 256, 1000, 4099]), group=rng.choice([0, 1]),
                            dtype=rng.choice(list(DT)), red=rng.choice(["sum", "min", "max"]), root=rng.randrange(64),
                            heap=rng.random() < 0.5, seed=seed * 100 + step,
                            cnts=[rng.randrange(0, 9) for _ in range(64)]))
    return D, M, program


def run_program(r, mlsl, world, D, M, program, dev="cpu"):
    if True:
        e = mlsl.env()
        dist = e.create_distribution(D, M)
        res = []
        for st in program:
            mem = _members(world, D, M, r, st["group"])
            P, idx = len(mem), mem.index(r)
            assert dist.get_process_count(st["group"]) == P and dist.get_process_idx(st["group"]) == idx
            dt, n, op, g = st["dtype"], st["n"], st["op"], st["group"]
            rop = {"sum": 0, "min": 1, "max": 2}[st["red"]]
            root = st["root"] % P

            def buf(t):     # same values, heap or foreign memory
                if not st["heap"]:
                    return t.clone().to(dev)
                h = mlsl.alloc_tensor(max(t.numel(), 1), t.dtype)[:t.numel()]
                h.copy_(t)
                return h

            if op == "barrier":
                dist.barrier(g)
                res.append(None)
            elif op in ("allreduce", "allreduce_inplace"):
                x = buf(_data(st["seed"], r, n, dt))
                y = x if op.endswith("inplace") else buf(torch.zeros(n, dtype=dt))
                e.wait(dist.all_reduce(x, y, n, DT[dt], rop, g))
                res.append(y.clone().cpu())
            elif op == "reduce_scatter":
                x, y = buf(_data(st["seed"], r, n * P, dt)), buf(torch.zeros(n, dtype=dt))
                e.wait(dist.reduce_scatter(x, y, n, DT[dt], rop, g))
                res.append(y.clone().cpu())
            elif op == "allgather":
                x, y = buf(_data(st["seed"], r, n, dt)), buf(torch.zeros(n * P, dtype=dt))
                e.wait(dist.all_gather(x, n, y, DT[dt], g))
                res.append(y.clone().cpu())
            elif op == "allgatherv":
                cnts = [st["cnts"][mem[p] % 64] + 1 for p in range(P)]
                x, y = buf(_data(st["seed"], r, cnts[idx], dt)), buf(torch.zeros(sum(cnts), dtype=dt))
                e.wait(dist.all_gatherv(x, cnts[idx], y, cnts, DT[dt], g))
                res.append(y.clone().cpu())
            elif op == "alltoall":
                x, y = buf(_data(st["seed"], r, n * P, dt)), buf(torch.zeros(n * P, dtype=dt))
                e.wait(dist.all_to_all(x, n, y, DT[dt], g))
                res.append(y.clone().cpu())
            elif op == "alltoallv":
                # count from member a to member b: a function of the global ranks only, so both sides agree
                cnt = lambda a, b: (st["cnts"][(mem[a] * 5 + mem[b]) % 64]) % 7      # noqa: E731
                sc = [cnt(idx, p) for p in range(P)]
                rc = [cnt(p, idx) for p in range(P)]
                so = [sum(sc[:p]) for p in range(P)]
                ro = [sum(rc[:p]) for p in range(P)]
                x, y = buf(_data(st["seed"], r, max(sum(sc), 1), dt)), buf(torch.zeros(max(sum(rc), 1), dtype=dt))
                e.wait(dist.all_to_allv(x, sc, so, y, rc, ro, DT[dt], g))
                res.append(y[:sum(rc)].clone().cpu())
            elif op == "bcast":
                x = buf(_data(st["seed"], r, n, dt))
                e.wait(dist.bcast(x, n, DT[dt], root, g))
                res.append(x.clone().cpu())
            elif op == "reduce":
                x, y = buf(_data(st["seed"], r, n, dt)), buf(torch.zeros(n, dtype=dt))
                e.wait(dist.reduce(x, y, n, DT[dt], rop, root, g))
                res.append(y.clone().cpu() if idx == root else None)
            elif op == "gather":
                x, y = buf(_data(st["seed"], r, n, dt)), buf(torch.zeros(n * P, dtype=dt))
                e.wait(dist.gather(x, n, y, DT[dt], root, g))
                res.append(y.clone().cpu() if idx == root else None)
            elif op == "scatter":
                x, y = buf(_data(st["seed"], r, n * P, dt)), buf(torch.zeros(n, dtype=dt))
                e.wait(dist.scatter(x, y, n, DT[dt], root, g))
                res.append(y.clone().cpu())
        e.delete_distribution(dist)
        return res


def run_fuzz(world, seed, backend):
    D, M, program = make_program(world, seed)
    dev = "cuda" if backend == "cuda" else "cpu"
    env = {"MLSL_HEAP_SIZE_GB": "0.5", "MLSL_WATCHDOG_SEC": "20"} if backend == "cuda" else None
    outs = run_ranks(world, lambda r, mlsl: run_program(r, mlsl, world, D, M, program, dev), backend=backend, env=env,
                     wait_mode="host" if backend == "cuda" else None)
    for r in range(world):
        check_rank(r, world, D, M, program, outs[r])


def check_rank(r, world, D, M, program, results):
    """compare what rank r got with the closed-form expectation (depends on the deterministic inputs only)"""
    def red(ts, how, dt):
        acc = ts[0].double() if dt != torch.uint8 else ts[0].to(torch.int64)
        for t in ts[1:]:
            t = t.double() if dt != torch.uint8 else t.to(torch.int64)
            acc = acc + t if how == "sum" else (torch.minimum(acc, t) if how == "min" else torch.maximum(acc, t))
        if dt == torch.uint8 and how == "sum":
            acc = acc % 256
        return acc

    for k, st in enumerate(program):
        dt, n, op = st["dtype"], st["n"], st["op"]
        tol = {torch.float32: 1e-5, torch.float64: 1e-12}.get(dt, 0)
        if True:
            mem = _members(world, D, M, r, st["group"])
            P, idx = len(mem), mem.index(r)
            root = st["root"] % P
            got = results[k]

            def same(a, b):
                assert a is not None and a.shape == b.shape, (op, k, r)
                assert torch.allclose(a.double(), b.double(), rtol=tol, atol=tol * 8), (op, k, r, dt)

            if op == "barrier":
                continue
            if op in ("allreduce", "allreduce_inplace"):
                same(got, red([_data(st["seed"], q, n, dt) for q in mem], st["red"], dt))
            elif op == "reduce_scatter":
                same(got, red([_data(st["seed"], q, n * P, dt)[idx * n:(idx + 1) * n] for q in mem], st["red"], dt))
            elif op == "allgather":
                same(got, torch.cat([_data(st["seed"], q, n, dt) for q in mem]))
            elif op == "allgatherv":
                cnts = [st["cnts"][mem[p] % 64] + 1 for p in range(P)]
                same(got, torch.cat([_data(st["seed"], mem[p], cnts[p], dt) for p in range(P)]))
            elif op == "alltoall":
                same(got, torch.cat([_data(st["seed"], q, n * P, dt)[idx * n:(idx + 1) * n] for q in mem]))
            elif op == "alltoallv":
                cnt = lambda a, b: (st["cnts"][(mem[a] * 5 + mem[b]) % 64]) % 7      # noqa: E731
                parts = []
                for p in range(P):
                    sc = [cnt(p, q) for q in range(P)]
                    src = _data(st["seed"], mem[p], max(sum(sc), 1), dt)
                    parts.append(src[sum(sc[:idx]):sum(sc[:idx]) + sc[idx]])
                same(got, torch.cat(parts) if parts else torch.zeros(0, dtype=dt))
            elif op == "bcast":
                same(got, _data(st["seed"], mem[root], n, dt))
            elif op == "reduce":
                if idx == root:
                    same(got, red([_data(st["seed"], q, n, dt) for q in mem], st["red"], dt))
            elif op == "gather":
                if idx == root:
                    same(got, torch.cat([_data(st["seed"], q, n, dt) for q in mem]))
            elif op == "scatter":
                same(got, _data(st["seed"], mem[root], n * P, dt)[idx * n:(idx + 1) * n])
