"""One rank of the torch.distributed "mlsl" backend checks; started by bin/mlslrun from test_torch_backend_cpu.py.
Every collective is compared with a value computed locally from the known per-rank inputs."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mlsl_b200.torch_backend  # noqa: E402,F401  (registers the backend)

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
DEV = torch.device("cpu")
if len(sys.argv) > 2 and sys.argv[2] == "cuda":   # one GPU per rank: torchrun --nproc-per-node N ... env cuda
    DEV = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(DEV)
dist.init_process_group("mlsl", init_method="env://" if sys.argv[1] == "env" else "file://" + sys.argv[1], rank=rank,
                        world_size=world)
assert dist.get_backend() == "mlsl" and dist.get_rank() == rank and dist.get_world_size() == world


def inp(r, n=1000, dtype=torch.float32):
    return (torch.arange(n, dtype=torch.float64) % 13 + r + 1).to(dtype).to(DEV)


def check(name, got, want):
    if not torch.equal(got.to(torch.float64), want.to(torch.float64)):
        print("rank %d: %s mismatch: %s vs %s" % (rank, name, got.flatten()[:6], want.flatten()[:6]), flush=True)
        sys.exit(1)


def collectives(group, ranks):
    P, me = len(ranks), ranks.index(rank)
    for dtype in (torch.float32, torch.bfloat16, torch.int32, torch.int64, torch.float64):
        t = inp(me, dtype=dtype)
        dist.all_reduce(t, group=group)
        check("all_reduce %s" % dtype, t, sum(inp(r, dtype=dtype).double() for r in range(P)).to(dtype))
    for op, fn in ((dist.ReduceOp.MAX, torch.maximum), (dist.ReduceOp.MIN, torch.minimum), (dist.ReduceOp.PRODUCT, torch.mul)):
        t = inp(me, 64) / 4
        dist.all_reduce(t, op=op, group=group)
        want = inp(0, 64) / 4
        for r in range(1, P):
            want = fn(want, inp(r, 64) / 4)
        check("all_reduce %s" % op, t, want)
    t = inp(me)
    dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
    check("all_reduce avg", t, sum(inp(r) for r in range(P)) / P)
    # asynchronous handle + non-contiguous tensor
    base = torch.stack([inp(me, 64), inp(me, 64) * 2], dim=1)
    col = base[:, 1]
    w = dist.all_reduce(col, group=group, async_op=True)
    while not w.is_completed():      # polling completes the request and runs the copy-back of the strided view
        pass
    w.wait()
    check("all_reduce strided", base[:, 1], 2 * sum(inp(r, 64) for r in range(P)))
    check("all_reduce strided untouched column", base[:, 0], inp(me, 64))
    # broadcast (any dtype travels as bytes)
    for root in range(P):
        b = inp(me, 33, torch.int64) if me != root else inp(100 + root, 33, torch.int64)
        dist.broadcast(b, src=ranks[root], group=group)
        check("broadcast", b, inp(100 + root, 33, torch.int64))
    flag = torch.tensor([me == 0], device=DEV)
    dist.broadcast(flag, src=ranks[0], group=group)
    assert bool(flag[0])
    # all_gather (list and flat), reduce_scatter (list and flat)
    outs = [torch.empty(50, device=DEV) for _ in range(P)]
    dist.all_gather(outs, inp(me, 50), group=group)
    for r in range(P):
        check("all_gather", outs[r], inp(r, 50))
    flat = torch.empty(P * 50, dtype=torch.int64, device=DEV)
    dist.all_gather_into_tensor(flat, inp(me, 50, torch.int64), group=group)
    check("all_gather_into_tensor", flat, torch.cat([inp(r, 50, torch.int64) for r in range(P)]))
    shard = torch.empty(40, device=DEV)
    dist.reduce_scatter_tensor(shard, inp(me, 40 * P), group=group)
    check("reduce_scatter_tensor", shard, sum(inp(r, 40 * P) for r in range(P))[me * 40:(me + 1) * 40])
    shard = torch.empty(40, device=DEV)
    dist.reduce_scatter(shard, list(inp(me, 40 * P).chunk(P)), op=dist.ReduceOp.AVG, group=group)
    check("reduce_scatter avg", shard, (sum(inp(r, 40 * P) for r in range(P)) / P)[me * 40:(me + 1) * 40])
    # reduce, gather, scatter
    for root in range(P):
        t = inp(me, 77)
        dist.reduce(t, dst=ranks[root], group=group)
        check("reduce", t, sum(inp(r, 77) for r in range(P)) if me == root else inp(me, 77))
        got = [torch.empty(9, device=DEV) for _ in range(P)] if me == root else None
        dist.gather(inp(me, 9), got, dst=ranks[root], group=group)
        if me == root:
            for r in range(P):
                check("gather", got[r], inp(r, 9))
        out = torch.empty(11, device=DEV)
        dist.scatter(out, [inp(10 * r + root, 11) for r in range(P)] if me == root else None, src=ranks[root], group=group)
        check("scatter", out, inp(10 * me + root, 11))
    # all_to_all: equal and unequal splits, tensor lists
    src = torch.cat([inp(me * P + r, 8) for r in range(P)])
    dst = torch.empty_like(src)
    dist.all_to_all_single(dst, src, group=group)
    check("all_to_all_single", dst, torch.cat([inp(r * P + me, 8) for r in range(P)]))
    in_splits = [(me + r) % 3 + 1 for r in range(P)]
    out_splits = [(r + me) % 3 + 1 for r in range(P)]
    src = torch.cat([torch.full((in_splits[r], 4), float(me * 100 + r), device=DEV) for r in range(P)])
    dst = torch.empty(sum(out_splits), 4, device=DEV)
    dist.all_to_all_single(dst, src, out_splits, in_splits, group=group)
    check("all_to_all_single splits", dst, torch.cat([torch.full((out_splits[r], 4), float(r * 100 + me), device=DEV) for r in range(P)]))
    outs = [torch.empty(5, dtype=torch.int64, device=DEV) for _ in range(P)]
    dist.all_to_all(outs, [inp(me * P + r, 5, torch.int64) for r in range(P)], group=group)
    for r in range(P):
        check("all_to_all", outs[r], inp(r * P + me, 5, torch.int64))
    dist.barrier(group=group)
    objs = [None] * P
    dist.all_gather_object(objs, {"rank": me}, group=group)
    assert [o["rank"] for o in objs] == list(range(P))


collectives(None, list(range(world)))
if world >= 4:
    # overlapping sub-groups created by their members only; every rank calls new_group for every group
    lists = [[0, 1], [2, 3], [1, 2, 3], [0, 3]]
    groups = [dist.new_group(l) for l in lists]
    for l, g in zip(lists, groups):
        if rank in l:
            collectives(g, l)
    # world traffic still works with the sub-groups alive, and a sub-group can be destroyed on its own
    t = torch.ones(10, device=DEV)
    dist.all_reduce(t)
    check("world after groups", t, torch.full((10,), float(world), device=DEV))
    if rank in lists[0]:
        dist.destroy_process_group(groups[0])

# point to point: a send and its recv are one operation of a two-member distribution made on first use
if world > 1:
    up, down = (rank + 1) % world, (rank - 1) % world
    for dtype, n in ((torch.float32, 1000), (torch.int64, 77), (torch.bfloat16, 4099)):
        out, got = inp(rank, n, dtype), torch.zeros(n, dtype=dtype, device=DEV)
        if rank % 2 == 0:                       # even ranks talk first: every message is a rendezvous of its two ranks
            dist.send(out, dst=up)
            dist.recv(got, src=down)
        else:
            dist.recv(got, src=down)
            dist.send(out, dst=up)
        check("send/recv ring %s" % dtype, got, inp(down, n, dtype))
    # both directions of a pair in one batch (pipeline schedules): the couple maps to the same operation on both sides
    if world % 2 == 0:
        mate = rank ^ 1
        a, b = inp(rank, 513), torch.zeros(513, device=DEV)
        base = torch.zeros(64, 2, device=DEV)
        ops = [dist.P2POp(dist.isend, a, mate), dist.P2POp(dist.irecv, b, mate)]
        if rank > mate:
            ops.reverse()                       # the higher rank lists its recv first: pairs match in program order
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        check("batch_isend_irecv", b, inp(mate, 513))
        if rank < mate:
            dist.send(inp(rank, 64), dst=mate)
        else:
            dist.recv(base[:, 1], src=mate)     # strided destination
            check("recv into a strided view", base[:, 1], inp(mate, 64))
            check("recv strided untouched column", base[:, 0], torch.zeros(64, device=DEV))
    dist.barrier()

# DistributedDataParallel end to end: parameters broadcast from rank 0, gradients averaged
torch.manual_seed(1234 + rank)
model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4)).to(DEV)
ddp = torch.nn.parallel.DistributedDataParallel(model, bucket_cap_mb=0.001, device_ids=[DEV.index] if DEV.type == "cuda" else None)
torch.manual_seed(99)
ref = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4)).to(DEV)
state = [p.detach().clone() for p in model.parameters()]
dist.broadcast_object_list(obj := [[s.tolist() for s in state] if rank == 0 else None], src=0)
with torch.no_grad():
    for p, s in zip(ref.parameters(), obj[0]):
        p.copy_(torch.tensor(s, device=DEV))
for p, q in zip(model.parameters(), ref.parameters()):
    check("ddp parameter broadcast", p.detach(), q.detach())
opt = torch.optim.SGD(ddp.parameters(), lr=0.1)
ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
for step in range(3):
    torch.manual_seed(7 * step)
    x_all, y_all = torch.randn(world, 8, 16).to(DEV), torch.randn(world, 8, 4).to(DEV)
    opt.zero_grad()
    torch.nn.functional.mse_loss(ddp(x_all[rank]), y_all[rank]).backward()
    opt.step()
    ropt.zero_grad()
    sum(torch.nn.functional.mse_loss(ref(x_all[r]), y_all[r]) for r in range(world)).div(world).backward()
    ropt.step()
for p, q in zip(model.parameters(), ref.parameters()):
    if not torch.allclose(p, q, atol=1e-5 if DEV.type == "cpu" else 1e-3, rtol=1e-4 if DEV.type == "cpu" else 1e-2):
        print("rank %d: DDP diverged from the single-process reference" % rank, flush=True)
        sys.exit(1)

# the same training with the quantised all-reduce as DDP communication hook: replicas stay identical, the result is close
torch.manual_seed(4321)
qmodel = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4)).to(DEV)
qddp = torch.nn.parallel.DistributedDataParallel(qmodel, device_ids=[DEV.index] if DEV.type == "cuda" else None)
qddp.register_comm_hook(None, mlsl_b200.torch_backend.compressed_allreduce_hook)
qref = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4)).to(DEV)
qref.load_state_dict(qmodel.state_dict())
qopt, qropt = torch.optim.SGD(qddp.parameters(), lr=0.1), torch.optim.SGD(qref.parameters(), lr=0.1)
for step in range(3):
    torch.manual_seed(11 * step)
    x_all, y_all = torch.randn(world, 8, 16).to(DEV), torch.randn(world, 8, 4).to(DEV)
    qopt.zero_grad()
    torch.nn.functional.mse_loss(qddp(x_all[rank]), y_all[rank]).backward()
    qopt.step()
    qropt.zero_grad()
    sum(torch.nn.functional.mse_loss(qref(x_all[r]), y_all[r]) for r in range(world)).div(world).backward()
    qropt.step()
flat = torch.cat([p.detach().reshape(-1) for p in qmodel.parameters()])
want = torch.cat([p.detach().reshape(-1) for p in qref.parameters()])
lo, hi = flat.clone(), flat.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
err = ((flat - want).norm() / want.norm()).item()
if os.environ.get("MLSL_TEST_VERBOSE") and rank == 0:
    print("quantised hook relative error %g" % err, flush=True)
if not torch.equal(lo, hi) or not err < 2e-2:
    print("rank %d: quantised DDP hook: replicas identical %s, relative error %g" % (rank, torch.equal(lo, hi), err), flush=True)
    sys.exit(1)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("torch backend OK", flush=True)
