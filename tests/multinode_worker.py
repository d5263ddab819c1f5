"""One rank of a two-level (node x local rank) job: mlsl_b200.parallel.multinode on the host backend inside each "node"
and gloo between them.  Started with torchrun-style variables by test_multinode_cpu.py."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mlsl_b200 import comm  # noqa: E402
from mlsl_b200.parallel import multinode  # noqa: E402

DEV = torch.device("cpu")
if len(sys.argv) > 1 and sys.argv[1] == "cuda":     # one GPU per rank; CUDA_VISIBLE_DEVICES selects the "node's" GPUs
    DEV = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(DEV)
hc = multinode.init_hybrid()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
assert hc.rank == rank and hc.world_size == world and comm.world_size() == int(os.environ["LOCAL_WORLD_SIZE"])


def inp(r, n):
    return (torch.arange(n, dtype=torch.float32) % 17 + r + 1).to(DEV)


for n in (1, 5, 1000, 4099, 1 << 16):
    t = inp(rank, n)
    hc.allreduce(t)
    assert torch.equal(t, sum(inp(r, n) for r in range(world))), ("allreduce", n)
    t = inp(rank, n)
    hc.allreduce(t, scale=1.0 / world)
    assert torch.allclose(t, sum(inp(r, n) for r in range(world)) / world, rtol=1e-6), ("allreduce scale", n)
    t = inp(rank, n)
    hc.allreduce(t, op="max")
    assert torch.equal(t, inp(world - 1, n)), ("allreduce max", n)
    for root in (0, world - 1, world // 2):
        t = inp(100 + root, n) if rank == root else torch.zeros(n, device=DEV)
        hc.bcast(t, root=root)
        assert torch.equal(t, inp(100 + root, n)), ("bcast", n, root)
    out = hc.allgather(inp(rank, n))
    assert torch.equal(out, torch.cat([inp(r, n) for r in range(world)])), ("allgather", n)
hc.barrier()

# data-parallel SGD over both levels against a single-process model on the whole batch
torch.manual_seed(3)
model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 2)).to(DEV)
ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 2)).to(DEV)
ref.load_state_dict(model.state_dict())
opt, ropt = torch.optim.SGD(model.parameters(), lr=0.1), torch.optim.SGD(ref.parameters(), lr=0.1)
for step in range(3):
    torch.manual_seed(50 + step)
    x, y = torch.randn(world, 4, 8).to(DEV), torch.randn(world, 4, 2).to(DEV)
    opt.zero_grad()
    torch.nn.functional.mse_loss(model(x[rank]), y[rank]).backward()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    hc.allreduce(flat, scale=1.0 / world)
    off = 0
    for p in model.parameters():
        p.grad.copy_(flat[off:off + p.numel()].view_as(p))
        off += p.numel()
    opt.step()
    ropt.zero_grad()
    sum(torch.nn.functional.mse_loss(ref(x[r]), y[r]) for r in range(world)).div(world).backward()
    ropt.step()
for p, q in zip(model.parameters(), ref.parameters()):
    assert torch.allclose(p, q, atol=1e-5 if DEV.type == "cpu" else 1e-3, rtol=1e-4 if DEV.type == "cpu" else 1e-2), "two-level data parallel training diverged"

# the packaged form: DistributedOptimizer with bucketed gradients in the symmetric heap, hooks, two-level all-reduce
import mlsl_b200 as mlsl  # noqa: E402
torch.manual_seed(8)
m2 = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 2)).to(DEV)
r2 = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 2)).to(DEV)
r2.load_state_dict(m2.state_dict())
dopt = mlsl.DistributedOptimizer(m2.parameters(), lr=0.05, momentum=0.9, hybrid=hc, bucket_mb=0.0005)
ropt2 = torch.optim.SGD(r2.parameters(), lr=0.05, momentum=0.9)
for step in range(4):
    torch.manual_seed(90 + step)
    x, y = torch.randn(world, 4, 8).to(DEV), torch.randn(world, 4, 2).to(DEV)
    dopt.zero_grad()
    torch.nn.functional.mse_loss(m2(x[rank]), y[rank]).backward()
    dopt.step()
    ropt2.zero_grad()
    sum(torch.nn.functional.mse_loss(r2(x[r]), y[r]) for r in range(world)).div(world).backward()
    ropt2.step()
for p, q in zip(m2.parameters(), r2.parameters()):
    assert torch.allclose(p, q, atol=1e-5 if DEV.type == "cpu" else 1e-3, rtol=1e-4 if DEV.type == "cpu" else 1e-2), \
        "DistributedOptimizer(hybrid=...) diverged"
assert len(dopt.buckets) > 1
dopt.close()
hc.barrier()
dist.destroy_process_group()
hc.finalize()
print("multinode OK %d" % rank, flush=True)
