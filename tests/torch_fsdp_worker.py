"""PyTorch FSDP2 (`fully_shard`), hybrid sharding on a 2-D DeviceMesh, DTensor tensor parallelism and functional
collectives over the "mlsl" torch.distributed backend, each compared with a single-process model fed the whole batch.  Started by bin/mlslrun from test_torch_backend_cpu.py."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mlsl_b200.torch_backend  # noqa: E402,F401

from torch.distributed.device_mesh import init_device_mesh  # noqa: E402
from torch.distributed.fsdp import fully_shard  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("mlsl", init_method="file://" + sys.argv[1], rank=rank, world_size=world)


def train(mesh, tag):
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.ReLU(), torch.nn.Linear(64, 16))
    ref = torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.ReLU(), torch.nn.Linear(64, 16))
    ref.load_state_dict(model.state_dict())
    for layer in model:
        if isinstance(layer, torch.nn.Linear):
            fully_shard(layer, mesh=mesh)
    fully_shard(model, mesh=mesh)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
    for step in range(3):
        torch.manual_seed(step)
        x, y = torch.randn(world, 4, 16), torch.randn(world, 4, 16)
        opt.zero_grad()
        torch.nn.functional.mse_loss(model(x[rank]), y[rank]).backward()
        opt.step()
        ropt.zero_grad()
        sum(torch.nn.functional.mse_loss(ref(x[r]), y[r]) for r in range(world)).div(world).backward()
        ropt.step()
    full = {k: v.full_tensor() if hasattr(v, "full_tensor") else v for k, v in model.state_dict().items()}
    for k, v in ref.state_dict().items():
        if not torch.allclose(full[k], v, atol=1e-5):
            print("rank %d: %s: %s differs by %g" % (rank, tag, k, (full[k] - v).abs().max()), flush=True)
            sys.exit(1)


train(init_device_mesh("cpu", (world,)), "fsdp")
if world % 2 == 0 and world > 2:
    train(init_device_mesh("cpu", (2, world // 2), mesh_dim_names=("replicate", "shard")), "hsdp")


def tensor_parallel(mesh):
    from torch.distributed.tensor.parallel import ColwiseParallel, RowwiseParallel, parallelize_module

    class MLP(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.up, self.down = torch.nn.Linear(16, 64), torch.nn.Linear(64, 16)

        def forward(self, x):
            return self.down(torch.relu(self.up(x)))

    torch.manual_seed(0)
    model, ref = MLP(), MLP()
    ref.load_state_dict(model.state_dict())
    parallelize_module(model, mesh, {"up": ColwiseParallel(), "down": RowwiseParallel()})
    torch.manual_seed(5)
    x = torch.randn(8, 16)
    out, want = model(x), ref(x)
    out.sum().backward()
    want.sum().backward()
    ok = torch.allclose(out, want, atol=1e-5) and torch.allclose(model.up.weight.grad.full_tensor(), ref.up.weight.grad,
                                                                  atol=1e-5)
    import torch.distributed._functional_collectives as fc
    t = torch.as_tensor(fc.all_reduce(torch.ones(4) * (rank + 1), "sum", dist.group.WORLD))
    ok = ok and torch.equal(t, torch.full((4,), world * (world + 1) / 2))
    if not ok:
        print("rank %d: tensor parallel / functional collectives mismatch" % rank, flush=True)
        sys.exit(1)


tensor_parallel(init_device_mesh("cpu", (world,)))
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("torch fsdp OK", flush=True)
