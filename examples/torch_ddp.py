"""Unmodified PyTorch DistributedDataParallel + ZeRO-style sharded tensors on the "mlsl" torch.distributed backend.

    bin/mlslrun -n 4 python examples/torch_ddp.py                          # CPU ranks, host backend
    torchrun --nproc-per-node 8 examples/torch_ddp.py --device cuda        # one B200 per rank, CUDA backend

Nothing below mentions the library except the import that registers the backend and the backend name: gradients are
bucketed and all-reduced by DDP, the parameter shards are gathered with all_gather_into_tensor and the gradient shards
reduced with reduce_scatter_tensor - the calls FSDP makes.
"""
import argparse
import os
import sys
import tempfile

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mlsl_b200.torch_backend  # noqa: E402,F401

ap = argparse.ArgumentParser()
ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"])
ap.add_argument("--steps", type=int, default=8)
args = ap.parse_args()

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
dev = torch.device("cpu")
if args.device == "cuda":
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
if "MASTER_ADDR" in os.environ and "MASTER_PORT" in os.environ:
    dist.init_process_group("mlsl", init_method="env://", rank=rank, world_size=world)
else:   # mlslrun: the ranks share a job id, use a file store named after it
    store = os.path.join(tempfile.gettempdir(), "torch_ddp_store_%s" % os.environ.get("MLSL_JOB_ID", "solo"))
    dist.init_process_group("mlsl", init_method="file://" + store, rank=rank, world_size=world)

torch.manual_seed(0)
model = torch.nn.Sequential(torch.nn.Linear(64, 256), torch.nn.GELU(), torch.nn.Linear(256, 64)).to(dev)
import mlsl_b200  # noqa: E402
with mlsl_b200.heap_pool():     # CUDA backend: DDP's gradient buckets are allocated in the symmetric heap -> zero-copy all-reduces
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index] if dev.type == "cuda" else None,
                                                    bucket_cap_mb=0.05)
opt = torch.optim.AdamW(ddp.parameters(), lr=1e-2)
torch.manual_seed(100 + rank)
x, y = torch.randn(32, 64, device=dev), torch.randn(32, 64, device=dev)
first = last = None
for step in range(args.steps):
    opt.zero_grad()
    loss = torch.nn.functional.mse_loss(ddp(x), y)
    loss.backward()
    opt.step()
    mean = loss.detach().clone()
    dist.all_reduce(mean, op=dist.ReduceOp.AVG)
    first = mean.item() if first is None else first
    last = mean.item()

# every replica holds the same parameters
flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
lo, hi = flat.clone(), flat.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
identical = bool(torch.equal(lo, hi))

# the FSDP pair: gather parameter shards, reduce-scatter gradients
pad = (-flat.numel()) % world
padded = torch.cat([flat, flat.new_zeros(pad)])
shard = padded.chunk(world)[rank].clone()
full = torch.empty_like(padded)
dist.all_gather_into_tensor(full, shard)
gshard = torch.empty_like(shard)
dist.reduce_scatter_tensor(gshard, padded * (rank + 1))
sharded_ok = bool(torch.equal(full, padded)) and bool(
    torch.allclose(gshard, padded.chunk(world)[rank] * (world * (world + 1) / 2), rtol=1e-5, atol=1e-6))

ok = identical and sharded_ok and last < first
print("rank %d: loss %.4f -> %.4f, replicas identical: %s, shard round trip: %s : %s"
      % (rank, first, last, identical, sharded_ok, "PASSED" if ok else "FAILED"), flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
