"""DistributedDataParallel step time on CPUs: torch.distributed backend "mlsl" (host backend) against gloo.
    bin/mlslrun -n 4 --bind none python bench/torch_ddp_cpu_bench.py --backend mlsl     (and --backend gloo)"""
import argparse
import json
import os
import sys
import tempfile
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
ap = argparse.ArgumentParser()
ap.add_argument("--backend", default="mlsl")
ap.add_argument("--width", type=int, default=1024)
ap.add_argument("--layers", type=int, default=4)
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
if args.backend == "mlsl":
    import mlsl_b200.torch_backend  # noqa: F401
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))
store = os.path.join(tempfile.gettempdir(), "ddpb_%s_%s" % (args.backend, os.environ.get("MLSL_JOB_ID", "solo")))
dist.init_process_group(args.backend, init_method="file://" + store, rank=rank, world_size=world)
torch.manual_seed(0)
layers = []
for _ in range(args.layers):
    layers += [torch.nn.Linear(args.width, args.width), torch.nn.ReLU()]
model = torch.nn.parallel.DistributedDataParallel(torch.nn.Sequential(*layers))
opt = torch.optim.SGD(model.parameters(), lr=0.01)
x, y = torch.randn(32, args.width), torch.randn(32, args.width)


def step():
    opt.zero_grad()
    torch.nn.functional.mse_loss(model(x), y).backward()
    opt.step()


for _ in range(5):
    step()
dist.barrier()
t0 = time.perf_counter()
for _ in range(args.steps):
    step()
ms = (time.perf_counter() - t0) * 1e3 / args.steps
t = torch.tensor([ms], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    n = sum(p.numel() for p in model.parameters())
    print(json.dumps({"backend": args.backend, "world": world, "params": n, "grad_MiB": round(n * 4 / 2**20, 1),
                      "ms_per_step": round(t.item(), 2)}))
dist.destroy_process_group()
