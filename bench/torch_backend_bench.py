"""all_reduce / all_gather_into_tensor / reduce_scatter_tensor through torch.distributed: backend "mlsl" against gloo (CPU)
or nccl (CUDA).  Wall-clock per call on CPU (host tensors, nothing asynchronous), CUDA events on GPU; max over ranks.

    bin/mlslrun -n 4 python bench/torch_backend_bench.py --backend mlsl
    bin/mlslrun -n 4 python bench/torch_backend_bench.py --backend gloo
    torchrun --nproc-per-node 8 bench/torch_backend_bench.py --backend mlsl --device cuda     (and --backend nccl)
"""
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
ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"])
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--max-mb", type=int, default=64)
args = ap.parse_args()
if args.backend == "mlsl":
    import mlsl_b200.torch_backend  # noqa: F401

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
dev = torch.device("cpu")
if args.device == "cuda":
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
if "MASTER_ADDR" in os.environ and "MASTER_PORT" in os.environ:
    dist.init_process_group(args.backend, init_method="env://", rank=rank, world_size=world)
else:
    store = os.path.join(tempfile.gettempdir(), "tbb_store_%s_%s" % (args.backend, os.environ.get("MLSL_JOB_ID", "solo")))
    dist.init_process_group(args.backend, init_method="file://" + store, rank=rank, world_size=world)


def timed(fn, iters):
    for _ in range(3):
        fn()
    dist.barrier()
    if dev.type == "cuda":
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
    else:
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        us = (time.perf_counter() - t0) * 1e6 / iters
    t = torch.tensor([us], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


rows = []
size = 4096
while size <= args.max_mb << 20:
    n = size // 4
    x = torch.ones(n, device=dev)
    shard = torch.ones(n // world, device=dev)
    full = torch.empty(n // world * world, device=dev)
    iters = args.iters if size <= 1 << 22 else max(3, args.iters // 4)
    row = {"bytes": size,
           "all_reduce_us": round(timed(lambda: dist.all_reduce(x), iters), 1),
           "all_gather_into_tensor_us": round(timed(lambda: dist.all_gather_into_tensor(full, shard), iters), 1),
           "reduce_scatter_tensor_us": round(timed(lambda: dist.reduce_scatter_tensor(shard, full), iters), 1)}
    rows.append(row)
    x.fill_(1.0)
    size *= 4
if rank == 0:
    print(json.dumps({"backend": args.backend, "device": args.device, "world": world, "rows": rows}))
dist.destroy_process_group()
