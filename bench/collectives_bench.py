"""All-gather / reduce-scatter / all-to-all / broadcast on N GPUs: this library's peer-memory kernels against NCCL, device
timed (CUDA events, max over ranks), symmetric-heap buffers for ours, plain device tensors for NCCL.
    torchrun --nproc-per-node 8 bench/collectives_bench.py [--max-mb 256]
One JSON line per (op, bytes) on rank 0; `bytes` = the larger buffer, busbw = bytes / t x (N-1)/N."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mlsl_b200 as mlsl  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--max-mb", type=int, default=256)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ.setdefault("MLSL_STREAM_MODE", "inline")
mlsl.init()
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    mlsl.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


size = 1 << 20
while size <= args.max_mb << 20:
    n = size // 4 // world * world            # elements of the whole (fp32)
    part = n // world
    whole = mlsl.alloc_tensor(n, torch.float32, zero=True)
    whole2 = mlsl.alloc_tensor(n, torch.float32, zero=True)
    shard = mlsl.alloc_tensor(part, torch.float32, zero=True)
    tw, tw2, ts = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda"), torch.zeros(part, device="cuda")
    iters = args.iters if size <= 64 << 20 else max(5, args.iters // 4)
    cases = {
        "all_gather": (lambda: mlsl.allgather(shard, out=whole), lambda: dist.all_gather_into_tensor(tw, ts)),
        "reduce_scatter": (lambda: mlsl.reduce_scatter(whole, out=shard), lambda: dist.reduce_scatter_tensor(ts, tw)),
        "all_to_all": (lambda: mlsl.alltoall(whole, out=whole2), lambda: dist.all_to_all_single(tw2, tw)),
        "broadcast": (lambda: mlsl.bcast(whole, root=0), lambda: dist.broadcast(tw, src=0)),
    }
    for op, (ours, nccl) in cases.items():
        a, b = timed(ours, iters), timed(nccl, iters)
        if rank == 0:
            f = 1.0 if op == "broadcast" else (world - 1) / world
            print(json.dumps({"op": op, "bytes": n * 4, "world": world, "ours_us": round(a * 1e3, 2), "nccl_us": round(b * 1e3, 2),
                              "ours_busbw_GBps": round(n * 4 / (a * 1e-3) / 1e9 * f, 1),
                              "nccl_busbw_GBps": round(n * 4 / (b * 1e-3) / 1e9 * f, 1), "speedup": round(b / a, 2)}), flush=True)
    for t in (whole, whole2, shard):
        mlsl.free_tensor(t)
    size *= 4
dist.destroy_process_group()
mlsl.finalize()
