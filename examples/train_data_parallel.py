#!/usr/bin/env python
"""Data-parallel training of a small MLP with mlsl_b200.DistributedOptimizer.

    torchrun --nproc-per-node 8 examples/train_data_parallel.py --mode fused      # one rank per GPU
    bin/mlslrun -n 4 python examples/train_data_parallel.py --steps 5             # CPU, host backend

mode "fused"     : per bucket ONE kernel = reduce-scatter of the gradients + AdamW/SGD on the owned shard + all-gather
                   of the updated parameters (the reference's "distributed update", sharded optimizer state)
mode "allreduce" : gradient all-reduce (optionally fp8-compressed with error feedback), local torch optimizer.
Gradient buckets start their communication from autograd hooks while backward is still running."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsl_b200 as mlsl  # noqa: E402
from mlsl_b200.models.mlp import MLP  # noqa: E402
from mlsl_b200.parallel.data_parallel import DistributedDataParallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--mode", default="fused", choices=["fused", "allreduce"])
    ap.add_argument("--optimizer", default="adamw", choices=["sgd", "adamw"])
    ap.add_argument("--compress", action="store_true")
    ap.add_argument("--width", type=int, default=256)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--lr", type=float, default=None)
    args = ap.parse_args()
    if args.lr is None:
        args.lr = 1e-3 if args.optimizer == "adamw" else 0.05
    use_cuda = torch.cuda.is_available() and os.environ.get("MLSL_BACKEND", "cuda") == "cuda"
    if use_cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    mlsl.init()
    rank, world = mlsl.rank(), mlsl.world_size()
    dev = "cuda" if use_cuda else "cpu"
    torch.manual_seed(1234 + rank)                       # different init per rank: the wrapper broadcasts rank 0's
    model = DistributedDataParallel(MLP(args.width, 4 * args.width, args.width, layers=2).to(dev))
    opt = mlsl.DistributedOptimizer(model.parameters(), lr=args.lr, optimizer=args.optimizer, mode=args.mode,
                                    compress=args.compress, weight_decay=0.01)
    gen = torch.Generator().manual_seed(99 + rank)       # every rank sees its own shard of the data
    target_w = torch.randn(args.width, args.width, generator=torch.Generator().manual_seed(7)) / args.width ** 0.5
    x = torch.randn(args.batch, args.width, generator=gen)        # a fixed batch per rank: the loss must go down
    y = x @ target_w
    x, y = x.to(dev), y.to(dev)
    losses = []
    for step in range(args.steps):
        loss = torch.nn.functional.mse_loss(model(x), y)
        opt.zero_grad()
        loss.backward()                                  # bucket communication starts from the gradient hooks
        opt.step()                                       # waits for it and applies the (sharded) update
        losses.append(float(loss))
        if rank == 0:
            print("step %d loss %.5f" % (step, losses[-1]), flush=True)
    # replicas must still agree bit for bit
    flat = torch.cat([p.detach().float().reshape(-1) for p in model.parameters()]).contiguous()
    ref = flat.clone()
    mlsl.bcast(ref, root=0)
    same = bool(torch.equal(flat, ref))
    print("[%d] replicas identical: %s, loss %.5f -> %.5f" % (rank, same, losses[0], losses[-1]), flush=True)
    opt.close()
    mlsl.finalize()
    return 0 if same and losses[-1] < losses[0] else 1


if __name__ == "__main__":
    sys.exit(main())
