#!/usr/bin/env python
"""Hybrid data x tensor (model) parallel training of a two-layer MLP.

    torchrun --nproc-per-node 8 examples/train_tensor_parallel.py --model-parts 4     # 2 replicas x 4-way TP
    bin/mlslrun -n 4 python examples/train_tensor_parallel.py --model-parts 2         # CPU, host backend

Distribution(data_parts, model_parts) builds the groups exactly like the reference (model_parts consecutive ranks form a
model group, equal positions of different model groups form a data group).  Inside a model group the hidden dimension
is split: ColumnParallelLinear -> ReLU -> RowParallelLinear, whose output is reduce-scattered over the token rows - on
the CUDA backend that GEMM and its reduce-scatter are one tcgen05 kernel (mlsl_b200.ops.gemm_reduce_scatter).  Across
the data group the weight gradients are averaged with Distribution all-reduces."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsl_b200 as mlsl  # noqa: E402
from mlsl_b200.parallel.tensor_parallel import ColumnParallelLinear, RowParallelLinear, gather_rows  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-parts", type=int, default=2)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--tokens", type=int, default=256)
    ap.add_argument("--width", type=int, default=128)
    ap.add_argument("--hidden", type=int, default=512)
    args = ap.parse_args()
    use_cuda = torch.cuda.is_available() and os.environ.get("MLSL_BACKEND", "cuda") == "cuda"
    if use_cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    env = mlsl.init()
    rank, world = mlsl.rank(), mlsl.world_size()
    M = args.model_parts
    assert world % M == 0, "world size must be a multiple of --model-parts"
    dist = env.create_distribution(world // M, M)
    d_idx, m_idx = dist.get_process_idx(0), dist.get_process_idx(1)      # GroupType.DATA, GroupType.MODEL
    dev = "cuda" if use_cuda else "cpu"
    dtype = torch.bfloat16 if use_cuda else torch.float32
    torch.manual_seed(100 + m_idx)                       # same shard on every replica, different shards inside a group
    fc1 = ColumnParallelLinear(args.width, args.hidden, bias=False, distribution=dist, dtype=dtype, device=dev)
    fc2 = RowParallelLinear(args.hidden, args.width, bias=False, distribution=dist, dtype=dtype, device=dev)
    params = [fc1.weight, fc2.weight]
    gen = torch.Generator().manual_seed(7 + d_idx)       # one data shard per replica, identical inside a model group
    target = torch.randn(args.width, args.width, generator=torch.Generator().manual_seed(3)) / args.width ** 0.5
    x = torch.randn(args.tokens, args.width, generator=gen)               # a fixed batch: the loss must go down
    y = (x @ target).to(dev)
    x = x.to(dev).to(dtype)
    losses = []
    for step in range(args.steps):
        out = gather_rows(fc2(torch.relu(fc1(x))), distribution=dist)    # [tokens, width] on every rank of the group
        loss = torch.nn.functional.mse_loss(out.float(), y)
        for p in params:
            p.grad = None
        loss.backward()
        for p in params:                                 # average the shard's gradient over the replicas
            g = p.grad.float().contiguous().view(-1)
            mlsl.allreduce(g, group="data", distribution=dist, scale=1.0 / (world // M))
            p.data.add_(g.view_as(p).to(p.dtype), alpha=-0.5)
        losses.append(float(loss))
        if rank == 0:
            print("step %d loss %.5f" % (step, losses[-1]), flush=True)
    ok = losses[-1] < losses[0]
    print("[%d] data idx %d model idx %d: loss %.5f -> %.5f %s" % (rank, d_idx, m_idx, losses[0], losses[-1],
                                                                "PASSED" if ok else "FAILED"), flush=True)
    env.delete_distribution(dist)
    mlsl.finalize()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
