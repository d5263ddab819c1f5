#!/usr/bin/env python
"""Data x tensor/sequence parallel training of a small transformer (mlsl_b200.models.gpt).

    torchrun --nproc-per-node 8 examples/train_parallel_transformer.py --model-parts 4      # 2 replicas x 4-way TP + SP
    bin/mlslrun -n 4 python examples/train_parallel_transformer.py --model-parts 2          # CPU, host backend

Inside a model group the activations are split over the token rows between the blocks and over heads / hidden units inside
them; every exchange is an all-gather in front of a GEMM or a reduce-scatter behind one (on the CUDA backend: the fused
kernels).  Gradients of the sharded weights are averaged over the data group; the replicated ones (LayerNorm, output
biases) are partial sums over the token shards and are first summed over the model group."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsl_b200 as mlsl  # noqa: E402
from mlsl_b200.models.gpt import ParallelTransformer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-parts", type=int, default=2)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--width", type=int, default=64)
    ap.add_argument("--heads", type=int, default=4)
    ap.add_argument("--tokens", type=int, default=64)
    ap.add_argument("--steps", type=int, default=8)
    args = ap.parse_args()
    use_cuda = torch.cuda.is_available() and os.environ.get("MLSL_BACKEND", "cuda") == "cuda"
    if use_cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    env = mlsl.init()
    world, M = mlsl.world_size(), args.model_parts
    assert world % M == 0 and args.tokens % M == 0
    replicas = world // M
    dist = env.create_distribution(replicas, M)
    rep, part = dist.get_process_idx(mlsl.GroupType.DATA), dist.get_process_idx(mlsl.GroupType.MODEL)
    dev = "cuda" if use_cuda else "cpu"
    torch.manual_seed(100 + part)                    # the same shard on every replica
    model = ParallelTransformer(args.layers, args.width, args.heads, distribution=dist, group="model", device=dev)
    with torch.no_grad():                            # replicated parameters start equal everywhere
        for p in model.replicated_parameters():
            mlsl.bcast(p.data, root=0, group="global", distribution=dist)
    replicated = {id(p) for p in model.replicated_parameters()}
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    gen = torch.Generator().manual_seed(7 + rep)     # one sequence per replica, a fixed batch: the loss must fall
    rows = args.tokens // M
    x_full = torch.randn(args.tokens, args.width, generator=gen)
    y_full = torch.tanh(x_full.roll(1, dims=0))      # predict the previous token's features
    x = x_full[part * rows:(part + 1) * rows].to(dev)
    y = y_full[part * rows:(part + 1) * rows].to(dev)
    losses = []
    for step in range(args.steps):
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(model(x), y, reduction="sum") / (args.tokens * args.width)
        loss.backward()
        for p in model.parameters():
            g = p.grad.contiguous().view(-1)
            if id(p) in replicated:
                mlsl.allreduce(g, group="model", distribution=dist)
            mlsl.allreduce(g, group="data", distribution=dist, scale=1.0 / replicas)
            p.grad.copy_(g.view_as(p.grad))
        opt.step()
        total = loss.detach().clone().view(1)
        mlsl.allreduce(total, group="model", distribution=dist)          # the shards' parts of the sequence loss
        losses.append(total.item())
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    lo, hi = flat.clone(), flat.clone()
    mlsl.allreduce(lo, op="min", group="data", distribution=dist)
    mlsl.allreduce(hi, op="max", group="data", distribution=dist)
    same = bool(torch.equal(lo, hi))
    ok = same and losses[-1] < losses[0]
    print("rank %d (replica %d, part %d/%d): loss %.4f -> %.4f, replicas identical: %s : %s"
          % (mlsl.rank(), rep, part, M, losses[0], losses[-1], same, "PASSED" if ok else "FAILED"), flush=True)
    env.delete_distribution(dist)
    mlsl.finalize()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
