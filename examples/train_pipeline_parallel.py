#!/usr/bin/env python
"""Data x pipeline parallel training of a deep MLP.

    torchrun --nproc-per-node 8 examples/train_pipeline_parallel.py --stages 4        # 2 replicas x 4 stages
    bin/mlslrun -n 4 python examples/train_pipeline_parallel.py --stages 2            # CPU, host backend

Distribution(replicas, stages): the model group of a rank is its pipeline (consecutive ranks = consecutive stages), the
data group the same stage of every replica.  A step runs `--micro` micro-batches through the GPipe schedule of
mlsl_b200.parallel.pipeline_parallel (one collective neighbour exchange per tick), then averages each stage's gradients
over its data group."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsl_b200 as mlsl  # noqa: E402
from mlsl_b200.parallel.pipeline_parallel import PipelineStage, bubble_fraction  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stages", type=int, default=2)
    ap.add_argument("--micro", type=int, default=4)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--width", type=int, default=64)
    ap.add_argument("--micro-batch", type=int, default=16)
    ap.add_argument("--layers-per-stage", type=int, default=2)
    ap.add_argument("--schedule", default="1f1b", choices=["gpipe", "1f1b"])
    args = ap.parse_args()
    use_cuda = torch.cuda.is_available() and os.environ.get("MLSL_BACKEND", "cuda") == "cuda"
    if use_cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    env = mlsl.init()
    world, S = mlsl.world_size(), args.stages
    assert world % S == 0, "world size must be a multiple of --stages"
    replicas = world // S
    dist = env.create_distribution(replicas, S)
    rep, stage = dist.get_process_idx(mlsl.GroupType.DATA), dist.get_process_idx(mlsl.GroupType.MODEL)
    dev = "cuda" if use_cuda else "cpu"
    torch.manual_seed(100 + stage)                        # the same stage weights on every replica
    layers = []
    for _ in range(args.layers_per_stage):
        layers += [torch.nn.Linear(args.width, args.width), torch.nn.Tanh()]
    block = torch.nn.Sequential(*layers).to(dev)
    shape = (args.micro_batch, args.width)
    st = PipelineStage(block, shape, shape, group="model", distribution=dist)
    opt = torch.optim.SGD(block.parameters(), lr=0.2)
    gen = torch.Generator().manual_seed(7 + rep)          # one data shard per replica, a fixed batch: the loss must fall
    xs = [torch.randn(*shape, generator=gen).to(dev) for _ in range(args.micro)]
    ys = [torch.tanh(x.roll(1, dims=1)) * 0.5 for x in xs]
    losses = []
    for step in range(args.steps):
        opt.zero_grad()
        loss = st.step(xs if st.is_first else None, loss_fn=torch.nn.functional.mse_loss if st.is_last else None,
                       targets=ys if st.is_last else None, num_micro=args.micro, schedule=args.schedule)
        for p in block.parameters():                      # average this stage's gradients over the replicas
            g = p.grad.contiguous().view(-1)
            mlsl.allreduce(g, group="data", distribution=dist, scale=1.0 / replicas)
            p.grad.copy_(g.view_as(p.grad))
        opt.step()
        if loss is not None:
            losses.append(loss.item())
    # replicas of a stage must have stayed identical
    flat = torch.cat([p.detach().reshape(-1) for p in block.parameters()])
    lo, hi = flat.clone(), flat.clone()
    mlsl.allreduce(lo, op="min", group="data", distribution=dist)
    mlsl.allreduce(hi, op="max", group="data", distribution=dist)
    same = bool(torch.equal(lo, hi))
    ok = same and (not losses or losses[-1] < losses[0])
    tail = "loss %.4f -> %.4f, " % (losses[0], losses[-1]) if losses else ""
    print("rank %d (replica %d, stage %d/%d): %sbubble %.0f %%, replicas identical: %s : %s"
          % (mlsl.rank(), rep, stage, S, tail, 100 * bubble_fraction(S, args.micro), same, "PASSED" if ok else "FAILED"), flush=True)
    env.delete_distribution(dist)
    mlsl.finalize()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
