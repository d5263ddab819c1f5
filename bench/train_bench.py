#!/usr/bin/env python
"""Data-parallel training throughput (BASELINE configs 3-5): ResNet-50 / BERT-large / MLP on synthetic data.

    torchrun --nproc-per-node N bench/train_bench.py --model resnet50 --impl mlsl --mode fused
    torchrun --nproc-per-node N bench/train_bench.py --model resnet50 --impl ddp            # NCCL baseline

impl mlsl : mlsl_b200.DistributedOptimizer - gradient buckets in the symmetric heap, one ParameterSet per bucket,
            communication started from autograd hooks while backward is still running;
              --mode fused      reduce-scatter + optimizer + all-gather as one kernel per bucket (distributed update)
              --mode allreduce  all-reduce with the 1/N scale fused in (+ --compress: fp8 transport), local optimizer
impl ddp  : torch DistributedDataParallel over NCCL + torch.optim (fused) - the baseline to beat.
Timing: CUDA events around K full steps (forward, backward, communication, optimizer), max over ranks.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(model, batch, seq):
    from mlsl_b200.models import MLP, resnet50
    from mlsl_b200.models.bert import BertConfig, BertEncoderModel
    if model == "resnet50":
        m = resnet50().cuda().to(memory_format=torch.channels_last)
        x = torch.randn(batch, 3, 224, 224, device="cuda").to(memory_format=torch.channels_last)
        y = torch.randint(0, 1000, (batch,), device="cuda")
        return m, (x, y), batch, "images/s", lambda out, y: torch.nn.functional.cross_entropy(out.float(), y)
    if model in ("bert-large", "bert-small"):
        cfg = BertConfig() if model == "bert-large" else BertConfig(layers=4, hidden=512, heads=8, ffn=2048)
        m = BertEncoderModel(cfg).cuda()
        x = torch.randint(0, cfg.vocab, (batch, seq), device="cuda")
        y = torch.randint(0, cfg.vocab, (batch, seq), device="cuda")
        return m, (x, y), batch * seq, "tokens/s", lambda out, y: torch.nn.functional.cross_entropy(
            out.view(-1, out.shape[-1]).float(), y.view(-1))
    m = MLP(4096, 16384, 4096, 6).cuda()
    x = torch.randn(batch, 4096, device="cuda")
    y = torch.randn(batch, 4096, device="cuda")
    return m, (x, y), batch, "samples/s", lambda out, y: torch.nn.functional.mse_loss(out.float(), y)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet50", choices=["resnet50", "bert-large", "bert-small", "mlp"])
    ap.add_argument("--impl", default="mlsl", choices=["mlsl", "ddp"])
    ap.add_argument("--mode", default="fused", choices=["fused", "allreduce"])
    ap.add_argument("--compress", action="store_true")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--seq", type=int, default=512)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--bucket-mb", type=float, default=64)
    ap.add_argument("--optimizer", default=None)
    ap.add_argument("--param-dtype", default="fp32", choices=["fp32", "bf16"],
                    help="bf16: parameters AND gradients in bf16 (half the gradient bytes on the wire; the fused sharded optimizer "
                         "keeps fp32 master weights and moments for its shard), no autocast")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    batch = args.batch or {"resnet50": 128, "bert-large": 8, "bert-small": 16, "mlp": 256}[args.model]
    kind = args.optimizer or ("adamw" if args.model.startswith("bert") else "sgd")
    torch.manual_seed(1234)
    model, (x, y), units, unit_name, loss_fn = build(args.model, batch, args.seq)
    nparam = sum(p.numel() for p in model.parameters())
    if args.param_dtype == "bf16":
        import contextlib
        model = model.to(torch.bfloat16)
        if x.is_floating_point():
            x = x.to(torch.bfloat16)
        amp = contextlib.nullcontext()
    else:
        amp = torch.autocast("cuda", dtype=torch.bfloat16)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    if args.impl == "mlsl":
        os.environ.setdefault("MLSL_BACKEND", "cuda")
        os.environ.setdefault("MLSL_HEAP_SIZE_GB", "%.1f" % max(4.0, nparam * 4 * 3.2 / 2 ** 30))
        os.environ.setdefault("MLSL_NUM_CHANNELS", "32")   # leave most SMs to the backward pass that runs concurrently
        import mlsl_b200 as mlsl
        mlsl.init()
        from mlsl_b200.parallel import broadcast_parameters
        if world > 1:
            broadcast_parameters(model)
        opt = mlsl.DistributedOptimizer(model.parameters(), lr=0.01 if kind == "sgd" else 1e-4, momentum=0.9,
                                        weight_decay=1e-4, optimizer=kind, mode=args.mode, bucket_mb=args.bucket_mb,
                                        compress=args.compress)
        net = model
        sync_max = lambda t: mlsl.allreduce(t, op="max")  # noqa: E731
    else:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], bucket_cap_mb=args.bucket_mb,
                                                        gradient_as_bucket_view=True)
        opt = (torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-4, fused=True) if kind == "adamw"
               else torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4, fused=True))
        sync_max = lambda t: dist.all_reduce(t, op=dist.ReduceOp.MAX)  # noqa: E731

    def step():
        opt.zero_grad(set_to_none=False)
        with amp:
            out = net(x)
        loss = loss_fn(out, y)
        loss.backward()
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
    sync_max(ms)
    torch.cuda.synchronize()
    ms = float(ms.item())
    if rank == 0:
        print(json.dumps({"model": args.model, "impl": args.impl, "mode": args.mode if args.impl == "mlsl" else "nccl-ddp",
                          "compress": args.compress, "n_gpus": world, "per_gpu_batch": batch, "params_M": round(nparam / 1e6, 1),
                          "ms_per_step": round(ms, 3), "throughput": round(units * world / (ms * 1e-3), 1),
                          "unit": unit_name, "optimizer": kind, "loss": round(float(loss), 4),
                          "dtype": "bf16 params/grads" if args.param_dtype == "bf16" else "bf16 autocast, fp32 params/grads"}))
    if args.impl == "mlsl":
        opt.close()
        import mlsl_b200 as mlsl
        mlsl.finalize()
    else:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
