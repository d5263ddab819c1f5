"""Sweep the grid / unroll of the large all-reduce kernel inside ONE job (Environment.set_tuning changes the knobs on
every rank at the same point of the program): torchrun --nproc-per-node N bench/tune_allreduce.py [--mb 1024,256,64]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
ap = argparse.ArgumentParser()
ap.add_argument("--mb", default="1024,256,64")
ap.add_argument("--channels", default="48,64,96,128,148")
ap.add_argument("--unroll", default="1,2,4")
ap.add_argument("--hybrid", default="", help="'pct:cta_pct,...' pairs of the multicast + peer-to-peer hybrid to try at the best grid")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("MLSL_BACKEND", "cuda")
os.environ.setdefault("MLSL_HEAP_SIZE_GB", "3.5")
os.environ.setdefault("MLSL_STREAM_MODE", "inline")
import mlsl_b200 as mlsl  # noqa: E402

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
env = mlsl.init()
nmax = max(int(m) for m in args.mb.split(",")) << 18
x = mlsl.alloc_tensor(nmax, torch.float32, zero=False)
y = mlsl.alloc_tensor(nmax, torch.float32, zero=False)
x.fill_(1.0)


def timed(n, iters):
    for _ in range(2):
        mlsl.allreduce(x[:n], out=y[:n], scale=1.0 / world)
    mlsl.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        mlsl.allreduce(x[:n], out=y[:n], scale=1.0 / world)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device="cuda")
    mlsl.allreduce(t, op="max")
    torch.cuda.synchronize()
    return float(t.item())


for mb in (int(m) for m in args.mb.split(",")):
    n = mb << 18
    best = None
    for ch in (int(c) for c in args.channels.split(",")):
        for u in (int(v) for v in args.unroll.split(",")):
            env.set_tuning("ar_channels", ch)
            env.set_tuning("ar_unroll", u)
            ms = timed(n, 8 if mb >= 256 else 20)
            alg = (mb << 20) / (ms * 1e-3) / 1e9
            row = {"mb": mb, "channels": ch, "unroll": u, "us": round(ms * 1e3, 1), "algbw_GBps": round(alg, 1),
                   "busbw_GBps": round(alg * 2 * (world - 1) / world, 1)}
            if best is None or ms < best[0]:
                best = (ms, row)
            if rank == 0:
                print(json.dumps(row), flush=True)
    if rank == 0:
        print("BEST", json.dumps(best[1]), flush=True)
    if args.hybrid:
        env.set_tuning("ar_channels", best[1]["channels"])
        env.set_tuning("ar_unroll", best[1]["unroll"])
        for pair in args.hybrid.split(","):
            pct, cta = (int(v) for v in pair.split(":"))
            for ch in sorted({best[1]["channels"], 148}):
                env.set_tuning("ar_channels", ch)
                env.set_tuning("ar_p2p_pct", pct)
                env.set_tuning("ar_p2p_cta_pct", cta)
                ms = timed(n, 8 if mb >= 256 else 20)
                alg = (mb << 20) / (ms * 1e-3) / 1e9
                if rank == 0:
                    print(json.dumps({"mb": mb, "channels": ch, "hybrid_p2p_pct": pct, "hybrid_cta_pct": cta, "us": round(ms * 1e3, 1),
                                      "algbw_GBps": round(alg, 1), "busbw_GBps": round(alg * 2 * (world - 1) / world, 1)}), flush=True)
        env.set_tuning("ar_p2p_pct", 0)
mlsl.finalize()
