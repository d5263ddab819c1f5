"""What the host link of every GPU delivers while ALL ranks of the node copy at once: pinned host -> device, device ->
host, and both directions together (the shape of the end-to-end all-reduce pipeline), per rank and as the node total.
    torchrun --nproc-per-node N bench/pcie_probe.py [--mb 1024] [--no-bind]
The ceiling the end-to-end number of bench.py is judged against (PCIe Gen5 x16: ~55 GB/s per direction per GPU when
nothing is shared; GPUs behind one switch / one socket share uplinks and memory controllers)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=1024)
ap.add_argument("--no-bind", action="store_true", help="MLSL_NUMA_BIND=0: leave the process where the launcher put it")
args = ap.parse_args()
if args.no_bind:
    os.environ["MLSL_NUMA_BIND"] = "0"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("MLSL_BACKEND", "cuda")
os.environ.setdefault("MLSL_HEAP_SIZE_GB", "0.5")
import mlsl_b200 as mlsl  # noqa: E402

mlsl.init()               # binds the process to the GPU's NUMA node before the pinned buffers are allocated
n = args.mb << 18
hin, hout = torch.ones(n).pin_memory(), torch.empty(n).pin_memory()
din, dout = torch.empty(n, device="cuda"), torch.ones(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, iters=4):
    res = []
    for _ in range(iters + 1):
        mlsl.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_stream(torch.cuda.current_stream())
        s2.wait_stream(torch.cuda.current_stream())
        if up:
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1))
    ms = min(res[1:])
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    mlsl.allreduce(t, op="max")
    torch.cuda.synchronize()
    return float(t.item())


out = {"world": world, "mb": args.mb, "numa_bind": not args.no_bind, "cpus": len(os.sched_getaffinity(0))}
for name, up, down in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
    ms = run(up, down)
    out[name + "_GBps_per_gpu_per_dir"] = round(n * 4 / (ms * 1e-3) / 1e9, 2)
if rank == 0:
    print(json.dumps(out), flush=True)
mlsl.finalize()
