#!/usr/bin/env python
"""Headline benchmark: all-reduce bus bandwidth (BASELINE.json: "allreduce bus GB/s vs msg size at 1/2/4/8 B200,
device-timed, max over ranks").

    python bench.py --gpus N --steps K --warmup W            # ours (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N ...            # the unmodified reference (CPU/MPI library) from baseline/_ref

One "step" = one fp32 SUM all-reduce of the headline message (1 GiB per rank, out of place, with the 1/N averaging
scale fused into the kernel) through the public API (mlsl_b200.allreduce -> Distribution::AllReduceEx ->
Environment::Wait).  value = bus bandwidth = bytes/time * 2(N-1)/N for N>1; for N=1 the factor is 0 by definition, so
the single-GPU value is the algorithm bandwidth bytes/time of the (one-kernel) local path.  Timed with CUDA events on
the launching stream, max over ranks.  The JSON line also carries a message-size sweep, the end-to-end number
(pinned host -> device -> all-reduce -> host every step) and the clocks seen during the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HEADLINE_BYTES = 1 << 30


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bytes", type=int, default=HEADLINE_BYTES, help="headline message size per rank")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--sweep-sizes", default="", help="comma separated byte sizes instead of the default 1 KiB..1 GiB x4 ladder")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--nccl", action="store_true", help="also time torch.distributed (NCCL) all_reduce for comparison")
    ap.add_argument("--compress", action="store_true", help="headline through the fp8-compressed transport")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of this rank's GPU, sampled while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            p = [x.strip() for x in s.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def busbw_factor(n):
    return 2.0 * (n - 1) / n if n > 1 else 1.0


def run_ours(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and rank == 0:
        print("warning: WORLD_SIZE=%d but --gpus %d (launch with torchrun for N>1)" % (world, args.gpus), file=sys.stderr)
    torch.cuda.set_device(local)
    os.environ.setdefault("MLSL_BACKEND", "cuda")
    S = args.bytes
    need_gb = 2 * S / 2 ** 30 + 1.0
    os.environ.setdefault("MLSL_HEAP_SIZE_GB", "%.2f" % max(need_gb, 3.5))
    os.environ.setdefault("MLSL_WATCHDOG_SEC", "60")
    # a pure collective loop: run the kernels directly on the caller's stream (no comm-stream hop, no events)
    os.environ.setdefault("MLSL_STREAM_MODE", "inline")
    use_graph = os.environ.get("MLSL_BENCH_GRAPH", "1") == "1"

    import mlsl_b200 as mlsl

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    env = mlsl.init()
    assert env.get_backend_name() == "cuda", "bench needs the CUDA backend"
    assert mlsl.world_size() == world
    n = S // 4
    x = mlsl.alloc_tensor(n, torch.float32, zero=False)
    y = mlsl.alloc_tensor(n, torch.float32, zero=False)
    x.fill_(1.0)
    scale = 1.0 / world

    def step_device(src, dst, count):
        mlsl.allreduce(src[:count], out=dst[:count], scale=scale, compress=args.compress)

    def timed(fn, steps, warm, graph=False):
        for _ in range(warm):
            fn()
        mlsl.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if graph:
            # launch-bound sizes: capture the K collectives in one CUDA graph (tickets live in device memory, so a
            # replay runs the full handshake again) and time the replay
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                for _ in range(steps):
                    fn()
            g.replay()
            torch.cuda.synchronize()
            mlsl.barrier()
            torch.cuda.synchronize()
            e0.record()
            g.replay()
            e1.record()
        else:
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
        torch.cuda.synchronize()
        mlsl.barrier()
        ms = e0.elapsed_time(e1) / steps
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        mlsl.allreduce(t, op="max")          # max over ranks, through our own library
        torch.cuda.synchronize()
        return float(t.item())

    # ---- headline ---------------------------------------------------------------------------------------------
    for _ in range(3):   # page everything in before the clock sampler starts
        step_device(x, y, n)
    torch.cuda.synchronize()
    sampler = ClockSampler(int(os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[local]) if
                           os.environ.get("CUDA_VISIBLE_DEVICES") else local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms = timed(lambda: step_device(x, y, n), args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if sampler else None
    torch.cuda.synchronize()
    ok = bool(torch.allclose(y[:1024], torch.ones(1024, device="cuda"), rtol=1e-3 if not args.compress else 0.1))
    value = S / (ms * 1e-3) / 1e9 * busbw_factor(world)

    # ---- sweep -------------------------------------------------------------------------------------------------
    sweep = []
    if not args.no_sweep:
        sizes = [1 << k for k in range(10, 31, 2)]
        if S not in sizes:
            sizes.append(S)
        if args.sweep_sizes:
            sizes = [int(v) for v in args.sweep_sizes.split(",")]
        for b in sorted(sizes):
            if b > S:
                continue
            cnt = b // 4
            it = 200 if b <= (1 << 16) else (60 if b <= (1 << 22) else (20 if b <= (1 << 26) else 6))
            gr = use_graph and b <= (1 << 22)
            m = timed(lambda: step_device(x, y, cnt), it, 5, graph=gr)
            sweep.append({"bytes": b, "us": round(m * 1e3, 3), "algbw_GBps": round(b / (m * 1e-3) / 1e9, 3),
                          "busbw_GBps": round(b / (m * 1e-3) / 1e9 * busbw_factor(world), 3), "cuda_graph": gr})

    # ---- NCCL comparison (optional) ---------------------------------------------------------------------------------
    nccl = None
    if args.nccl and world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        xt = torch.ones(n, device="cuda")
        nccl = []
        for b in sorted(set([1 << k for k in range(10, 31, 2)] + [S])):
            if b > S:
                continue
            v = xt[:b // 4]
            it = 200 if b <= (1 << 16) else (60 if b <= (1 << 22) else (20 if b <= (1 << 26) else 6))

            def f():
                dist.all_reduce(v)
                v.mul_(scale)          # the separate elementwise kernel our epilogue fuses away
            m = timed(f, it, 5)
            nccl.append({"bytes": b, "us": round(m * 1e3, 3),
                         "busbw_GBps": round(b / (m * 1e-3) / 1e9 * busbw_factor(world), 3)})
            xt.fill_(1.0)
        dist.destroy_process_group()

    # ---- end to end: pinned host -> device, all-reduce, device -> host, every step ------------------------------------
    e2e = None
    if not args.no_e2e:
        hin = torch.ones(n, dtype=torch.float32).pin_memory()
        hout = torch.empty(n, dtype=torch.float32).pin_memory()

        def step_e2e():
            # the public call with host-resident buffers: the library pipelines H2D / all-reduce / D2H in chunks
            mlsl.allreduce(hin, out=hout, scale=scale)

        try:
            ms_e = timed(step_e2e, max(3, min(args.steps, 10)), 3)
            e2e = {"value": round(S / (ms_e * 1e-3) / 1e9 * busbw_factor(world), 3), "unit": "GB/s",
                   "h2d_bytes_per_step": S, "d2h_bytes_per_step": S, "ms_per_step": round(ms_e, 4),
                   "how": "mlsl.allreduce(pinned_host_in, out=pinned_host_out): chunked H2D -> NVLink all-reduce -> D2H pipeline",
                   "correct": bool(torch.allclose(hout[:1024], torch.ones(1024), rtol=1e-3) and
                                   torch.allclose(hout[-1024:], torch.ones(1024), rtol=1e-3))}
        except Exception as ex:  # noqa: BLE001 - the device-timed headline must still be reported
            e2e = {"error": repr(ex)[:300]}

    # kernels per timed step: giant messages on the multicast (NVLS) path are issued as several launches
    # (csrc/cuda/cuda_backend.cu: >= 1.5 x MLSL_NVLS_CHUNK_MB on a group that spans the multicast object)
    launches_per_step = 1
    chunk = int(os.environ.get("MLSL_NVLS_CHUNK_MB", "256")) << 20
    nvls = ("NVLS" in mlsl.env().describe_backend() and world >= int(os.environ.get("MLSL_NVLS_MIN_RANKS", "4"))
            and os.environ.get("MLSL_NVLS", "1") != "0" and not args.compress)
    if nvls and chunk and S >= chunk + chunk // 2:
        launches_per_step = -(-S // chunk)
    out = {
        "metric": "allreduce_busbw_GBps" if world > 1 else "allreduce_algbw_GBps_single_gpu",
        "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic", "impl": "ours",
        "config": {"model": "allreduce fp32 SUM, %d MiB per rank, out of place, fused 1/N scale" % (S >> 20),
                   "global_batch": None, "seq_len": None, "parallelism": "dp%d" % world, "message_bytes": S,
                   "l2": "inputs (1 GiB) larger than L2, no flush needed", "transport": "fp8" if args.compress else "fp32",
                   "api": "mlsl_b200.allreduce -> Distribution::AllReduceEx -> Environment::Wait",
                   "backend": env.get_backend_name(), "backend_detail": env.describe_backend(), "stream_mode": os.environ.get("MLSL_STREAM_MODE"),
                   "correct": ok,
                   "value_definition": "bus bandwidth in the nccl-tests sense: S / t x 2(N-1)/N, a per-link figure that stays "
                                       "CONSTANT under ideal weak scaling (S per rank fixed); total bytes reduced per second = "
                                       "N x S / t.  N = 1 has no link: the value is S / t of the on-device scale-copy"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps * launches_per_step,
        "sweep": sweep, "nccl": nccl,
    }
    mlsl.free_tensor(x)
    mlsl.free_tensor(y)
    mlsl.finalize()
    if rank == 0:
        print(json.dumps(out))


def run_reference(args):
    """The unmodified reference library (CPU, Intel MPI runtime) driven by baseline/ref_bench.py."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    try:
        import ref_bench
        print(json.dumps(ref_bench.run(args.gpus, args.steps, max(args.warmup, 3), args.bytes)))
    except Exception as e:  # noqa: BLE001 - must never fail the driver
        print(json.dumps({"impl": "reference", "unavailable": ("%s: %s" % (type(e).__name__, e))[:300].replace("\n", " ")}))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
