#!/usr/bin/env python
"""Fused GEMM + reduce-scatter (tcgen05 kernel, csrc/cuda/gemm_rs.cu) vs the unfused baseline
(torch.matmul = cuBLAS, then torch.distributed reduce_scatter_tensor = NCCL).
    torchrun --nproc-per-node N bench/gemm_rs_bench.py [--shapes M,N,K ...]
K is the PER-RANK reduction length (the K-slice of a row-parallel layer).  Device-timed, max over ranks; reports the
achieved fraction of the roofline max(2MNK / bf16 peak, bytes over NVLink / 770 GB/s)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", nargs="*", default=["8192,8192,1024", "8192,8192,4096", "16384,4096,2048", "4096,12288,1536"])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-nccl", action="store_true")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    os.environ.setdefault("MLSL_BACKEND", "cuda")
    os.environ.setdefault("MLSL_HEAP_SIZE_GB", "6")
    os.environ.setdefault("MLSL_STREAM_MODE", "inline")
    import mlsl_b200 as mlsl
    from mlsl_b200.ops import gemm_reduce_scatter
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    mlsl.init()
    peaks = {"bf16_tflops": 1732.9}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    use_nccl = world > 1 and not args.no_nccl
    if use_nccl:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def timed(fn, iters):
        for _ in range(3):
            fn()
        mlsl.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device="cuda")
        mlsl.allreduce(t, op="max")
        torch.cuda.synchronize()
        return float(t.item())

    rows = []
    for shp in args.shapes:
        M, N, K = (int(v) for v in shp.split(","))
        a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
        w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
        out = mlsl.alloc_tensor((M // world, N), torch.bfloat16)
        ms_fused = timed(lambda: gemm_reduce_scatter(a, w, out=out, group="global"), args.iters)
        ms_gemm = timed(lambda: torch.matmul(a, w.t()), args.iters)
        ms_unfused = None
        if use_nccl:
            import torch.distributed as dist
            c = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            o = torch.empty(M // world, N, device="cuda", dtype=torch.bfloat16)

            def base():
                torch.matmul(a, w.t(), out=c)
                dist.reduce_scatter_tensor(o, c)
            ms_unfused = timed(base, args.iters)
        flops = 2.0 * M * N * K
        t_math = flops / (peaks["bf16_tflops"] * 1e12) * 1e3
        t_link = (M * N * 2.0 * (world - 1) / world) / 770e9 * 1e3
        roof = max(t_math, t_link)
        rows.append({"M": M, "N": N, "K_per_rank": K, "ranks": world, "fused_ms": round(ms_fused, 4),
                     "cublas_gemm_only_ms": round(ms_gemm, 4), "cublas_plus_nccl_rs_ms": None if ms_unfused is None else round(ms_unfused, 4),
                     "fused_tflops": round(flops / ms_fused / 1e9, 1), "roofline_ms": round(roof, 4),
                     "roofline_bound": "math" if t_math >= t_link else "nvlink", "frac_of_roofline": round(roof / ms_fused, 3),
                     "speedup_vs_unfused": None if ms_unfused is None else round(ms_unfused / ms_fused, 3)})
        mlsl.free_tensor(out)
    if rank == 0:
        for r in rows:
            print(json.dumps(r))
    mlsl.finalize()
    if use_nccl:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
