"""One GPU: k_gemm_rs (cta_group::1) vs k_gemm_rs2 (cta_group::2, MLSL_GEMM_2CTA=1) vs cuBLAS, plus a numerics check."""
import os
import sys

os.environ.setdefault("MLSL_BACKEND", "cuda")
os.environ.setdefault("MLSL_HEAP_SIZE_GB", "4")
os.environ.setdefault("MLSL_STREAM_MODE", "inline")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mlsl_b200 as mlsl  # noqa: E402
from mlsl_b200.ops import gemm_reduce_scatter  # noqa: E402

torch.cuda.set_device(0)
torch.cuda.set_stream(torch.cuda.Stream())
mlsl.init()


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for M, N, K in ((512, 768, 512), (8192, 8192, 4096), (8192, 8192, 1024), (16384, 4096, 2048)):
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    out = mlsl.alloc_tensor((M, N), torch.bfloat16)
    res = {}
    for flag in ("0", "1"):
        os.environ["MLSL_GEMM_2CTA"] = flag
        res[flag] = timed(lambda: gemm_reduce_scatter(a, w, out=out, group="global"))
        res["out" + flag] = out.float().clone()
    ref = (a.float() @ w.float().t())
    res["cublas"] = timed(lambda: torch.matmul(a, w.t()))
    err = [(res["out" + f] - ref).abs().max().item() / ref.abs().max().item() for f in ("0", "1")]
    tf = lambda ms: 2.0 * M * N * K / ms / 1e9  # noqa: E731
    print("M %d N %d K %d: 1cta %.4f ms (%.0f TF) 2cta %.4f ms (%.0f TF) cublas %.4f ms (%.0f TF) relerr %.2e %.2e" % (
        M, N, K, res["0"], tf(res["0"]), res["1"], tf(res["1"]), res["cublas"], tf(res["cublas"]), err[0], err[1]), flush=True)
mlsl.finalize()
