import os

import torch

from .. import comm
from ..api import DataType


def allgather_gemm(x_shard, w, out_dtype=None, group="model", distribution=None, fused=None):
    """Column-parallel linear layer under sequence parallelism: every rank holds a row shard x_shard[M/P, K] and needs
    y[M, N] = concat_rows(x_0 .. x_{P-1}) @ w[N, K].T.  Returns (y, x_full); x_full[M, K] is kept for the backward pass.

    fused (default on the CUDA backend, MLSL_AG_GEMM=0 switches it off; bf16, M % (128 P) == 0, N % 256 == 0, K % 64 == 0): ONE
    kernel - copy CTAs stream the peers' shards over NVLink into x_full with bulk copies while tensor-core CTAs already
    multiply the row tiles that have landed (csrc/cuda/ag_gemm.cu).  Otherwise: Distribution all-gather, then matmul."""
    d = distribution if distribution is not None else comm.world_distribution()
    g = comm._group(group)
    P = d.get_process_count(g)
    rows, K = x_shard.shape
    N = w.shape[0]
    M = rows * P
    if fused is None:          # tuning knob `ag_gemm` (MLSL_AG_GEMM, default on): the fused kernel whenever the shape allows
        fused = comm.is_device() and comm.env().get_tuning("ag_gemm") != 0
    ok = (fused and comm.is_device() and x_shard.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and
          M % (128 * P) == 0 and N % 256 == 0 and K % 64 == 0)
    out_dtype = out_dtype or x_shard.dtype
    if not ok:
        if P == 1:
            full = x_shard
        else:
            full = comm.allgather(x_shard.contiguous().view(-1), group=group, distribution=distribution).view(M, K)
        return (full @ w.t()).to(out_dtype), full
    assert out_dtype in (torch.bfloat16, torch.float32)
    x_shard, w = x_shard.contiguous(), w.contiguous()
    full = torch.empty(M, K, dtype=torch.bfloat16, device=x_shard.device)
    y = torch.empty(M, N, dtype=out_dtype, device=x_shard.device)
    comm._sync_stream()
    req = d.all_gather_gemm(x_shard, w, full, y, M, N, K, DataType.FLOAT if out_dtype == torch.float32 else DataType.BF16, g)
    comm.Work(comm.env(), req, y, (x_shard, w, full, y)).wait()
    return y, full
