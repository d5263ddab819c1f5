import torch

from .. import comm
from ..api import DataType


def gemm_reduce_scatter(a, w, out=None, out_dtype=torch.bfloat16, group="model", distribution=None, async_op=False):
    """Row-parallel linear layer in one kernel: out[M/P, N] = reduce_scatter_rows(a[M, K_local] @ w[N, K_local].T).

    a, w: bf16, contiguous (w has the nn.Linear weight layout).  The tcgen05 GEMM's epilogue pushes every partial tile
    into the owning rank's staging area over NVLink while the tensor core runs the next tile; the owner sums the P
    partials.  Returns this rank's rows ([M/P, N])."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.is_contiguous() and w.is_contiguous()
    M, K = a.shape
    N, K2 = w.shape
    assert K == K2
    d = distribution if distribution is not None else comm.world_distribution()
    g = comm._group(group)
    P = d.get_process_count(g)
    if out is None:
        out = comm.alloc_tensor((M // P, N), out_dtype, zero=False)
    comm._sync_stream()
    req = d.gemm_reduce_scatter(a, w, out, M, N, K, DataType.FLOAT if out.dtype == torch.float32 else DataType.BF16, g)
    work = comm.Work(comm.env(), req, out, (a, w, out))
    return work if async_op else work.wait()
