"""Tensor (model) parallel linear layers on the model group of a Distribution.

This is the reference's "model parallelism" (OT_CC with modelParts > 1, reference src/mlsl_impl.cpp:139-175) as PyTorch
modules: the weight's reduction dimension is split over the model group, every rank produces a full-size PARTIAL sum,
forward needs a ReduceScatter and backward the matching AllGather (activation exchange "case 1").  On the CUDA backend
the forward GEMM and its reduce-scatter are ONE kernel (mlsl_b200.ops.gemm_reduce_scatter: tcgen05 GEMM whose epilogue
pushes the partial tiles to their owner over NVLink); everywhere else it is matmul + Distribution::ReduceScatter.

    ColumnParallelLinear : W[out/P, in]  y_local = x @ W^T           (no communication forward, all-reduce of dX backward)
    RowParallelLinear    : W[out, in/P]  y[M/P, out] = reduce_scatter_rows(x_local @ W^T)   (rows = tokens)
"""
import torch

from .. import comm


def _group_info(distribution, group):
    d = distribution if distribution is not None else comm.world_distribution()
    g = comm._group(group)
    return d, g, d.get_process_count(g), d.get_process_idx(g)


class _RowParallelMatmul(torch.autograd.Function):
    """y[M/P, N] = reduce_scatter_rows(x[M, K_local] @ w[N, K_local]^T)"""

    @staticmethod
    def forward(ctx, x, w, distribution, group, fused):
        d, g, P, idx = _group_info(distribution, group)
        ctx.save_for_backward(x, w)
        ctx.cfg = (distribution, group)
        ctx.mlsl_state = comm._state()
        M, N = x.shape[0], w.shape[0]
        use_fused = (fused and comm.is_device() and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and
                     M % (128 * P) == 0 and N % 256 == 0 and x.shape[1] % 64 == 0)
        if use_fused:
            from ..ops import gemm_reduce_scatter
            out = gemm_reduce_scatter(x.contiguous(), w.contiguous(), group=group, distribution=distribution)
            return out.clone()      # the op's result buffer is heap memory owned by the library
        partial = (x @ w.t()).contiguous()
        if P == 1:
            return partial
        out = comm.reduce_scatter(partial.view(-1), group=group, distribution=distribution)
        return out.view(M // P, N).clone()

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        distribution, group = ctx.cfg
        gy = gy.contiguous()
        with comm.use_state(ctx.mlsl_state):
            d, g, P, idx = _group_info(distribution, group)
            if P > 1:
                full = comm.allgather(gy.view(-1), group=group, distribution=distribution).view(gy.shape[0] * P, gy.shape[1])
            else:
                full = gy
        gx = full.to(w.dtype) @ w             # [M, K_local]
        gw = full.to(x.dtype).t() @ x         # [N, K_local]
        return gx, gw, None, None, None


class _CopyToModelGroup(torch.autograd.Function):
    """identity forward, all-reduce backward (the input of a column-parallel layer is replicated)"""

    @staticmethod
    def forward(ctx, x, distribution, group):
        ctx.cfg = (distribution, group)
        ctx.mlsl_state = comm._state()
        return x

    @staticmethod
    def backward(ctx, gx):
        distribution, group = ctx.cfg
        with comm.use_state(ctx.mlsl_state):
            d, g, P, idx = _group_info(distribution, group)
            if P > 1:
                gx = gx.contiguous().clone()
                comm.allreduce(gx.view(-1), group=group, distribution=distribution)
        return gx, None, None


class _SeqParallelColumnMatmul(torch.autograd.Function):
    """y[M, N/P] = all_gather_rows(x[M/P, K]) @ w[N/P, K]^T   (Megatron-style sequence parallelism in front of a column
    parallel layer).  Forward: all-gather + GEMM (one kernel when the fused op is enabled, mlsl_b200.ops.allgather_gemm);
    backward: dX[M/P, K] = reduce_scatter_rows(dY @ w) - the same shape of problem as the row-parallel forward, so it
    runs on the fused GEMM + reduce-scatter kernel - and dW = dY^T @ X_full with the gathered X kept from forward."""

    @staticmethod
    def forward(ctx, x, w, distribution, group, fused):
        from ..ops import allgather_gemm
        ctx.cfg = (distribution, group, fused)
        ctx.mlsl_state = comm._state()
        y, full = allgather_gemm(x, w, group=group, distribution=distribution, fused=fused)
        ctx.save_for_backward(full, w)
        return y

    @staticmethod
    def backward(ctx, gy):
        full, w = ctx.saved_tensors
        distribution, group, fused = ctx.cfg
        gy = gy.contiguous()
        with comm.use_state(ctx.mlsl_state):
            d, g, P, idx = _group_info(distribution, group)
            M, Nl = gy.shape
            K = w.shape[1]
            use_fused = (fused is not False and comm.is_device() and gy.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and
                         P > 1 and M % (128 * P) == 0 and K % 256 == 0 and Nl % 64 == 0)
            if use_fused:
                from ..ops import gemm_reduce_scatter
                gx = gemm_reduce_scatter(gy, w.t().contiguous(), group=group, distribution=distribution).clone()
            else:
                partial = (gy.to(w.dtype) @ w).contiguous()                 # [M, K]: this rank's share of dX for ALL rows
                if P == 1:
                    gx = partial
                else:
                    gx = comm.reduce_scatter(partial.view(-1), group=group, distribution=distribution).view(M // P, K).clone()
        gw = gy.to(full.dtype).t() @ full                                   # [N/P, K]
        return gx, gw, None, None, None


class ColumnParallelLinear(torch.nn.Module):
    """Output features are split over the model group.  sequence_parallel=False: the input is replicated, [M, in] ->
    [M, out/P] (all-reduce of dX in backward).  sequence_parallel=True: the input arrives split over the token rows,
    [M/P, in] -> [M, out/P]: all-gather + GEMM forward, GEMM + reduce-scatter backward - together with
    RowParallelLinear (whose output is row-split again) no activation is ever replicated between the two layers' ends."""

    def __init__(self, in_features, out_features, bias=True, distribution=None, group="model", dtype=None, device=None,
                 sequence_parallel=False, fused=None):
        super().__init__()
        self.distribution, self.group = distribution, group
        self.sequence_parallel, self.fused = sequence_parallel, fused
        _, _, P, idx = _group_info(distribution, group)
        assert out_features % P == 0
        self.weight = torch.nn.Parameter(torch.empty(out_features // P, in_features, dtype=dtype, device=device))
        self.bias = torch.nn.Parameter(torch.zeros(out_features // P, dtype=dtype, device=device)) if bias else None
        torch.nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)

    def forward(self, x):
        if self.sequence_parallel:
            y = _SeqParallelColumnMatmul.apply(x, self.weight, self.distribution, self.group, self.fused)
            return y if self.bias is None else y + self.bias
        x = _CopyToModelGroup.apply(x, self.distribution, self.group)
        return torch.nn.functional.linear(x, self.weight, self.bias)


class RowParallelLinear(torch.nn.Module):
    """Input features are split over the model group; the output rows (tokens) come back reduce-scattered:
    [M, in/P] -> [M/P, out]."""

    def __init__(self, in_features, out_features, bias=True, distribution=None, group="model", dtype=None, device=None,
                 fused=True):
        super().__init__()
        self.distribution, self.group, self.fused = distribution, group, fused
        _, _, P, idx = _group_info(distribution, group)
        assert in_features % P == 0
        self.weight = torch.nn.Parameter(torch.empty(out_features, in_features // P, dtype=dtype, device=device))
        self.bias = torch.nn.Parameter(torch.zeros(out_features, dtype=dtype, device=device)) if bias else None
        torch.nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)

    def forward(self, x):
        y = _RowParallelMatmul.apply(x, self.weight, self.distribution, self.group, self.fused)
        return y if self.bias is None else y + self.bias


def gather_rows(x, distribution=None, group="model", replicated_downstream=True):
    """[M/P, N] -> [M, N]: differentiable all-gather of the token dimension.  Backward: when the consumer is replicated
    on every rank (each computes the same full gradient) this rank's rows are simply sliced out; when the consumer is
    itself model parallel (each rank holds a PARTIAL gradient) the gradients are reduce-scattered."""
    return _GatherRows.apply(x, distribution, group, replicated_downstream)


class _GatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, distribution, group, replicated):
        ctx.cfg = (distribution, group, replicated)
        ctx.mlsl_state = comm._state()
        d, g, P, idx = _group_info(distribution, group)
        if P == 1:
            return x
        out = comm.allgather(x.contiguous().view(-1), group=group, distribution=distribution)
        return out.view(x.shape[0] * P, *x.shape[1:]).clone()

    @staticmethod
    def backward(ctx, gy):
        distribution, group, replicated = ctx.cfg
        with comm.use_state(ctx.mlsl_state):
            d, g, P, idx = _group_info(distribution, group)
            if P == 1:
                return gy, None, None, None
            rows = gy.shape[0] // P
            if replicated:
                return gy[idx * rows:(idx + 1) * rows].contiguous(), None, None, None
            out = comm.reduce_scatter(gy.contiguous().view(-1), group=group, distribution=distribution)
            return out.view(rows, *gy.shape[1:]).clone(), None, None, None
