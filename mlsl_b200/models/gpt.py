"""Megatron-style transformer blocks on the library's tensor + sequence parallel layers.

Between the blocks the activations are split over the token rows of the model group ([M/P, d] per rank, M = tokens of
one sequence), inside a block over heads / hidden units:

    x[M/P, d] -> LN -> ColumnParallelLinear(sequence_parallel)  all-gather + GEMM   -> qkv[M, 3d/P]   (H/P heads per rank)
              -> causal attention over the local heads                               -> ctx[M, d/P]
              -> RowParallelLinear                               GEMM + reduce-scatter -> [M/P, d]  (+ residual)
              -> LN -> ColumnParallelLinear(sequence_parallel) -> GELU -> RowParallelLinear -> [M/P, d]  (+ residual)

so no activation is ever replicated and every communication is one of the two fused kernels on the CUDA backend
(mlsl_b200.ops.allgather_gemm / gemm_reduce_scatter); on the host backend the same graph runs on all-gather /
reduce-scatter + matmul.  The reference has no model code (it is a communication library); this is the layer stack its
model-parallel OT_CC operations (reference src/mlsl_impl.cpp:139-175) were written for, in today's shape.
"""
import math

import torch

from ..parallel.tensor_parallel import ColumnParallelLinear, RowParallelLinear, _group_info


class ParallelTransformerBlock(torch.nn.Module):
    def __init__(self, d_model, n_heads, d_ff=None, distribution=None, group="model", dtype=None, device=None, causal=True,
                 fused=None):
        super().__init__()
        _, _, P, _ = _group_info(distribution, group)
        if n_heads % P or d_model % n_heads:
            raise ValueError("n_heads (%d) must be a multiple of the group size (%d) and divide d_model (%d)"
                             % (n_heads, P, d_model))
        d_ff = d_ff or 4 * d_model
        self.P, self.heads_local, self.head_dim, self.causal = P, n_heads // P, d_model // n_heads, causal
        kw = dict(distribution=distribution, group=group, dtype=dtype, device=device)
        self.ln1 = torch.nn.LayerNorm(d_model, dtype=dtype, device=device)
        self.ln2 = torch.nn.LayerNorm(d_model, dtype=dtype, device=device)
        # per rank: [q | k | v] of its own heads
        self.qkv = ColumnParallelLinear(d_model, 3 * d_model, sequence_parallel=True, fused=fused, **kw)
        self.proj = RowParallelLinear(d_model, d_model, **kw)
        self.fc1 = ColumnParallelLinear(d_model, d_ff, sequence_parallel=True, fused=fused, **kw)
        self.fc2 = RowParallelLinear(d_ff, d_model, **kw)

    def attention(self, qkv):
        M = qkv.shape[0]
        q, k, v = qkv.view(M, 3, self.heads_local, self.head_dim).permute(1, 2, 0, 3)      # [h, M, hd] each
        out = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=self.causal,
                                                               scale=1.0 / math.sqrt(self.head_dim))
        return out.permute(1, 0, 2).reshape(M, self.heads_local * self.head_dim)

    def forward(self, x):
        """x: this rank's token rows [M/P, d_model]; returns the same layout."""
        x = x + self.proj(self.attention(self.qkv(self.ln1(x))))
        return x + self.fc2(torch.nn.functional.gelu(self.fc1(self.ln2(x))))

    def layer_norm_parameters(self):
        """Replicated parameters whose gradients are partial sums over the token shards: all-reduce them over the model
        group before the optimizer step (the sharded weights need nothing)."""
        return [p for m in (self.ln1, self.ln2) for p in m.parameters()] + \
               [b for b in (self.proj.bias, self.fc2.bias) if b is not None]


class ParallelTransformer(torch.nn.Module):
    """n_layers blocks; input and output are token-row shards [M/P, d_model]."""

    def __init__(self, n_layers, d_model, n_heads, **kw):
        super().__init__()
        self.blocks = torch.nn.ModuleList([ParallelTransformerBlock(d_model, n_heads, **kw) for _ in range(n_layers)])

    def forward(self, x):
        for b in self.blocks:
            x = b(x)
        return x

    def replicated_parameters(self):
        return [p for b in self.blocks for p in b.layer_norm_parameters()]
