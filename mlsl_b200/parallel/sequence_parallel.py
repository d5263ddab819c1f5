"""Sequence (context) parallelism on a Distribution group - an extension: the reference scales the batch and the
feature maps only, but it exposes exactly the primitives long-context training needs: AlltoAll (the
head <-> sequence transpose of DeepSpeed-Ulysses) and the send/recv-list op it declares and never wires up
(reference src/comm.hpp:212-248), which is the KV rotation of ring attention.

    ulysses_attention : every rank holds seq/P tokens of ALL heads; one all-to-all gives it ALL tokens of heads/P
                        heads, attention runs locally, a second all-to-all restores the token split.
    ring_attention    : every rank keeps its queries; key/value blocks travel round the ring (comm.ring_shift) while the
                        partial results are merged with the online-softmax rule, so no rank ever holds the full sequence.

Layout everywhere: [tokens, heads, head_dim] with the token dimension split over the group in rank order.  Both are
differentiable (autograd Functions whose backward is the mirrored collective)."""
import math

import torch

from .. import comm


def _info(distribution, group):
    d = distribution if distribution is not None else comm.world_distribution()
    g = comm._group(group)
    return d, d.get_process_count(g), d.get_process_idx(g)


class _AllToAll(torch.autograd.Function):
    """x viewed as P equal chunks along dim 0 after the caller's permutation; chunk p goes to rank p."""

    @staticmethod
    def forward(ctx, x, distribution, group):
        ctx.cfg = (distribution, group)
        ctx.mlsl_state = comm._state()
        out = comm.alltoall(x.contiguous().view(-1), group=group, distribution=distribution)
        return out.view(x.shape).clone()

    @staticmethod
    def backward(ctx, g):
        distribution, group = ctx.cfg
        with comm.use_state(ctx.mlsl_state):
            out = comm.alltoall(g.contiguous().view(-1), group=group, distribution=distribution)
            return out.view(g.shape).clone(), None, None


def seq_to_heads(x, distribution=None, group="model"):
    """[S/P, H, D] (my tokens, all heads) -> [S, H/P, D] (all tokens, my heads)."""
    d, P, _ = _info(distribution, group)
    if P == 1:
        return x
    s, H, D = x.shape
    assert H % P == 0, "the number of heads must be divisible by the sequence-parallel group size"
    # chunk p = heads of rank p: [P, s, H/P, D]
    send = x.view(s, P, H // P, D).permute(1, 0, 2, 3).contiguous()
    recv = _AllToAll.apply(send, distribution, group)            # [P(source rank = token block), s, H/P, D]
    return recv.reshape(P * s, H // P, D)


def heads_to_seq(x, distribution=None, group="model"):
    """[S, H/P, D] -> [S/P, H, D]: inverse of seq_to_heads."""
    d, P, _ = _info(distribution, group)
    if P == 1:
        return x
    S, h, D = x.shape
    s = S // P
    send = x.view(P, s, h, D).contiguous()                       # chunk p = token block of rank p
    recv = _AllToAll.apply(send, distribution, group)            # [P(source rank = head block), s, h, D]
    return recv.permute(1, 0, 2, 3).reshape(s, P * h, D)


def _attention(q, k, v, causal, q_offset=0, k_offset=0):
    """plain softmax attention on [T, H, D] tensors -> (out [T, H, D], log-sum-exp [H, T]); fp32 accumulation"""
    scale = 1.0 / math.sqrt(q.shape[-1])
    s = torch.einsum("qhd,khd->hqk", q.float(), k.float()) * scale
    if causal:
        qi = torch.arange(q.shape[0], device=q.device).view(-1, 1) + q_offset
        ki = torch.arange(k.shape[0], device=q.device).view(1, -1) + k_offset
        s = s.masked_fill((ki > qi).unsqueeze(0), float("-inf"))
    lse = torch.logsumexp(s, dim=-1)                             # [H, T]; -inf where a query sees no key of this block
    p = torch.exp(s - torch.nan_to_num(lse, neginf=0.0).unsqueeze(-1))
    p = torch.where(torch.isinf(lse).unsqueeze(-1), torch.zeros_like(p), p)
    return torch.einsum("hqk,khd->qhd", p, v.float()), lse


def ulysses_attention(q, k, v, causal=False, distribution=None, group="model", attn_fn=None):
    """q, k, v: [S/P, H, D] local token blocks.  Returns [S/P, H, D]."""
    qh, kh, vh = (seq_to_heads(t, distribution, group) for t in (q, k, v))
    if attn_fn is not None:
        o = attn_fn(qh, kh, vh, causal)
    else:
        o = _attention(qh, kh, vh, causal)[0].to(q.dtype)
    return heads_to_seq(o, distribution, group)


class _RingShift(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, shift, distribution, group):
        ctx.cfg = (shift, distribution, group)
        ctx.mlsl_state = comm._state()
        return comm.ring_shift(x.contiguous().view(-1), shift, group=group, distribution=distribution).view(x.shape).clone()

    @staticmethod
    def backward(ctx, g):
        shift, distribution, group = ctx.cfg
        with comm.use_state(ctx.mlsl_state):
            back = comm.ring_shift(g.contiguous().view(-1), -shift, group=group, distribution=distribution)
            return back.view(g.shape).clone(), None, None, None


def ring_shift(x, shift=1, distribution=None, group="model"):
    """differentiable comm.ring_shift"""
    d, P, _ = _info(distribution, group)
    return x if P == 1 else _RingShift.apply(x, shift, distribution, group)


def ring_attention(q, k, v, causal=False, distribution=None, group="model"):
    """q, k, v: [S/P, H, D] local token blocks (rank i holds tokens [i*S/P, (i+1)*S/P)).  Returns [S/P, H, D].
    P steps: attend to the resident KV block, pass it on, merge with the running result by log-sum-exp weights."""
    d, P, idx = _info(distribution, group)
    T = q.shape[0]
    out, lse = None, None
    kv = torch.stack([k, v])                                     # one message per step
    for step in range(P):
        src = (idx - step) % P                                   # owner of the block that is resident now
        if not (causal and src > idx):                           # a block entirely in the future contributes nothing
            o_b, lse_b = _attention(q, kv[0], kv[1], causal, q_offset=idx * T, k_offset=src * T)
            if out is None:
                out, lse = o_b, lse_b
            else:
                new = torch.logaddexp(lse, lse_b)
                w_old = torch.exp(lse - new).transpose(0, 1).unsqueeze(-1)      # [T, H, 1]
                w_new = torch.exp(lse_b - new).transpose(0, 1).unsqueeze(-1)
                out, lse = out * w_old + o_b * w_new, new
        else:
            # keep the block in the autograd graph: the rotation's backward is a collective every rank must enter
            out = out + 0.0 * kv.float().sum()
        if step + 1 < P:
            kv = ring_shift(kv, 1, distribution, group)
    return out.to(q.dtype)
