"""Expert parallelism: a mixture-of-experts layer whose experts are spread over a Distribution group, tokens routed with
Distribution::AlltoAllv (dispatch) and returned with the inverse exchange (combine).

AlltoAllv is part of the reference's collective set (reference include/mlsl.hpp:700-720, src/comm_ep.cpp AlltoAllv path) but
nothing in it uses it; this is the layer it exists for.  Routing is drop-less top-k: every (token, expert) assignment is
sent, the per-rank row counts travel first in one small all-to-all, so the exchange is exactly as large as the routing
decision and needs no capacity factor.  Both exchanges are autograd functions whose backward is the mirrored exchange.
"""
import torch

from .. import comm


class _RowExchange(torch.autograd.Function):
    """y = rows of x redistributed: send_rows[p] consecutive rows go to rank p, recv_rows[p] arrive from rank p."""

    @staticmethod
    def forward(ctx, x, send_rows, recv_rows, dist, group):
        ctx.cfg = (send_rows, recv_rows, dist, group, comm._state())
        d = x.shape[1]
        out = comm.alltoallv(x.contiguous().view(-1), [r * d for r in send_rows], [r * d for r in recv_rows], group=group,
                             distribution=dist)
        return out.view(sum(recv_rows), d)

    @staticmethod
    def backward(ctx, gy):
        send_rows, recv_rows, dist, group, state = ctx.cfg
        d = gy.shape[1]
        with comm.use_state(state):
            gx = comm.alltoallv(gy.contiguous().view(-1), [r * d for r in recv_rows], [r * d for r in send_rows],
                                group=group, distribution=dist)
        return gx.view(sum(send_rows), d), None, None, None, None


class ExpertParallelMoE(torch.nn.Module):
    """num_experts two-layer MLP experts, num_experts / P of them on every rank of the group; a replicated gate.

    forward(x[T, d]) -> y[T, d] = sum over the token's top-k experts of softmax-weight * expert(x).  Every rank routes its
    OWN tokens (the group is also the data dimension for the tokens, as in expert-parallel transformer layers); the
    gate's gradient is therefore a per-rank partial and has to be averaged like any data-parallel parameter."""

    def __init__(self, d_model, d_hidden, num_experts, top_k=1, group="model", distribution=None, dtype=torch.float32,
                 device=None):
        super().__init__()
        self.dist = distribution if distribution is not None else comm.world_distribution()
        self.group = group
        g = comm._group(group)
        self.P, self.idx = self.dist.get_process_count(g), self.dist.get_process_idx(g)
        if num_experts % self.P:
            raise ValueError("%d experts cannot be spread evenly over %d ranks" % (num_experts, self.P))
        self.E, self.El, self.k = num_experts, num_experts // self.P, top_k
        kw = {"dtype": dtype, "device": device}
        self.gate = torch.nn.Linear(d_model, num_experts, bias=False, **kw)
        self.w1 = torch.nn.Parameter(torch.empty(self.El, d_model, d_hidden, **kw))
        self.w2 = torch.nn.Parameter(torch.empty(self.El, d_hidden, d_model, **kw))
        for w in (self.w1, self.w2):
            torch.nn.init.normal_(w, std=w.shape[1] ** -0.5)

    def route(self, x):
        """-> (token index, expert index, weight) of every assignment, sorted by expert"""
        probs = torch.softmax(self.gate(x).float(), dim=-1)
        w, e = torch.topk(probs, self.k, dim=-1)
        w = (w / w.sum(-1, keepdim=True)).to(x.dtype)
        tok = torch.arange(x.shape[0], device=x.device).repeat_interleave(self.k)
        e, w = e.reshape(-1), w.reshape(-1)
        order = torch.sort(e, stable=True).indices
        return tok[order], e[order], w[order]

    def forward(self, x):
        T, d = x.shape
        tok, e, w = self.route(x)
        per_expert = torch.bincount(e, minlength=self.E)                              # my rows for every global expert
        # row counts per (source rank, local expert) on the receiving side
        per32 = per_expert.to(torch.int32).contiguous()
        theirs = per32
        if self.P > 1:
            theirs = torch.empty_like(per32)
            comm.alltoall(per32, out=theirs, group=self.group, distribution=self.dist)
        send_rows = per_expert.view(self.P, self.El).sum(1).tolist()
        recv_matrix = theirs.view(self.P, self.El)                                    # [src rank][local expert]
        recv_rows = recv_matrix.sum(1).tolist()
        rows = x[tok]                                                                 # sorted by expert = by dest rank
        if self.P > 1:
            rows = _RowExchange.apply(rows, send_rows, recv_rows, self.dist, self.group)
        # arrived rows: for every source rank, its rows for my expert 0, 1, ...; run each expert on its slices
        counts = recv_matrix.tolist()
        out = torch.empty_like(rows)
        starts, off = [], 0
        for src in range(self.P):
            starts.append([])
            for le in range(self.El):
                starts[src].append(off)
                off += counts[src][le]
        for le in range(self.El):
            idx = [torch.arange(starts[s][le], starts[s][le] + counts[s][le], device=x.device) for s in range(self.P)]
            idx = torch.cat(idx) if idx else torch.empty(0, dtype=torch.long, device=x.device)
            if idx.numel() == 0:
                continue
            h = torch.relu(rows[idx] @ self.w1[le])
            out = out.index_copy(0, idx, h @ self.w2[le])
        if self.P > 1:
            out = _RowExchange.apply(out, recv_rows, send_rows, self.dist, self.group)
        y = torch.zeros_like(x)
        return y.index_add(0, tok, out * w.unsqueeze(1))
