"""Ulysses (all-to-all) and ring (send/recv-list rotation) attention over a Distribution group vs single-process
softmax attention: forward values and the gradients of q, k, v."""
import pytest
import torch

from conftest import run_ranks

S, H, D = 32, 4, 8


def _qkv():
    g = torch.Generator().manual_seed(11)
    return [torch.randn(S, H, D, generator=g) for _ in range(3)] + [torch.randn(S, H, D, generator=g)]


def _reference(causal):
    q, k, v, t = _qkv()
    q, k, v = (x.clone().requires_grad_(True) for x in (q, k, v))
    s = torch.einsum("qhd,khd->hqk", q, k) / D ** 0.5
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(S, S, dtype=torch.bool), 1).unsqueeze(0), float("-inf"))
    o = torch.einsum("hqk,khd->qhd", torch.softmax(s, -1), v)
    ((o - t) ** 2).sum().backward()
    return o.detach(), q.grad, k.grad, v.grad


@pytest.mark.parametrize("kind", ["ulysses", "ring"])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("world", [1, 2, 4])
def test_sequence_parallel_attention(kind, causal, world):
    def body(r, mlsl):
        from mlsl_b200.parallel.sequence_parallel import ring_attention, ulysses_attention
        e = mlsl.env()
        dist = e.create_distribution(1, world)
        q, k, v, t = _qkv()
        n = S // world
        sl = slice(r * n, (r + 1) * n)
        ql, kl, vl = (x[sl].clone().requires_grad_(True) for x in (q, k, v))
        fn = ulysses_attention if kind == "ulysses" else ring_attention
        o = fn(ql, kl, vl, causal=causal, distribution=dist, group="model")
        ((o - t[sl]) ** 2).sum().backward()
        out = (o.detach(), ql.grad, kl.grad, vl.grad)
        e.delete_distribution(dist)
        return out

    outs = run_ranks(world, body)
    ref = _reference(causal)
    n = S // world
    for r, got in enumerate(outs):
        for a, b in zip(got, ref):
            assert torch.allclose(a, b[r * n:(r + 1) * n], rtol=1e-4, atol=1e-5)


def test_ring_shift_collective():
    def body(r, mlsl):
        x = torch.arange(6, dtype=torch.float32) + 10 * r
        return mlsl.ring_shift(x, 1), mlsl.ring_shift(x, -2)

    outs = run_ranks(4, body)
    for r, (a, b) in enumerate(outs):
        assert torch.equal(a, torch.arange(6, dtype=torch.float32) + 10 * ((r - 1) % 4))
        assert torch.equal(b, torch.arange(6, dtype=torch.float32) + 10 * ((r + 2) % 4))
