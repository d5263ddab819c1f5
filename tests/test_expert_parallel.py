"""Expert-parallel mixture of experts (mlsl_b200/parallel/expert_parallel.py) against a dense single-process evaluation
with the same gate and expert weights: outputs, input gradients, expert-weight gradients and gate gradients."""
import pytest
import torch

from conftest import run_ranks

D, H, E, T = 8, 16, 4, 10


def _weights():
    g = torch.Generator().manual_seed(3)
    return (torch.randn(E, D, generator=g) * 0.7, torch.randn(E, D, H, generator=g) * 0.4, torch.randn(E, H, D, generator=g) * 0.3)


def _tokens(rank):
    g = torch.Generator().manual_seed(40 + rank)
    return torch.randn(T, D, generator=g), torch.randn(T, D, generator=g)


def _dense(x, gate, w1, w2, k):
    probs = torch.softmax(x @ gate.t(), -1)
    w, e = torch.topk(probs, k, -1)
    w = w / w.sum(-1, keepdim=True)
    y = torch.zeros_like(x)
    for j in range(k):
        for ex in range(E):
            m = e[:, j] == ex
            if m.any():
                y = y + torch.zeros_like(x).index_add(0, m.nonzero().view(-1), (torch.relu(x[m] @ w1[ex]) @ w2[ex]) * w[m, j:j + 1])
    return y


@pytest.mark.parametrize("world,k", [(1, 1), (2, 1), (2, 2), (4, 2)])
def test_moe_matches_dense(world, k):
    gate0, w10, w20 = _weights()
    # reference: every rank's tokens through the dense layer; expert / gate gradients summed over the ranks' losses
    gate, w1, w2 = (t.clone().requires_grad_(True) for t in (gate0, w10, w20))
    want_y, want_gx = [], []
    for r in range(world):
        x, t = _tokens(r)
        x = x.clone().requires_grad_(True)
        y = _dense(x, gate, w1, w2, k)
        ((y - t) ** 2).sum().backward()
        want_y.append(y.detach())
        want_gx.append(x.grad.clone())

    def body(r, mlsl):
        from mlsl_b200.parallel.expert_parallel import ExpertParallelMoE
        dist = mlsl.env().create_distribution(1, world)
        moe = ExpertParallelMoE(D, H, E, top_k=k, group="model", distribution=dist)
        El = E // world
        with torch.no_grad():
            moe.gate.weight.copy_(gate0)
            moe.w1.copy_(w10[r * El:(r + 1) * El])
            moe.w2.copy_(w20[r * El:(r + 1) * El])
        x, t = _tokens(r)
        x = x.clone().requires_grad_(True)
        y = moe(x)
        ((y - t) ** 2).sum().backward()
        ggate = moe.gate.weight.grad.clone().contiguous()
        mlsl.allreduce(ggate.view(-1), group="model", distribution=dist)       # partial per rank -> sum
        mlsl.env().delete_distribution(dist)
        return y.detach(), x.grad.clone(), moe.w1.grad.clone(), moe.w2.grad.clone(), ggate

    res = run_ranks(world, body)
    El = E // world
    for r in range(world):
        y, gx, gw1, gw2, gg = res[r]
        assert torch.allclose(y, want_y[r], atol=1e-5, rtol=1e-4), (r, (y - want_y[r]).abs().max())
        assert torch.allclose(gx, want_gx[r], atol=1e-5, rtol=1e-4)
        assert torch.allclose(gw1, w1.grad[r * El:(r + 1) * El], atol=1e-5, rtol=1e-4)
        assert torch.allclose(gw2, w2.grad[r * El:(r + 1) * El], atol=1e-5, rtol=1e-4)
        assert torch.allclose(gg, gate.grad, atol=1e-5, rtol=1e-4)


def test_alltoallv_tensor_api():
    def body(r, mlsl):
        P = mlsl.world_size()
        send = [(r + p) % 3 for p in range(P)]
        x = torch.cat([torch.full((send[p],), float(100 * r + p)) for p in range(P)] + [torch.empty(0)])
        out = mlsl.alltoallv(x, send, group="global")
        want = torch.cat([torch.full(((p + r) % 3,), float(100 * p + r)) for p in range(P)] + [torch.empty(0)])
        assert torch.equal(out, want), (r, out, want)
        with pytest.raises(ValueError):
            mlsl.alltoallv(x, send[:-1], group="global")
        return True

    assert run_ranks(3, body) == [True] * 3
