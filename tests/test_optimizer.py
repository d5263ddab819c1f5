"""DistributedOptimizer (Session/ParameterSet graph API under the hood) vs a single-process torch optimizer fed with
the averaged gradient.  CPU (host backend) always; the same check on the CUDA kernels is marked gpu."""
import pytest
import torch

from conftest import run_ranks


import threading

_init_lock = threading.Lock()


def _model(dtype=torch.float32):
    # the global RNG is shared by the in-process rank threads: seed + construct atomically
    with _init_lock:
        torch.manual_seed(7)
        m = torch.nn.Sequential(torch.nn.Linear(24, 48), torch.nn.Tanh(), torch.nn.Linear(48, 48), torch.nn.Tanh(),
                                torch.nn.Linear(48, 5))
    return m.to(dtype)


def _batch(rank, step):
    g = torch.Generator().manual_seed(1000 * step + rank)
    return torch.randn(6, 24, generator=g), torch.randn(6, 5, generator=g)


def _reference(world, steps, kind, kw):
    m = _model()
    opt = (torch.optim.AdamW(m.parameters(), **kw) if kind == "adamw" else torch.optim.SGD(m.parameters(), **kw))
    for s in range(steps):
        opt.zero_grad()
        loss = 0
        for r in range(world):
            x, y = _batch(r, s)
            loss = loss + torch.nn.functional.mse_loss(m(x), y) / world
        loss.backward()
        opt.step()
    return torch.cat([p.detach().reshape(-1) for p in m.parameters()])


def _train(world, steps, kind, mode, backend, kw, bucket_mb=0.004, compress=False):
    def body(r, mlsl):
        dev = "cuda" if backend == "cuda" else "cpu"
        m = _model().to(dev)
        okw = dict(lr=kw["lr"], weight_decay=kw.get("weight_decay", 0.0), optimizer=kind, mode=mode, bucket_mb=bucket_mb,
                   compress=compress)
        if kind == "sgd":
            okw["momentum"] = kw.get("momentum", 0.0)
        opt = mlsl.DistributedOptimizer(m.parameters(), **okw)
        assert len(opt.buckets) > 1            # several buckets: exercises per-bucket overlap bookkeeping
        for s in range(steps):
            opt.zero_grad()
            x, y = _batch(r, s)
            torch.nn.functional.mse_loss(m(x.to(dev)), y.to(dev)).backward()
            opt.step()
        if backend == "cuda":
            torch.cuda.current_stream().synchronize()
        out = torch.cat([p.detach().reshape(-1).float().cpu() for p in m.parameters()])
        opt.close()
        return out

    env = {"MLSL_HEAP_SIZE_GB": "0.25"} if backend == "cuda" else None
    return run_ranks(world, body, backend=backend, env=env)


@pytest.mark.parametrize("mode", ["fused", "allreduce"])
@pytest.mark.parametrize("kind,kw", [("sgd", dict(lr=0.05, momentum=0.9, weight_decay=0.01)),
                                     ("adamw", dict(lr=0.01, weight_decay=0.02))])
def test_distributed_optimizer_cpu(mode, kind, kw):
    world, steps = 3, 4
    outs = _train(world, steps, kind, mode, "host", kw)
    ref = _reference(world, steps, kind, kw)
    for o in outs:
        assert torch.allclose(o, ref, rtol=2e-4, atol=2e-5), (o - ref).abs().max()
        assert torch.equal(o, outs[0])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fused", "allreduce"])
@pytest.mark.parametrize("kind,kw", [("sgd", dict(lr=0.05, momentum=0.9, weight_decay=0.01)),
                                     ("adamw", dict(lr=0.01, weight_decay=0.02))])
def test_distributed_optimizer_gpu(mode, kind, kw):
    world, steps = 4, 4
    outs = _train(world, steps, kind, mode, "cuda", kw)
    ref = _reference(world, steps, kind, kw)
    for o in outs:
        assert torch.allclose(o, ref, rtol=5e-4, atol=5e-5), (o - ref).abs().max()
        assert torch.equal(o, outs[0])


@pytest.mark.gpu
def test_distributed_optimizer_gpu_fp8_gradients():
    world, steps = 4, 3
    kw = dict(lr=0.05, momentum=0.9)
    outs = _train(world, steps, "sgd", "allreduce", "cuda", kw, compress=True)
    ref = _reference(world, steps, "sgd", kw)
    for o in outs:
        assert (o - ref).abs().max() < 0.05        # fp8 transport: close, not equal
        assert torch.equal(o, outs[0])
