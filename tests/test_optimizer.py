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


def _train(world, steps, kind, mode, backend, kw, bucket_mb=0.004, compress=False, model_zero_grad=False, dtype=torch.float32):
    def body(r, mlsl):
        dev = "cuda" if backend == "cuda" else "cpu"
        m = _model(dtype).to(dev)
        okw = dict(lr=kw["lr"], weight_decay=kw.get("weight_decay", 0.0), optimizer=kind, mode=mode, bucket_mb=bucket_mb,
                   compress=compress)
        if kind == "sgd":
            okw["momentum"] = kw.get("momentum", 0.0)
        opt = mlsl.DistributedOptimizer(m.parameters(), **okw)
        assert len(opt.buckets) > 1            # several buckets: exercises per-bucket overlap bookkeeping
        for s in range(steps):
            if model_zero_grad:
                m.zero_grad()                  # set_to_none=True: autograd allocates fresh gradients outside the buckets
            else:
                opt.zero_grad()
            x, y = _batch(r, s)
            torch.nn.functional.mse_loss(m(x.to(dev).to(dtype)).float(), y.to(dev)).backward()
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


@pytest.mark.parametrize("mode,kind,kw", [("fused", "adamw", dict(lr=0.01, weight_decay=0.02)),
                                          ("fused", "sgd", dict(lr=0.05, momentum=0.9, weight_decay=0.01)),
                                          ("allreduce", "sgd", dict(lr=0.05, momentum=0.9, weight_decay=0.01))])
def test_bf16_parameters_and_gradients(mode, kind, kw):
    """bf16 parameters: the gradient buckets travel in bf16 (half the bytes); the fused sharded optimizer keeps fp32 master
    weights and moments for its shard, so the result tracks the fp32 reference to bf16 resolution and replicas stay identical."""
    world, steps = 2, 4
    outs = _train(world, steps, kind, mode, "host", kw, dtype=torch.bfloat16)
    ref = _reference(world, steps, kind, kw)
    for o in outs:
        assert torch.equal(o, outs[0])
        assert torch.allclose(o, ref, rtol=0.05, atol=0.02), (o - ref).abs().max()


@pytest.mark.parametrize("mode", ["fused", "allreduce"])
def test_model_zero_grad_set_to_none_keeps_replicas_in_sync(mode):
    """The usual `model.zero_grad()` detaches p.grad from the bucket; the hooks move the fresh gradients back in."""
    kw = dict(lr=0.05, momentum=0.9, weight_decay=0.01)
    ref = _reference(2, 3, "sgd", kw)
    for out in _train(2, 3, "sgd", mode, "host", kw, model_zero_grad=True):
        assert torch.allclose(out, ref, rtol=1e-4, atol=1e-5)


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


@pytest.mark.parametrize("mode,kind,kw", [("fused", "adamw", dict(lr=1e-2, weight_decay=0.01)),
                                          ("fused", "sgd", dict(lr=0.05, momentum=0.9)),
                                          ("allreduce", "adamw", dict(lr=1e-2, weight_decay=0.01))])
def test_checkpoint_resume_and_reshard(mode, kind, kw):
    """state_dict() after 3 steps, 2 more steps -> reference.  A fresh optimizer that loads the checkpoint and runs the
    same 2 steps must land on the same weights - also when the job is resumed on a different number of ranks (every
    rank feeds the same batch here, so the averaged gradient does not depend on the world size)."""
    def make(mlsl, m):
        okw = dict(lr=kw["lr"], weight_decay=kw.get("weight_decay", 0.0), optimizer=kind, mode=mode, bucket_mb=0.004)
        if kind == "sgd":
            okw["momentum"] = kw.get("momentum", 0.0)
        return mlsl.DistributedOptimizer(m.parameters(), **okw)

    def step(m, opt, s):
        opt.zero_grad()
        x, y = _batch(0, s)
        torch.nn.functional.mse_loss(m(x), y).backward()
        opt.step()

    def flat(m):
        return torch.cat([p.detach().reshape(-1).float() for p in m.parameters()])

    saved = {}

    def first(r, mlsl):
        m = _model()
        opt = make(mlsl, m)
        for s in range(3):
            step(m, opt, s)
        sd = opt.state_dict()
        if r == 0:
            saved["sd"] = sd
        for s in range(3, 5):
            step(m, opt, s)
        out = flat(m)
        opt.close()
        return out

    ref = run_ranks(2, first)[0]

    def resumed(r, mlsl):
        m = _model()
        with torch.no_grad():
            for p in m.parameters():
                p.add_(1.0)                       # make sure the weights really come from the checkpoint
        opt = make(mlsl, m)
        opt.load_state_dict(saved["sd"])
        assert opt.steps == 3
        for s in range(3, 5):
            step(m, opt, s)
        out = flat(m)
        opt.close()
        return out

    same_world = run_ranks(2, resumed)
    assert torch.equal(same_world[0], ref) and torch.equal(same_world[1], ref)
    other_world = run_ranks(4, resumed)
    for o in other_world:
        assert torch.allclose(o, ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("mode", ["fused", "allreduce"])
def test_gradient_accumulation_with_no_sync(mode):
    """Two micro-batches per step: the first backward runs under no_sync() (no exchange, the bucket only accumulates), the
    second starts the exchange of the sum.  Reference: one process, the mean over ranks of the summed micro-batch gradients."""
    world, steps, kw = 2, 3, dict(lr=0.05, momentum=0.9)

    def body(r, mlsl):
        m = _model()
        opt = mlsl.DistributedOptimizer(m.parameters(), lr=kw["lr"], momentum=kw["momentum"], mode=mode, bucket_mb=0.004)
        started = []
        orig = opt._start
        opt._start = lambda b: (started.append(1), orig(b))[1]
        for s in range(steps):
            opt.zero_grad()
            with opt.no_sync():
                x, y = _batch(r, 2 * s)
                torch.nn.functional.mse_loss(m(x), y).backward()
                assert not started, "no_sync() must not start any exchange"
            x, y = _batch(r, 2 * s + 1)
            torch.nn.functional.mse_loss(m(x), y).backward()
            assert len(started) == len(opt.buckets)
            opt.step()
            started.clear()
        out = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
        opt.close()
        return out

    outs = run_ranks(world, body)
    m = _model()
    ref_opt = torch.optim.SGD(m.parameters(), **kw)
    for s in range(steps):
        ref_opt.zero_grad()
        loss = 0
        for r in range(world):
            for micro in (2 * s, 2 * s + 1):
                x, y = _batch(r, micro)
                loss = loss + torch.nn.functional.mse_loss(m(x), y) / world
        loss.backward()
        ref_opt.step()
    ref = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
    for o in outs:
        assert torch.allclose(o, ref, rtol=2e-4, atol=2e-5), (o - ref).abs().max()


def test_model_stays_usable_after_close():
    """close() frees the buckets the parameters were views of: they must own their values again, and a second optimizer on
    the same model must continue from them."""
    def body(r, mlsl):
        m = _model()
        opt = mlsl.DistributedOptimizer(m.parameters(), lr=0.05, mode="fused", bucket_mb=0.004)
        x, y = _batch(r, 0)
        opt.zero_grad()
        torch.nn.functional.mse_loss(m(x), y).backward()
        opt.step()
        before = torch.cat([p.detach().reshape(-1).clone() for p in m.parameters()])
        opt.close()
        junk = [mlsl.alloc_tensor(4096, torch.float32) for _ in range(8)]      # recycle the freed blocks
        for j in junk:
            j.fill_(123.0)
        after = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
        assert torch.equal(before, after) and all(p.grad is None for p in m.parameters())
        opt2 = mlsl.DistributedOptimizer(m.parameters(), lr=0.05, mode="allreduce", bucket_mb=0.004)
        opt2.zero_grad()
        torch.nn.functional.mse_loss(m(x), y).backward()
        opt2.step()
        out = torch.cat([p.detach().reshape(-1).clone() for p in m.parameters()])
        opt2.close()
        for j in junk:
            mlsl.free_tensor(j)
        return out

    outs = run_ranks(2, body)
    assert torch.equal(outs[0], outs[1]) and torch.isfinite(outs[0]).all()
