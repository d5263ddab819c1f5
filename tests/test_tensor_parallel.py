"""Column/Row-parallel linear layers (model parallelism of the reference, activation exchange case 1) vs a single
process MLP: forward values and all gradients.  CPU (host backend) always; CUDA (fused GEMM+RS forward) marked gpu."""
import threading

import pytest
import torch

from conftest import run_ranks

_lock = threading.Lock()
D_IN, D_HID, D_OUT, M = 64, 512, 256, 256


def _full_weights():
    g = torch.Generator().manual_seed(5)
    w1 = torch.randn(D_HID, D_IN, generator=g) * 0.1
    w2 = torch.randn(D_OUT, D_HID, generator=g) * 0.1
    x = torch.randn(M, D_IN, generator=g)
    t = torch.randn(M, D_OUT, generator=g)
    return w1, w2, x, t


def _run(world, backend, dtype, tol):
    def body(r, mlsl):
        from mlsl_b200.parallel.tensor_parallel import ColumnParallelLinear, RowParallelLinear, gather_rows
        dev = "cuda" if backend == "cuda" else "cpu"
        w1, w2, x, t = _full_weights()
        e = mlsl.env()
        dist = e.create_distribution(1, world)          # pure model parallelism
        with _lock:
            col = ColumnParallelLinear(D_IN, D_HID, bias=False, distribution=dist, dtype=dtype, device=dev)
            row = RowParallelLinear(D_HID, D_OUT, bias=False, distribution=dist, dtype=dtype, device=dev)
        hs = D_HID // world
        with torch.no_grad():
            col.weight.copy_(w1[r * hs:(r + 1) * hs].to(dtype))
            row.weight.copy_(w2[:, r * hs:(r + 1) * hs].to(dtype))
        xin = x.to(dev).to(dtype).requires_grad_(True)
        h = torch.relu(col(xin))
        y_rows = row(h)                                  # [M / world, D_OUT]
        y = gather_rows(y_rows, distribution=dist)       # [M, D_OUT]
        loss = ((y.float() - t.to(dev)) ** 2).mean()
        loss.backward()
        if backend == "cuda":
            torch.cuda.current_stream().synchronize()
        out = (y.detach().float().cpu(), xin.grad.float().cpu(), col.weight.grad.float().cpu(), row.weight.grad.float().cpu())
        e.delete_distribution(dist)
        return out

    env = {"MLSL_HEAP_SIZE_GB": "0.5", "MLSL_WATCHDOG_SEC": "20"} if backend == "cuda" else None
    outs = run_ranks(world, body, backend=backend, env=env)
    w1, w2, x, t = _full_weights()
    w1r, w2r, xr = (v.to(dtype).float().requires_grad_(True) for v in (w1, w2, x))
    yr = torch.relu(xr @ w1r.t()) @ w2r.t()
    ((yr - t) ** 2).mean().backward()
    hs = D_HID // world
    for r, (y, gx, gw1, gw2) in enumerate(outs):
        s = max(1.0, yr.abs().max().item())
        assert (y - yr.detach()).abs().max().item() <= tol * s
        assert (gx - xr.grad).abs().max().item() <= tol * max(1e-3, xr.grad.abs().max().item()) * 4
        assert (gw1 - w1r.grad[r * hs:(r + 1) * hs]).abs().max().item() <= tol * max(1e-3, w1r.grad.abs().max().item()) * 4
        assert (gw2 - w2r.grad[:, r * hs:(r + 1) * hs]).abs().max().item() <= tol * max(1e-3, w2r.grad.abs().max().item()) * 4


@pytest.mark.parametrize("world", [1, 2, 4])
def test_tensor_parallel_mlp_cpu(world):
    _run(world, "host", torch.float32, 1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2])
def test_tensor_parallel_mlp_gpu_fused_gemm_rs(world):
    _run(world, "cuda", torch.bfloat16, 3e-2)


def _ag_gemm_case(world, backend, fused, dtype, tol, M=512, N=256, K=128):
    def body(r, mlsl):
        from mlsl_b200.ops import allgather_gemm
        dev = "cuda" if backend == "cuda" else "cpu"
        e = mlsl.env()
        dist = e.create_distribution(1, world)
        g = torch.Generator().manual_seed(31)
        x = (torch.randn(M, K, generator=g) * 0.5).to(dtype)
        w = (torch.randn(N, K, generator=g) * 0.5).to(dtype)
        rows = M // world
        outs = []
        for _ in range(2):      # twice: flag epochs / staging reuse
            y, full = allgather_gemm(x[r * rows:(r + 1) * rows].to(dev), w.to(dev), out_dtype=torch.float32,
                                     group="model", distribution=dist, fused=fused)
            if backend == "cuda":
                torch.cuda.current_stream().synchronize()
            outs.append((y.float().cpu(), full.float().cpu()))
        e.delete_distribution(dist)
        assert torch.equal(outs[0][0], outs[1][0])
        return outs[0]

    env = {"MLSL_HEAP_SIZE_GB": "0.5", "MLSL_WATCHDOG_SEC": "20"} if backend == "cuda" else None
    outs = run_ranks(world, body, backend=backend, env=env)
    g = torch.Generator().manual_seed(31)
    x = (torch.randn(M, K, generator=g) * 0.5).to(dtype).float()
    w = (torch.randn(N, K, generator=g) * 0.5).to(dtype).float()
    ref = x @ w.t()
    for y, full in outs:
        assert torch.equal(full, x)
        assert (y - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("world", [1, 2, 4])
def test_allgather_gemm_unfused_cpu(world):
    _ag_gemm_case(world, "host", False, torch.float32, 1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("world,M,N,K", [(1, 256, 256, 64), (2, 512, 512, 256), (4, 1024, 768, 512)])
def test_allgather_gemm_fused_gpu(world, M, N, K):
    _ag_gemm_case(world, "cuda", True, torch.bfloat16, 2e-2, M, N, K)


@pytest.mark.parametrize("world", [1, 2, 4])
def test_sequence_parallel_mlp_block_cpu(world):
    """Megatron-SP MLP: tokens split over the model group at both ends ([M/P, D] -> [M/P, D_OUT]); fc1 = all-gather +
    GEMM, fc2 = GEMM + reduce-scatter; values and every gradient against the single-process MLP."""
    def body(r, mlsl):
        from mlsl_b200.parallel.tensor_parallel import ColumnParallelLinear, RowParallelLinear
        w1, w2, x, t = _full_weights()
        e = mlsl.env()
        dist = e.create_distribution(1, world)
        with _lock:
            col = ColumnParallelLinear(D_IN, D_HID, bias=False, distribution=dist, sequence_parallel=True)
            row = RowParallelLinear(D_HID, D_OUT, bias=False, distribution=dist)
        hs, rows = D_HID // world, M // world
        with torch.no_grad():
            col.weight.copy_(w1[r * hs:(r + 1) * hs])
            row.weight.copy_(w2[:, r * hs:(r + 1) * hs])
        xin = x[r * rows:(r + 1) * rows].clone().requires_grad_(True)
        y = row(torch.relu(col(xin)))                      # [M/P, D_OUT]
        loss = ((y - t[r * rows:(r + 1) * rows]) ** 2).sum() / (M * D_OUT)
        loss.backward()
        out = (y.detach(), xin.grad, col.weight.grad, row.weight.grad)
        e.delete_distribution(dist)
        return out

    outs = run_ranks(world, body)
    w1, w2, x, t = _full_weights()
    w1r, w2r, xr = (v.clone().requires_grad_(True) for v in (w1, w2, x))
    yr = torch.relu(xr @ w1r.t()) @ w2r.t()
    ((yr - t) ** 2).mean().backward()
    hs, rows = D_HID // world, M // world
    for r, (y, gx, gw1, gw2) in enumerate(outs):
        assert torch.allclose(y, yr.detach()[r * rows:(r + 1) * rows], rtol=1e-4, atol=1e-5)
        assert torch.allclose(gx, xr.grad[r * rows:(r + 1) * rows], rtol=1e-4, atol=1e-6)
        assert torch.allclose(gw1, w1r.grad[r * hs:(r + 1) * hs], rtol=1e-4, atol=1e-6)
        assert torch.allclose(gw2, w2r.grad[:, r * hs:(r + 1) * hs], rtol=1e-4, atol=1e-6)
