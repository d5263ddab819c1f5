"""Sequence-parallel MLP block on the CUDA backend: fc1 = all-gather + matmul (the fused kernel stays opt-in), fc2 forward
and fc1 backward on the tcgen05 GEMM + reduce-scatter kernel, bf16.  Last-sorted file: first hardware run at round end."""
import threading

import pytest
import torch

from conftest import run_ranks

pytestmark = pytest.mark.gpu
_lock = threading.Lock()
D_IN, D_HID, D_OUT, M = 256, 512, 256, 512      # shapes the fused GEMM + reduce-scatter accepts for 2 ranks


def _full():
    g = torch.Generator().manual_seed(9)
    w1 = torch.randn(D_HID, D_IN, generator=g) * 0.05
    w2 = torch.randn(D_OUT, D_HID, generator=g) * 0.05
    x = torch.randn(M, D_IN, generator=g)
    t = torch.randn(M, D_OUT, generator=g)
    return w1, w2, x, t


def test_sequence_parallel_mlp_block_device():
    world = 2

    def body(r, mlsl):
        from mlsl_b200.parallel.tensor_parallel import ColumnParallelLinear, RowParallelLinear
        w1, w2, x, t = _full()
        e = mlsl.env()
        dist = e.create_distribution(1, world)
        with _lock:
            col = ColumnParallelLinear(D_IN, D_HID, bias=False, distribution=dist, sequence_parallel=True,
                                       dtype=torch.bfloat16, device="cuda")
            row = RowParallelLinear(D_HID, D_OUT, bias=False, distribution=dist, dtype=torch.bfloat16, device="cuda")
        hs, rows = D_HID // world, M // world
        with torch.no_grad():
            col.weight.copy_(w1[r * hs:(r + 1) * hs].to(torch.bfloat16))
            row.weight.copy_(w2[:, r * hs:(r + 1) * hs].to(torch.bfloat16))
        xin = x[r * rows:(r + 1) * rows].cuda().to(torch.bfloat16).requires_grad_(True)
        y = row(torch.relu(col(xin)))
        loss = ((y.float() - t[r * rows:(r + 1) * rows].cuda()) ** 2).sum() / (M * D_OUT)
        loss.backward()
        torch.cuda.current_stream().synchronize()
        out = (y.detach().float().cpu(), xin.grad.float().cpu(), col.weight.grad.float().cpu(), row.weight.grad.float().cpu())
        e.delete_distribution(dist)
        return out

    outs = run_ranks(world, body, backend="cuda", env={"MLSL_HEAP_SIZE_GB": "0.5", "MLSL_WATCHDOG_SEC": "20"})
    w1, w2, x, t = _full()
    w1r, w2r, xr = (v.to(torch.bfloat16).float().requires_grad_(True) for v in (w1, w2, x))
    yr = torch.relu(xr @ w1r.t()) @ w2r.t()
    ((yr - t) ** 2).mean().backward()
    hs, rows = D_HID // world, M // world
    tol = 4e-2

    def close(a, b):
        return (a - b).abs().max().item() <= tol * max(1e-3, b.abs().max().item())

    for r, (y, gx, gw1, gw2) in enumerate(outs):
        assert close(y, yr.detach()[r * rows:(r + 1) * rows])
        assert close(gx, xr.grad[r * rows:(r + 1) * rows])
        assert close(gw1, w1r.grad[r * hs:(r + 1) * hs])
        assert close(gw2, w2r.grad[:, r * hs:(r + 1) * hs])
