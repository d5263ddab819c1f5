"""CUDA twins of the pipeline- and expert-parallel tests (device tensors, loop-back ranks on one GPU)."""
import os
import threading

import pytest
import torch

from conftest import run_ranks

pytestmark = pytest.mark.gpu
# pipeline and expert parallelism run here by default (validated on hardware in round 2); the transformer block is
# torch-heavy (cuBLASLt workspaces, many first-use kernels per thread) and stays opt-in in loop-back mode - it runs one rank
# per GPU in examples/ and tests/mp_gpu_check.py
_opt_in = pytest.mark.skipif(os.environ.get("MLSL_TEST_STRATEGIES_GPU") != "1", reason="loop-back opt-in (torch-heavy)")
_lock = threading.Lock()
ENV = {"MLSL_HEAP_SIZE_GB": "0.5", "MLSL_WATCHDOG_SEC": "20"}


def test_pipeline_parallel_device():
    from test_pipeline_parallel import D_IN, MB, _blocks, _data, _reference
    stages, micro = 3, 4
    want_loss, want_grads = _reference(stages, micro)

    def body(r, mlsl):
        from mlsl_b200.parallel.pipeline_parallel import PipelineStage
        dist = mlsl.env().create_distribution(1, stages)
        with _lock:
            block = _blocks(stages)[r].cuda()
        st = PipelineStage(block, (MB, D_IN), (MB, D_IN), group="model", distribution=dist)
        xs, ys = _data(micro)
        xs, ys = [x.cuda() for x in xs], [y.cuda() for y in ys]
        loss = st.step(xs if st.is_first else None, loss_fn=torch.nn.functional.mse_loss if st.is_last else None,
                       targets=ys if st.is_last else None, num_micro=micro)
        torch.cuda.current_stream().synchronize()
        mlsl.env().delete_distribution(dist)
        return (loss.item() if loss is not None else None), [p.grad.float().cpu() for p in block.parameters()]

    res = run_ranks(stages, body, backend="cuda", env=ENV)
    assert abs(res[stages - 1][0] - want_loss) < 1e-4
    for r in range(stages):
        for got, want in zip(res[r][1], want_grads[r]):
            assert torch.allclose(got, want, atol=1e-4, rtol=1e-3), (r, (got - want).abs().max())


def test_expert_parallel_device():
    from test_expert_parallel import D, E, H, _dense, _tokens, _weights
    world, k = 2, 2
    gate0, w10, w20 = _weights()
    gate, w1, w2 = (t.clone().requires_grad_(True) for t in (gate0, w10, w20))
    want = []
    for r in range(world):
        x, t = _tokens(r)
        x = x.clone().requires_grad_(True)
        y = _dense(x, gate, w1, w2, k)
        ((y - t) ** 2).sum().backward()
        want.append((y.detach(), x.grad.clone()))

    def body(r, mlsl):
        from mlsl_b200.parallel.expert_parallel import ExpertParallelMoE
        dist = mlsl.env().create_distribution(1, world)
        with _lock:
            moe = ExpertParallelMoE(D, H, E, top_k=k, group="model", distribution=dist, device="cuda")
        El = E // world
        with torch.no_grad():
            moe.gate.weight.copy_(gate0)
            moe.w1.copy_(w10[r * El:(r + 1) * El])
            moe.w2.copy_(w20[r * El:(r + 1) * El])
        x, t = _tokens(r)
        x = x.cuda().requires_grad_(True)
        y = moe(x)
        ((y - t.cuda()) ** 2).sum().backward()
        torch.cuda.current_stream().synchronize()
        mlsl.env().delete_distribution(dist)
        return y.detach().cpu(), x.grad.cpu(), moe.w1.grad.cpu(), moe.w2.grad.cpu()

    res = run_ranks(world, body, backend="cuda", env=ENV)
    El = E // world
    for r in range(world):
        y, gx, gw1, gw2 = res[r]
        assert torch.allclose(y, want[r][0], atol=1e-3, rtol=1e-2)      # TF32-free fp32 matmuls, different summation order
        assert torch.allclose(gx, want[r][1], atol=1e-3, rtol=1e-2)
        assert torch.allclose(gw1, w1.grad[r * El:(r + 1) * El], atol=1e-3, rtol=1e-2)
        assert torch.allclose(gw2, w2.grad[r * El:(r + 1) * El], atol=1e-3, rtol=1e-2)


@_opt_in
def test_parallel_transformer_block_device():
    """bf16 block on the fused kernels' shapes (GEMM + reduce-scatter always; all-gather + GEMM with MLSL_AG_GEMM=1)."""
    from test_parallel_transformer import _full, _reference
    D, H, FF, M, world = 512, 8, 2048, 512, 2
    want_y, want = _reference(D, H, FF, M, cast=torch.bfloat16)

    def body(r, mlsl):
        from mlsl_b200.models.gpt import ParallelTransformerBlock
        f = {k: v.cuda() for k, v in _full(D, H, FF, M).items()}
        dist = mlsl.env().create_distribution(1, world)
        with _lock:
            blk = ParallelTransformerBlock(D, H, d_ff=FF, distribution=dist, group="model", dtype=torch.bfloat16, device="cuda")
        dl, fl, rows = D // world, FF // world, M // world
        sl, fs, rs = slice(r * dl, (r + 1) * dl), slice(r * fl, (r + 1) * fl), slice(r * rows, (r + 1) * rows)
        with torch.no_grad():
            blk.ln1.weight.copy_(f["ln1w"]), blk.ln1.bias.copy_(f["ln1b"])
            blk.ln2.weight.copy_(f["ln2w"]), blk.ln2.bias.copy_(f["ln2b"])
            blk.qkv.weight.copy_(torch.cat([f["wq"][sl], f["wk"][sl], f["wv"][sl]]))
            blk.qkv.bias.copy_(torch.cat([f["bq"][sl], f["bk"][sl], f["bv"][sl]]))
            blk.proj.weight.copy_(f["wo"][:, sl]), blk.proj.bias.copy_(f["bo"])
            blk.fc1.weight.copy_(f["w1"][fs]), blk.fc1.bias.copy_(f["b1"][fs])
            blk.fc2.weight.copy_(f["w2"][:, fs]), blk.fc2.bias.copy_(f["b2"])
        x = f["x"][rs].to(torch.bfloat16).requires_grad_(True)
        y = blk(x)
        ((y.float() - f["t"][rs]) ** 2).sum().backward()
        torch.cuda.current_stream().synchronize()
        mlsl.env().delete_distribution(dist)
        return y.detach().float().cpu(), x.grad.float().cpu(), blk.fc2.weight.grad.float().cpu()

    res = run_ranks(world, body, backend="cuda", env=ENV)
    for r, (y, gx, gw2) in enumerate(res):
        rows, fl = M // world, FF // world
        for got, ref in ((y, want_y[r * rows:(r + 1) * rows]), (gx, want["x"][r * rows:(r + 1) * rows]),
                         (gw2, want["w2"][:, r * fl:(r + 1) * fl])):
            assert (got - ref).abs().max().item() <= 6e-2 * max(1.0, ref.abs().max().item())
