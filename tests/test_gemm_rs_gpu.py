"""tcgen05 GEMM fused with the reduce-scatter of its partial sums vs a plain PyTorch fp32 reference."""
import pytest
import torch

from conftest import run_ranks

pytestmark = pytest.mark.gpu


def _mats(rank, M, N, K):
    g = torch.Generator().manual_seed(77 + rank)
    a = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.5).to(torch.bfloat16)
    return a, w


@pytest.mark.parametrize("world,M,N,K", [(1, 128, 256, 64), (1, 256, 768, 512), (2, 512, 256, 256), (4, 1024, 512, 320),
                                         (2, 2048, 1024, 1024)])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_gemm_reduce_scatter(world, M, N, K, out_dtype):
    def body(r, mlsl):
        from mlsl_b200.ops import gemm_reduce_scatter
        a, w = _mats(r, M, N, K)
        a, w = a.cuda(), w.cuda()
        outs = []
        for _ in range(2):   # twice: accumulator double-buffering / staging reuse across launches
            out = gemm_reduce_scatter(a, w, out_dtype=out_dtype, group="global")
            torch.cuda.current_stream().synchronize()
            outs.append(out.float().cpu())
        assert torch.equal(outs[0], outs[1])
        return outs[0]

    outs = run_ranks(world, body, backend="cuda", env={"MLSL_HEAP_SIZE_GB": "0.5", "MLSL_WATCHDOG_SEC": "20"}, wait_mode="host")
    ref = torch.zeros(M, N, dtype=torch.float32)
    for r in range(world):
        a, w = _mats(r, M, N, K)
        # partials travel as bf16, exactly like a bf16 GEMM followed by a bf16 reduce-scatter
        ref += (a.float() @ w.float().t()).to(torch.bfloat16).float()
    rows = M // world
    for r, o in enumerate(outs):
        exp = ref[r * rows:(r + 1) * rows]
        tol = 2e-2 * max(1.0, exp.abs().max().item())
        assert o.shape == exp.shape
        assert (o - exp).abs().max().item() <= tol, (o - exp).abs().max().item()


@pytest.mark.parametrize("world,M,N,K", [(1, 256, 256, 64), (1, 512, 768, 512), (2, 1024, 512, 384), (2, 2048, 1024, 1024)])
def test_gemm_reduce_scatter_two_cta(world, M, N, K):
    """Same check for the CTA-pair kernel (tcgen05.mma.cta_group::2, 256 x 256 tiles, MLSL_GEMM_2CTA=1)."""
    def body(r, mlsl):
        from mlsl_b200.ops import gemm_reduce_scatter
        a, w = _mats(r, M, N, K)
        a, w = a.cuda(), w.cuda()
        outs = []
        for _ in range(2):
            out = gemm_reduce_scatter(a, w, out_dtype=torch.float32, group="global")
            torch.cuda.current_stream().synchronize()
            outs.append(out.float().cpu())
        assert torch.equal(outs[0], outs[1])
        return outs[0]

    outs = run_ranks(world, body, backend="cuda", env={"MLSL_HEAP_SIZE_GB": "0.5", "MLSL_WATCHDOG_SEC": "20", "MLSL_GEMM_2CTA": "1"}, wait_mode="host")
    ref = torch.zeros(M, N, dtype=torch.float32)
    for r in range(world):
        a, w = _mats(r, M, N, K)
        ref += (a.float() @ w.float().t()).to(torch.bfloat16).float()
    rows = M // world
    for r, o in enumerate(outs):
        exp = ref[r * rows:(r + 1) * rows]
        assert (o - exp).abs().max().item() <= 2e-2 * max(1.0, exp.abs().max().item())
