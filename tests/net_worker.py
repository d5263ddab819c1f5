"""One rank of a multi-node (net backend, TCP) job: the randomised collective program of test_fuzz_cpu.py, a fused
distributed update and a failure-free finalize.  Started by bin/mlslrun --nnodes ... from test_net_backend_cpu.py."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mlsl_b200 as mlsl  # noqa: E402
from test_fuzz_cpu import check_rank, make_program, run_program  # noqa: E402


def main():
    seed = int(sys.argv[1])
    mlsl.init()
    r, world = mlsl.rank(), mlsl.world_size()
    assert mlsl.env().get_backend_name() == "net", mlsl.env().describe_backend()
    D, M, program = make_program(world, seed)
    results = run_program(r, mlsl, world, D, M, program)
    check_rank(r, world, D, M, program, results)
    # sharded optimizer over the wire: two steps of AdamW must leave every replica with identical weights
    torch.manual_seed(5)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
    opt = mlsl.DistributedOptimizer(model.parameters(), lr=1e-2, optimizer="adamw", mode="fused", bucket_mb=0.002)
    g = torch.Generator().manual_seed(100 + r)
    for _ in range(2):
        opt.zero_grad()
        torch.nn.functional.mse_loss(model(torch.randn(8, 16, generator=g)), torch.randn(8, 4, generator=g)).backward()
        opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).contiguous()
    ref = flat.clone()
    mlsl.bcast(ref, root=0)
    assert torch.equal(flat, ref)
    opt.close()
    # bf16 parameters and gradients through the sharded optimizer (fp32 master weights in the owner's shard; with
    # MLSL_NET_HIER_KB=0 the two-level route): replicas identical, and close to plain fp32 AdamW on the averaged gradient
    torch.manual_seed(7)
    ref_model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
    model_h = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
    model_h.load_state_dict(ref_model.state_dict())
    model_h = model_h.to(torch.bfloat16)
    ref_opt = torch.optim.AdamW(ref_model.parameters(), lr=1e-2, weight_decay=0.01)
    opt = mlsl.DistributedOptimizer(model_h.parameters(), lr=1e-2, weight_decay=0.01, optimizer="adamw", mode="fused", bucket_mb=0.002)
    gens = [torch.Generator().manual_seed(300 + q) for q in range(world)]
    for _ in range(3):
        batches = [(torch.randn(8, 16, generator=gq), torch.randn(8, 4, generator=gq)) for gq in gens]
        ref_opt.zero_grad()
        for x, y in batches:
            (torch.nn.functional.mse_loss(ref_model(x), y) / world).backward()
        ref_opt.step()
        opt.zero_grad()
        x, y = batches[r]
        torch.nn.functional.mse_loss(model_h(x.to(torch.bfloat16)).float(), y).backward()
        opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in model_h.parameters()]).contiguous()
    same = flat.clone()
    mlsl.bcast(same, root=0)
    assert torch.equal(flat, same), "bf16 replicas differ"
    want = torch.cat([p.detach().reshape(-1) for p in ref_model.parameters()])
    assert torch.allclose(flat.float(), want, rtol=0.05, atol=0.02), (flat.float() - want).abs().max()
    opt.close()
    # quantised all-reduce over the wire (CT_QUANTIZATION: block-scaled FP8 + error feedback): close to the exact mean,
    # bitwise identical on every rank, and the error feedback keeps the running mean of repeated reductions unbiased
    n = 5000 + 37 * seed
    gq = torch.Generator().manual_seed(900 + r)
    exact_sum, quant_sum = torch.zeros(n), torch.zeros(n)
    for it in range(6):
        x = torch.randn(n, generator=gq)
        exact = x.clone()
        mlsl.allreduce(exact, scale=1.0 / world)
        q = x.clone()
        mlsl.allreduce(q, scale=1.0 / world, compress=True)
        same = q.clone()
        mlsl.bcast(same, root=0)
        assert torch.equal(q, same), "quantised all-reduce differs between ranks"
        rel = ((q - exact).norm() / exact.norm()).item()
        assert 0 < rel < 0.08, rel
        exact_sum += exact
        quant_sum += q
    drift = ((quant_sum - exact_sum).norm() / exact_sum.norm()).item()
    assert drift < 0.08, drift
    # the persistent (ParameterSet) flavour keeps one residual per request: error feedback across iterations
    torch.manual_seed(6)
    model2 = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
    opt = mlsl.DistributedOptimizer(model2.parameters(), lr=1e-2, mode="allreduce", compress=True, bucket_mb=0.002)
    for _ in range(3):
        opt.zero_grad()
        torch.nn.functional.mse_loss(model2(torch.randn(8, 16, generator=g)), torch.randn(8, 4, generator=g)).backward()
        opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in model2.parameters()]).contiguous()
    ref = flat.clone()
    mlsl.bcast(ref, root=0)
    assert torch.equal(flat, ref)
    opt.close()
    mlsl.finalize()
    print("NET OK rank %d of %d (D=%d M=%d, %d ops)" % (r, world, D, M, len(program)), flush=True)


if __name__ == "__main__":
    main()
