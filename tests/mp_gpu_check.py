"""Multi-process, multi-GPU correctness check (one rank per GPU, launched by torchrun):
    python -m torch.distributed.run --nproc-per-node N tests/mp_gpu_check.py
Every collective over real NVLink peers (CUDA IPC / VMM mappings, NVLS multicast when N >= 4) against a reference
computed from the deterministic per-rank inputs.  Prints one PASSED/FAILED line per check on rank 0; exit code 1 on failure."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsl_b200 as mlsl  # noqa: E402


def make(rank, n, dtype, seed=0):
    g = torch.Generator().manual_seed(999 + 13 * rank + seed)
    if dtype == torch.int32:
        return torch.randint(0, 9, (n,), generator=g, dtype=torch.int32)
    return (torch.rand(n, generator=g) * 4 - 2).to(dtype)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MLSL_HEAP_SIZE_GB", "2")
    os.environ.setdefault("MLSL_WATCHDOG_SEC", "30")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    env = mlsl.init()
    r, W = mlsl.rank(), mlsl.world_size()
    fails = []

    def check(name, ok):
        t = torch.tensor([1.0 if ok else 0.0], device="cuda")
        mlsl.allreduce(t, op="min")
        torch.cuda.synchronize()
        good = bool(t.item() > 0.5)
        if r == 0:
            print("%s: %s" % (name, "PASSED" if good else "FAILED"), flush=True)
        if not good:
            fails.append(name)

    if r == 0:
        print("backend:", env.describe_backend(), flush=True)
    for dtype, tol in ((torch.float32, 1e-5), (torch.bfloat16, 5e-2), (torch.float16, 2e-2), (torch.float64, 1e-12), (torch.int32, 0)):
        for n in (1, 1000, 262147, 8 * 1024 * 1024 + 5):
            x = mlsl.alloc_tensor(n, dtype)
            x.copy_(make(r, n, dtype))
            y = mlsl.alloc_tensor(n, dtype)
            mlsl.allreduce(x, out=y, scale=0.5 if dtype.is_floating_point else 1.0)
            ref = sum(make(p, n, dtype).double() for p in range(W)) * (0.5 if dtype.is_floating_point else 1.0)
            torch.cuda.synchronize()
            err = (y.double().cpu() - ref).abs().max().item() / max(1.0, ref.abs().max().item())
            check("allreduce %s n=%d (err %.2e)" % (str(dtype).split(".")[-1], n, err), err <= tol)
            mlsl.allreduce(x)   # in place
            torch.cuda.synchronize()
            err = (x.double().cpu() - ref / (0.5 if dtype.is_floating_point else 1.0)).abs().max().item() / max(1.0, ref.abs().max().item())
            check("allreduce in-place %s n=%d" % (str(dtype).split(".")[-1], n), err <= 2 * tol)
            mlsl.free_tensor(x)
            mlsl.free_tensor(y)
    n = 100003
    x = mlsl.alloc_tensor(n * W, torch.float32)
    x.copy_(make(r, n * W, torch.float32))
    shard = mlsl.reduce_scatter(x)
    full = mlsl.allgather(shard)
    ref = sum(make(p, n * W, torch.float32).double() for p in range(W)).float()
    torch.cuda.synchronize()
    check("reduce_scatter", torch.allclose(shard.cpu(), ref[r * n:(r + 1) * n], rtol=1e-5, atol=1e-5))
    check("allgather", torch.allclose(full.cpu(), ref, rtol=1e-5, atol=1e-5))
    a = (torch.arange(W * n, dtype=torch.float32) + 1000 * r).cuda()
    a2a = mlsl.alltoall(a)
    torch.cuda.synchronize()
    ok = all(torch.equal(a2a[p * n:(p + 1) * n].cpu(), (torch.arange(W * n, dtype=torch.float32) + 1000 * p)[r * n:(r + 1) * n]) for p in range(W))
    check("alltoall", ok)
    b = (make(W - 1, n, torch.float32) if r == W - 1 else torch.zeros(n)).cuda()
    mlsl.bcast(b, root=W - 1)
    torch.cuda.synchronize()
    check("bcast", torch.equal(b.cpu(), make(W - 1, n, torch.float32)))
    q = mlsl.alloc_tensor(n, torch.float32)
    src = mlsl.alloc_tensor(n, torch.float32)
    src.copy_(make(r, n, torch.float32))
    mlsl.allreduce(src, out=q, compress=True)
    refq = sum(make(p, n, torch.float32).double() for p in range(W)).float()
    torch.cuda.synchronize()
    check("allreduce fp8-compressed", ((q.cpu() - refq).abs().max() / refq.abs().max()).item() < 0.1)
    # ---- mid sizes: the multi-CTA flag-in-data kernel (one-shot / two-shot), changing sizes back to back -------------
    for k, nb in enumerate((8200, 65536, 300004, 1 << 20, 70004, 1 << 19)):
        n = nb // 4
        x = make(r, n, torch.float32, seed=k).cuda()                     # foreign buffer: read in place by the kernel
        y = mlsl.alloc_tensor(n, torch.float32)
        mlsl.allreduce(x, out=y, scale=1.0 / W)
        ref = sum(make(p, n, torch.float32, seed=k).double() for p in range(W)) / W
        torch.cuda.synchronize()
        err = (y.double().cpu() - ref).abs().max().item()
        check("allreduce mid %d B (err %.1e)" % (nb, err), err <= 1e-5)
        mlsl.free_tensor(y)
    # ---- large symmetric buffers: NVLS flavours (multimem ld_reduce reduce-scatter, multimem.st bcast) and the bulk-copy
    #      (cp.async.bulk) gather-like collectives, random data, in place and out of place -------------------------------
    n = (4 << 20) // 4 + 12
    x = mlsl.alloc_tensor(n * W, torch.float32)
    x.copy_(make(r, n * W, torch.float32, seed=7))
    shard = mlsl.alloc_tensor(n, torch.float32)
    mlsl.reduce_scatter(x, out=shard, scale=0.25)
    ref = (sum(make(p, n * W, torch.float32, seed=7).double() for p in range(W)) * 0.25).float()
    torch.cuda.synchronize()
    check("reduce_scatter 4 MiB shards (symmetric heap)", torch.allclose(shard.cpu(), ref[r * n:(r + 1) * n], rtol=1e-5, atol=1e-5))
    full = mlsl.alloc_tensor(n * W, torch.float32)
    mlsl.allgather(shard, out=full)
    torch.cuda.synchronize()
    check("allgather 4 MiB shards (bulk copy)", torch.allclose(full.cpu(), ref, rtol=1e-5, atol=1e-5))
    full[r * n:(r + 1) * n].mul_(2.0)
    mlsl.allgather(full[r * n:(r + 1) * n], out=full)                    # in place: my shard already sits at r * n
    torch.cuda.synchronize()
    check("allgather in place", torch.allclose(full.cpu(), ref * 2, rtol=1e-5, atol=1e-5))
    a2a = mlsl.alloc_tensor(n * W, torch.float32)
    mlsl.alltoall(x, out=a2a)
    torch.cuda.synchronize()
    ok = all(torch.equal(a2a[p * n:(p + 1) * n].cpu(), make(p, n * W, torch.float32, seed=7)[r * n:(r + 1) * n]) for p in range(W))
    check("alltoall 4 MiB blocks (bulk copy)", ok)
    for root in (0, W - 1):
        b = mlsl.alloc_tensor(n * 4 + 3, torch.float32)
        b.copy_(make(r, n * 4 + 3, torch.float32, seed=11))
        mlsl.bcast(b, root=root)
        torch.cuda.synchronize()
        check("bcast 16 MiB from %d (symmetric heap)" % root, torch.equal(b.cpu(), make(root, n * 4 + 3, torch.float32, seed=11)))
        mlsl.free_tensor(b)
    g = mlsl.gather(shard, root=W - 1)
    torch.cuda.synchronize()
    check("gather", r != W - 1 or torch.allclose(g.cpu(), ref, rtol=1e-5, atol=1e-5))
    # ---- sub-groups (data x model), the randomised collective programs and the Session graph twin ----------------------
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_fuzz_cpu import check_rank, make_program, run_program
    for seed in (2, 4, 5):
        D, M, program = make_program(W, seed)
        res = run_program(r, mlsl, W, D, M, program, "cuda")
        ok = True
        try:
            check_rank(r, W, D, M, program, res)
        except AssertionError as e:
            ok = False
            print("rank %d fuzz seed %d: %r" % (r, seed, e), flush=True)
        check("random collective program seed %d (%d x %d)" % (seed, D, M), ok)
    for (D, M) in ((2, W // 2), (W // 2, 2)) if W >= 4 else ():
        dist = env.create_distribution(D, M)
        for grp, P in (("data", D), ("model", M)):
            t = mlsl.alloc_tensor(200003, torch.float32)
            t.fill_(float(r + 1))
            mlsl.allreduce(t, group=grp, distribution=dist)
            mem = [q for q in range(W) if (q % M == r % M if grp == "data" else q // M == r // M)]
            torch.cuda.synchronize()
            check("sub-group allreduce %dx%d %s" % (D, M, grp), bool((t == float(sum(q + 1 for q in mem))).all().item()) and len(mem) == P)
            mlsl.free_tensor(t)
        env.delete_distribution(dist)
    # ---- tcgen05 GEMMs fused with their collectives, one rank per GPU (k_gemm_rs2 on CTA pairs by default, k_ag_gemm) ----------
    from mlsl_b200.ops.ag_gemm import allgather_gemm
    from mlsl_b200.ops.gemm_rs import gemm_reduce_scatter
    M, N, K = 256 * W, 512, 256
    gen = torch.Generator().manual_seed(4242)
    a_full = (torch.randn(W, M, K, generator=gen) * 0.25).to(torch.bfloat16)          # rank p multiplies a_full[p] @ w_full[p].T
    w_full = (torch.randn(W, N, K, generator=gen) * 0.25).to(torch.bfloat16)
    out = gemm_reduce_scatter(a_full[r].cuda(), w_full[r].cuda(), out_dtype=torch.float32, group="data")
    ref = sum(a_full[p].float() @ w_full[p].float().t() for p in range(W))
    torch.cuda.synchronize()
    rows = M // W
    err = (out.cpu() - ref[r * rows:(r + 1) * rows]).abs().max().item() / ref.abs().max().item()
    check("gemm + reduce-scatter %dx%dx%d (rel err %.1e)" % (M, N, K, err), err < 2e-2)
    xs = a_full[0][r * rows:(r + 1) * rows].contiguous().cuda()                        # row shard of ONE matrix
    y, xg = allgather_gemm(xs, w_full[0].cuda(), out_dtype=torch.float32, group="data")
    torch.cuda.synchronize()
    ref2 = a_full[0].float() @ w_full[0].float().t()
    err2 = (y.cpu() - ref2).abs().max().item() / ref2.abs().max().item()
    check("all-gather + gemm %dx%dx%d (rel err %.1e)" % (M, N, K, err2), err2 < 2e-2 and torch.equal(xg.cpu(), a_full[0]))
    # ---- device-heap expansion: one allocation bigger than the whole initial heap (MLSL_HEAP_SIZE_GB=2); the new chunk is
    #      mapped by every peer's watcher thread, then peers read / write it like any other heap memory ----------------------
    if "VMM" in env.describe_backend():
        nbig = int(2.25 * (1 << 30)) // 4
        big = mlsl.alloc_tensor(nbig, torch.float32, zero=False)
        big.fill_(float(r + 1))
        mlsl.allreduce(big)
        torch.cuda.synchronize()
        want = float(W * (W + 1) // 2)
        check("heap growth: all-reduce inside a 2.25 GiB allocation on a 2 GiB heap (%s)" % env.describe_backend().split(",")[2].strip(),
              bool((big[:1 << 20] == want).all().item()) and bool((big[-(1 << 20):] == want).all().item()))
        mlsl.free_tensor(big)
    mlsl.finalize()
    if r == 0:
        print("mp_gpu_check: %s" % ("ALL PASSED" if not fails else "FAILED: %s" % fails), flush=True)
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
