"""Every Distribution collective on the host backend vs a plain PyTorch reference (in-process virtual ranks).

Covers what the reference's own tests never exercise (SURVEY section 4 "gaps"): Reduce/Gather/Scatter/AlltoAllv/
AllGatherv/Barrier/ReduceScatter through Distribution, MIN/MAX, DOUBLE/BYTE/BF16, non-divisible sizes,
non-power-of-two groups, user buffers vs Environment::Alloc buffers.
"""
import pytest
import torch

from conftest import run_ranks

DT = {torch.float32: 0, torch.float64: 1, torch.uint8: 2, torch.bfloat16: 3, torch.float16: 4, torch.int32: 5}


def _make(rank, n, dtype, seed=0):
    g = torch.Generator().manual_seed(1234 + 17 * rank + seed)
    if dtype in (torch.uint8, torch.int32):
        return torch.randint(0, 7, (n,), generator=g, dtype=torch.int32).to(dtype)
    return (torch.rand(n, generator=g, dtype=torch.float32) * 4 - 2).to(dtype)


def _ref_reduce(tensors, op):
    acc = tensors[0].to(torch.float64 if tensors[0].dtype.is_floating_point else torch.int64)
    for t in tensors[1:]:
        t = t.to(acc.dtype)
        acc = acc + t if op == "sum" else (torch.minimum(acc, t) if op == "min" else torch.maximum(acc, t))
    return acc


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.bfloat16, torch.int32, torch.uint8])
@pytest.mark.parametrize("op", ["sum", "min", "max"])
def test_allreduce(world, dtype, op):
    n = 1000 + world  # deliberately not divisible by the group size
    if dtype == torch.uint8 and op == "sum":
        pytest.skip("byte sums wrap around; covered by test_byte_sum_wraps")

    def body(r, mlsl):
        x = _make(r, n, dtype)
        heap = mlsl.alloc_tensor(n, dtype)
        heap.copy_(x)
        mlsl.allreduce(heap, op=op)          # zero-copy path (symmetric heap)
        user = x.clone()
        mlsl.allreduce(user, op=op)          # staged path (foreign buffer)
        return heap.clone(), user

    outs = run_ranks(world, body)
    ref = _ref_reduce([_make(r, n, dtype) for r in range(world)], op)
    tol = 0 if not dtype.is_floating_point else (1e-6 if dtype != torch.bfloat16 else 4e-2)
    for heap, user in outs:
        assert torch.allclose(heap.to(ref.dtype), ref, rtol=tol, atol=tol * 4)
        assert torch.equal(heap, user)
        assert torch.equal(heap, outs[0][0])  # bitwise identical on every rank


def test_byte_sum_wraps():
    def body(r, mlsl):
        t = torch.full((64,), 200, dtype=torch.uint8)
        mlsl.allreduce(t)
        return t

    for t in run_ranks(2, body):
        assert int(t[0]) == (400 % 256)


@pytest.mark.parametrize("world", [2, 4])
def test_allreduce_scale_and_out_of_place(world):
    n = 4096

    def body(r, mlsl):
        x = _make(r, n, torch.float32)
        out = torch.zeros(n)
        mlsl.allreduce(x, out=out, scale=1.0 / world)
        return x, out

    outs = run_ranks(world, body)
    ref = _ref_reduce([_make(r, n, torch.float32) for r in range(world)], "sum") / world
    for r, (x, out) in enumerate(outs):
        assert torch.equal(x, _make(r, n, torch.float32))      # send buffer untouched
        assert torch.allclose(out.double(), ref, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("world", [2, 3])
def test_reduce_scatter_allgather_roundtrip(world):
    n = 257

    def body(r, mlsl):
        x = _make(r, n * world, torch.float32)
        shard = mlsl.reduce_scatter(x)
        full = mlsl.allgather(shard)
        # in-place variants: result lands at offset 0 / contribution sits in the rank's slot
        y = x.clone()
        d = mlsl.world_distribution()
        e = mlsl.env()
        e.wait(d.reduce_scatter(y, y, n, 0, 0, 0))
        z = torch.zeros(n * world)
        z[r * n:(r + 1) * n] = shard
        e.wait(d.all_gather(z, n, z, 0, 0))
        return shard, full, y[:n].clone(), z

    outs = run_ranks(world, body)
    ref = _ref_reduce([_make(r, n * world, torch.float32) for r in range(world)], "sum").float()
    for r, (shard, full, inplace, z) in enumerate(outs):
        assert torch.allclose(shard, ref[r * n:(r + 1) * n], rtol=1e-6, atol=1e-6)
        assert torch.allclose(full, ref, rtol=1e-6, atol=1e-6)
        assert torch.equal(inplace, shard)
        assert torch.equal(z, full)


def test_bcast_reduce_gather_scatter():
    world, n = 4, 333

    def body(r, mlsl):
        d, e = mlsl.world_distribution(), mlsl.env()
        b = _make(1, n, torch.float32) if r == 1 else torch.zeros(n)
        mlsl.bcast(b, root=1)
        x = _make(r, n, torch.float64)
        red = torch.zeros(n, dtype=torch.float64)
        mlsl.reduce(x, out=red, root=2, op="max")
        g = torch.zeros(n * world) if r == 3 else torch.zeros(1)
        mine = _make(r, n, torch.float32)
        e.wait(d.gather(mine, n, g, 0, 3, 0))
        src = torch.arange(n * world, dtype=torch.float32) if r == 0 else torch.zeros(1)
        sc = torch.zeros(n)
        e.wait(d.scatter(src, sc, n, 0, 0, 0))
        return b, red, g, sc

    outs = run_ranks(world, body)
    for r, (b, red, g, sc) in enumerate(outs):
        assert torch.equal(b, _make(1, n, torch.float32))
        assert torch.equal(sc, torch.arange(n * world, dtype=torch.float32)[r * n:(r + 1) * n])
    assert torch.equal(outs[2][1], _ref_reduce([_make(r, n, torch.float64) for r in range(world)], "max"))
    assert torch.equal(outs[3][2], torch.cat([_make(r, n, torch.float32) for r in range(world)]))


@pytest.mark.parametrize("world", [2, 3, 4])
def test_alltoall_and_v(world):
    n = 50

    def body(r, mlsl):
        d, e = mlsl.world_distribution(), mlsl.env()
        x = torch.arange(world * n, dtype=torch.float32) + 1000 * r
        y = mlsl.alltoall(x)
        # v-variant: rank r sends (p + 1) elements to peer p, taken from offset 10 * p
        sc = [p + 1 for p in range(world)]
        so = [10 * p for p in range(world)]
        rc = [r + 1] * world
        ro = [p * (r + 1) for p in range(world)]
        out = torch.zeros(world * (r + 1))
        e.wait(d.all_to_allv(x, sc, so, out, rc, ro, 0, 0))
        # all_gatherv: rank p contributes p + 2 elements
        cnts = [p + 2 for p in range(world)]
        gv = torch.zeros(sum(cnts))
        part = x[:r + 2].clone()   # the raw API does not keep buffers alive: hold the reference until wait()
        e.wait(d.all_gatherv(part, r + 2, gv, cnts, 0, 0))
        return y, out, gv

    outs = run_ranks(world, body)
    for r, (y, out, gv) in enumerate(outs):
        for p in range(world):
            src = torch.arange(world * n, dtype=torch.float32) + 1000 * p
            assert torch.equal(y[p * n:(p + 1) * n], src[r * n:(r + 1) * n])
            assert torch.equal(out[p * (r + 1):(p + 1) * (r + 1)], src[10 * r:10 * r + r + 1])
        exp = torch.cat([(torch.arange(world * n, dtype=torch.float32) + 1000 * p)[:p + 2] for p in range(world)])
        assert torch.equal(gv, exp)


def test_send_recv_list_ring_shift():
    world, n = 4, 16

    def body(r, mlsl):
        d, e = mlsl.world_distribution(), mlsl.env()
        x = torch.full((n,), float(r))
        out = torch.zeros(n)
        nxt, prv = (r + 1) % world, (r - 1) % world
        sc = [n if p == nxt else 0 for p in range(world)]
        rc = [n if p == prv else 0 for p in range(world)]
        e.wait(d.send_recv_list(x, sc, [0] * world, out, rc, [0] * world, 0, 0))
        return out

    for r, out in enumerate(run_ranks(world, body)):
        assert torch.equal(out, torch.full((n,), float((r - 1) % world)))


def test_test_polling_and_barrier():
    def body(r, mlsl):
        w = mlsl.allreduce(torch.ones(100000), async_op=True)
        spins = 0
        while not w.is_completed():
            spins += 1
        assert w.is_completed()
        mlsl.barrier()
        return float(w.result[0])

    assert run_ranks(3, body) == [3.0, 3.0, 3.0]


def test_quantized_allreduce_matches_definition():
    """fp8 block-scaled transport with error feedback: bounded error, identical on all ranks, residual converges."""
    world, n = 4, 5000

    def body(r, mlsl):
        x = _make(r, n, torch.float32) * 3
        outs = []
        for _ in range(3):
            y = torch.zeros(n)
            mlsl.allreduce(x, out=y, compress=True)
            outs.append(y)
        return outs

    outs = run_ranks(world, body)
    ref = _ref_reduce([_make(r, n, torch.float32) * 3 for r in range(world)], "sum").float()
    for it in range(3):
        for r in range(1, world):
            assert torch.equal(outs[r][it], outs[0][it])
        err = (outs[0][it] - ref).abs().max() / ref.abs().max()
        assert err < 0.08, err
    # a request created per call has a fresh residual, so every iteration gives the same answer
    assert torch.equal(outs[0][0], outs[0][1])


def test_quantized_allreduce_through_user_plugin():
    """Environment::SetQuantizationParams with a lib_path: the host backend dlopen()s the user's library and runs the
    all-reduce on its opaque blocks (268-byte blocks of 256 int8 + header, the geometry of the reference's test)."""
    import os
    from conftest import ROOT
    plugin = os.path.join(ROOT, "bin", "libmlsl_quant_sample.so")
    assert os.path.exists(plugin), "make builds bin/libmlsl_quant_sample.so"
    world, n = 4, 3000      # 11.7 blocks: a partial tail block, more ranks than evenly divides

    def body(r, mlsl):
        env = mlsl.env()
        env.set_quantization_params(plugin, "sample_compress", "sample_decompress", "sample_reduce_sum", 268, 256)
        assert env.get_quantization_params()["block_size"] == 268
        x = _make(r, n, torch.float32) * 3
        y = torch.zeros(n)
        mlsl.allreduce(x, out=y, compress=True, scale=0.5)
        # persistent request (ParameterSet with compression): the error-feedback residual carries over
        from mlsl_b200.api import CompressionType, DataType, OperationType
        sess = env.create_session()
        sess.set_global_minibatch_size(world)
        dist = env.create_distribution(world, 1)
        ri = sess.create_operation_reg_info(OperationType.CC)
        ri.add_input(1, 1, DataType.FLOAT)
        ri.add_output(1, 1, DataType.FLOAT)
        ri.add_parameter_set(n, 1, DataType.FLOAT, False, CompressionType.QUANTIZATION)
        op = sess.get_operation(sess.add_operation(ri, dist))
        sess.commit()
        ps = op.get_parameter_set(0)
        g = mlsl.alloc_tensor(n, torch.float32)
        acc = torch.zeros(n)
        for _ in range(8):
            g.copy_(x)
            ps.start_gradient_comm(g)
            ps.wait_gradient_comm()
            acc += g
        return y, acc / 8

    outs = run_ranks(world, body)
    ref = _ref_reduce([_make(r, n, torch.float32) * 3 for r in range(world)], "sum").float()
    for r in range(world):
        assert torch.equal(outs[r][0], outs[0][0]) and torch.equal(outs[r][1], outs[0][1])
    err1 = (outs[0][0] - 0.5 * ref).abs().max() / ref.abs().max()
    assert err1 < 0.03, err1
    # with error feedback the time average converges towards the exact sum
    err8 = (outs[0][1] - ref).abs().max() / ref.abs().max()
    assert err8 < err1 * 2 and err8 < 0.02, (err1, err8)


@pytest.mark.parametrize("world", [1, 4])
def test_rma_window_put_get_fence(world):
    """[ext] one-sided windows: every rank puts a tagged block into its right neighbour, fences, then gets a block back
    from its left neighbour's window; range errors are reported."""
    n = 64

    def body(r, mlsl):
        from mlsl_b200.api import GroupType
        from mlsl_b200._lib import MLSLError
        d = mlsl.world_distribution()
        mem = mlsl.alloc_tensor(2 * n, torch.float32)
        mem[:n] = float(r)                      # first half: my own data; second half: filled by my left neighbour
        win = d.create_window(mem, GroupType.GLOBAL)
        assert win.get_size((r + 1) % world) == 2 * n * 4
        win.fence()
        src = torch.full((n,), 100.0 + r)
        win.put(src, (r + 1) % world, target_disp=n)
        win.fence()
        got_put = mem[n:].clone()               # what my left neighbour put here
        fetched = torch.zeros(n)
        win.get(fetched, (r - 1) % world, target_disp=0)
        win.fence()
        with pytest.raises(MLSLError, match="exceeds"):
            win.put(src, (r + 1) % world, target_disp=2 * n - 1)
        win.free()
        return got_put, fetched

    for r, (got_put, fetched) in enumerate(run_ranks(world, body)):
        left = (r - 1) % world
        assert torch.equal(got_put, torch.full((n,), 100.0 + left))
        assert torch.equal(fetched, torch.full((n,), float(left)))


def test_tensor_api_rejects_inconsistent_arguments():
    """Shapes / dtypes / names that would make the native call read or write out of bounds are refused up front."""
    def body(r, mlsl):
        bad = []

        def refused(fn, exc=ValueError):
            try:
                fn()
                bad.append("accepted")
            except exc:
                pass

        refused(lambda: mlsl.allreduce(torch.arange(12.).view(3, 4).t()))                       # not contiguous
        refused(lambda: mlsl.allreduce(torch.ones(4), out=torch.ones(3)))                       # too small
        refused(lambda: mlsl.allreduce(torch.ones(4), out=torch.ones(4, dtype=torch.float64)))  # other dtype
        refused(lambda: mlsl.reduce_scatter(torch.ones(5)))
        refused(lambda: mlsl.alltoall(torch.ones(5)))
        refused(lambda: mlsl.allgather(torch.ones(4), out=torch.ones(4)))
        refused(lambda: mlsl.reduce(torch.ones(4), out=torch.ones(2)))
        refused(lambda: mlsl.allreduce(torch.ones(4), op="prod"))
        refused(lambda: mlsl.allreduce(torch.ones(4), group="nobody"))
        refused(lambda: mlsl.allreduce(torch.ones(4, dtype=torch.int64)), TypeError)
        x = torch.ones(4)
        mlsl.allreduce(x)            # still consistent on every rank after the refused calls
        return bad, float(x[0])

    for bad, v in run_ranks(2, body):
        assert bad == [] and v == 2.0


def test_gather_and_scatter_tensor_api():
    """mlsl.gather / mlsl.scatter (Distribution::Gather / Scatter) for every root, a dtype the library reduces and one that
    only travels as bytes."""
    def body(r, mlsl):
        P = mlsl.world_size()
        for dtype in (torch.float32, torch.int64):
            for root in range(P):
                mine = (torch.arange(5) + 10 * r).to(dtype)
                got = mlsl.gather(mine, root=root, group="global")
                if r == root:
                    assert torch.equal(got, torch.cat([(torch.arange(5) + 10 * p).to(dtype) for p in range(P)]))
                else:
                    assert got is None
                whole = (torch.arange(3 * P) + 100 * root).to(dtype)
                part = mlsl.scatter(whole if r == root else torch.empty(3, dtype=dtype), root=root, group="global")
                assert torch.equal(part, whole[3 * r:3 * r + 3]), (r, root, part)
        with pytest.raises(ValueError):
            mlsl.scatter(torch.zeros(3 * P + 1), root=r, group="global") if P > 1 else (_ for _ in ()).throw(ValueError())
        return True

    assert run_ranks(3, body) == [True] * 3


def test_fp8_codec_matches_its_definition():
    """The table / bit-pattern FP8 (E4M3) codec of the quantised all-reduce against the libm definition on every 1021st
    float bit pattern (the whole 2^32 range takes 15 s: bin/quant_codec_check 1)."""
    import os
    import subprocess
    from conftest import ROOT
    p = subprocess.run([os.path.join(ROOT, "bin", "quant_codec_check"), "1021"], stdout=subprocess.PIPE, text=True, timeout=120)
    assert p.returncode == 0 and " 0 mismatches" in p.stdout, p.stdout


def test_allgatherv_tensor_api():
    def body(r, mlsl):
        P = mlsl.world_size()
        counts = [(p * 3) % 5 + 1 for p in range(P)]
        mine = torch.full((counts[r],), float(r + 1))
        got = mlsl.allgatherv(mine, counts, group="global")
        assert torch.equal(got, torch.cat([torch.full((counts[p],), float(p + 1)) for p in range(P)]))
        with pytest.raises(ValueError):
            mlsl.allgatherv(mine, counts[:-1], group="global")
        return True

    assert run_ranks(4, body) == [True] * 4


def test_compressed_allreduce_keeps_error_feedback_across_calls():
    """mlsl.allreduce(compress=True) on the same buffers: the residual of call k is added to the input of call k+1 (the
    reference keys its residual by the buffer address, quant/quant.c:153-167), so the quantisation error of the running
    MEAN of the results shrinks with the number of calls instead of staying at the single-call level."""
    import torch
    world, n, calls = 2, 4096, 24

    def body(r, mlsl):
        g = torch.Generator().manual_seed(7 + r)
        x = mlsl.alloc_tensor(n, torch.float32)
        x.copy_(torch.randn(n, generator=g))
        y = mlsl.alloc_tensor(n, torch.float32)
        outs = []
        for _ in range(calls):
            mlsl.allreduce(x, out=y, compress=True)
            outs.append(y.clone())
        return x.clone(), torch.stack(outs)

    res = run_ranks(world, body)
    exact = sum(r[0] for r in res)
    outs = res[0][1]
    err_first = (outs[0] - exact).abs().mean().item()
    err_mean = (outs.mean(0) - exact).abs().mean().item()
    assert err_first > 0
    assert err_mean < 0.35 * err_first, (err_first, err_mean)


@pytest.mark.parametrize("mx", ["0", "1"])
def test_compressed_allreduce_formats(mx):
    """Both fp8 wire formats of the compressed all-reduce - one fp32 scale per 128 elements, or the MX layout (one ue8m0
    power-of-two scale per 32 elements, MLSL_QUANT_MX=1) - stay within e4m3 precision of the exact sum, identically on every rank,
    also when magnitudes differ by orders between neighbouring 32-element groups (where the finer MX scales help)."""
    import torch
    world, n = 4, 5000

    def body(r, mlsl):
        g = torch.Generator().manual_seed(11 + r)
        x = torch.randn(n, generator=g)
        x[::64] *= 1000.0                      # a few large entries per 128-block
        y = torch.zeros(n)
        mlsl.allreduce(x, out=y, compress=True)
        return x, y

    res = run_ranks(world, body, env={"MLSL_QUANT_MX": mx})
    exact = sum(r[0] for r in res)
    for _, y in res:
        assert torch.equal(y, res[0][1])
    err = (res[0][1] - exact).abs()
    assert err.max() <= 0.07 * exact.abs().max()               # e4m3: 3 mantissa bits, two quantisation steps
    small = torch.ones(n, dtype=torch.bool)
    small[::64] = False                                          # the entries that share a scale with a 1000x larger neighbour
    if mx == "1":                                                # MX: only the 32-group of the outlier is coarse
        far = small.clone()
        for k in range(0, n, 64):
            far[k:k + 32] = False
        assert err[far].max() < 0.5, err[far].max()              # ~N(0, 2) sums quantised at their own scale



@pytest.mark.parametrize("world", [16, 48])
def test_many_ranks_world_and_subgroups(world):
    """Well past the eight ranks of one GPU node: the host backend serves up to 64 ranks (kMaxHostRanks) - world-wide collectives
    and the data / model groups of a (world / 4) x 4 distribution."""
    def body(r, mlsl):
        W = mlsl.world_size()
        x = torch.full((1000,), float(r + 1))
        mlsl.allreduce(x)
        a = mlsl.alltoall(torch.arange(W * 3, dtype=torch.float32) + 1000 * r)
        g = mlsl.allgather(torch.tensor([float(r)]))
        e = mlsl.env()
        d = e.create_distribution(W // 4, 4)
        y = torch.ones(64)
        mlsl.allreduce(y, group="model", distribution=d)
        z = torch.ones(64)
        mlsl.allreduce(z, group="data", distribution=d)
        e.delete_distribution(d)
        return float(x[0]), a.tolist(), float(g.sum()), float(y[0]), float(z[0])

    for r, (x0, a, gsum, y0, z0) in enumerate(run_ranks(world, body)):
        assert x0 == world * (world + 1) / 2 and gsum == world * (world - 1) / 2
        assert a == [1000.0 * q + 3 * r + k for q in range(world) for k in range(3)]
        assert (y0, z0) == (4.0, world / 4)
