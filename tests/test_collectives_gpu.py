"""CUDA peer-memory kernels vs plain PyTorch fp32/fp64 references.

Runs on ONE GPU: the ranks are in-process virtual ranks (threads) that all use cuda:0, every rank with its own
slab, stream and kernels - the peer pointers the kernels dereference are simply other allocations on the same
device, so the full handshake / pull / push protocol of the NVLink kernels is exercised.
"""
import pytest
import torch

from conftest import run_ranks

pytestmark = pytest.mark.gpu


def _make(rank, n, dtype, seed=0):
    g = torch.Generator().manual_seed(4321 + 31 * rank + seed)
    if dtype in (torch.uint8, torch.int32):
        return torch.randint(0, 7, (n,), generator=g, dtype=torch.int32).to(dtype)
    return (torch.rand(n, generator=g, dtype=torch.float32) * 4 - 2).to(dtype)


def _ref_reduce(tensors, op):
    acc = tensors[0].to(torch.float64 if tensors[0].dtype.is_floating_point else torch.int64)
    for t in tensors[1:]:
        t = t.to(acc.dtype)
        acc = acc + t if op == "sum" else (torch.minimum(acc, t) if op == "min" else torch.maximum(acc, t))
    return acc


def _gpu(fn, world, **kw):
    kw.setdefault("wait_mode", "host")     # no autograd hooks here: blocking waits keep `.cpu()` off spinning streams
    return run_ranks(world, fn, backend="cuda", env={"MLSL_HEAP_SIZE_GB": "0.5", "MLSL_WATCHDOG_SEC": "20", **kw.pop("env", {})}, **kw)


def test_native_library_is_the_cuda_path():
    def body(r, mlsl):
        return mlsl.env().get_backend_name(), mlsl.is_device()

    assert _gpu(body, 2) == [("cuda", True)] * 2


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16, torch.float64, torch.int32])
@pytest.mark.parametrize("op", ["sum", "max"])
def test_allreduce(world, dtype, op):
    n = 100003  # odd: exercises the vector body and the scalar tail

    def body(r, mlsl):
        x = _make(r, n, dtype).cuda()
        heap = mlsl.alloc_tensor(n, dtype)
        heap.copy_(x)
        mlsl.allreduce(heap, op=op)                # zero-copy: buffer lives in the symmetric heap
        user = x.clone()
        mlsl.allreduce(user, op=op)                # foreign cudaMalloc buffer: staged through the heap
        torch.cuda.current_stream().synchronize()
        return heap.cpu(), user.cpu()

    outs = _gpu(body, world)
    ref = _ref_reduce([_make(r, n, dtype) for r in range(world)], op)
    tol = 0 if not dtype.is_floating_point else {torch.float32: 1e-6, torch.float64: 1e-12}.get(dtype, 4e-2)
    for heap, user in outs:
        assert torch.allclose(heap.to(ref.dtype), ref, rtol=tol, atol=tol * 4)
        assert torch.equal(heap, user)
        assert torch.equal(heap, outs[0][0])        # bitwise identical on every rank


@pytest.mark.parametrize("n", [1, 7, 1024, 262144 + 3, 4 * 1024 * 1024])
def test_allreduce_sizes_scale_out_of_place(n):
    world = 4

    def body(r, mlsl):
        x = mlsl.alloc_tensor(n, torch.float32)
        x.copy_(_make(r, n, torch.float32))
        y = mlsl.alloc_tensor(n, torch.float32)
        mlsl.allreduce(x, out=y, scale=0.25)
        torch.cuda.current_stream().synchronize()
        return x.cpu(), y.cpu()

    outs = _gpu(body, world)
    ref = (_ref_reduce([_make(r, n, torch.float32) for r in range(world)], "sum") * 0.25)
    for r, (x, y) in enumerate(outs):
        assert torch.equal(x, _make(r, n, torch.float32))
        assert torch.allclose(y.double(), ref, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("dtype,op", [(torch.float32, "sum"), (torch.bfloat16, "sum"), (torch.float64, "max"), (torch.uint8, "min")])
def test_allreduce_mid_flag_in_data_kernel(world, dtype, op):
    """8 KiB .. 1 MiB: the multi-CTA flag-in-data kernel, one-shot and two-shot, back to back with changing sizes (the
    arena is reused with alternating parity), in place and out of place, heap and foreign buffers, fused scale."""
    es = torch.empty((), dtype=dtype).element_size()
    sizes = [8200 // es + 1, 65536 // es, 300000 // es + 3, (1 << 20) // es, 70000 // es + 1]

    def body(r, mlsl):
        outs = []
        for k, n in enumerate(sizes):
            x = _make(r + k, n, dtype).cuda()
            if k % 2 == 0:
                heap = mlsl.alloc_tensor(n, dtype)
                heap.copy_(x)
                mlsl.allreduce(heap, op=op)
                outs.append(heap.clone())
            else:
                y = torch.empty_like(x)
                mlsl.allreduce(x, out=y, op=op, scale=0.5 if dtype.is_floating_point and op == "sum" else 1.0)
                outs.append(y)
        torch.cuda.current_stream().synchronize()
        return [o.cpu() for o in outs]

    res = _gpu(body, world)
    tol = 0 if not dtype.is_floating_point else {torch.float32: 1e-6, torch.float64: 1e-12}.get(dtype, 4e-2)
    for k, n in enumerate(sizes):
        ref = _ref_reduce([_make(r + k, n, dtype) for r in range(world)], op)
        if k % 2 == 1 and dtype.is_floating_point and op == "sum":
            ref = ref * 0.5
        for r in range(world):
            got = res[r][k]
            assert torch.allclose(got.to(ref.dtype), ref, rtol=tol, atol=tol * 4), (k, n, r)
            assert torch.equal(got, res[0][k])          # bit-identical on every rank


@pytest.mark.parametrize("n", [1001, 5001])
def test_unaligned_buffers(n):
    """4-byte aligned views: <= 8 KiB the LL kernel goes byte-wise, above it the handshake kernel takes its scalar path.
    Rank 1 passes a 16-byte aligned view instead - the choice of kernel may not depend on a rank's own alignment."""
    world = 2

    def body(r, mlsl):
        base = mlsl.alloc_tensor(n + 8, torch.float32)
        v = base[1:n + 1] if r == 0 else base[4:n + 4]
        v.copy_(_make(r, n, torch.float32))
        mlsl.allreduce(v)
        f = torch.zeros(n + 3, device="cuda")[3:]       # foreign (torch-owned) view, 4-byte aligned
        f.copy_(_make(r, n, torch.float32))
        mlsl.allreduce(f)
        torch.cuda.current_stream().synchronize()
        return v.cpu(), f.cpu()

    outs = _gpu(body, world)
    ref = _ref_reduce([_make(r, n, torch.float32) for r in range(world)], "sum").float()
    for v, f in outs:
        assert torch.allclose(v, ref, rtol=1e-6, atol=1e-6) and torch.equal(v, outs[0][0])
        assert torch.allclose(f, ref, rtol=1e-6, atol=1e-6) and torch.equal(f, outs[0][1])


@pytest.mark.parametrize("world", [2, 3, 8])
def test_reduce_scatter_allgather(world):
    n = 4099

    def body(r, mlsl):
        x = mlsl.alloc_tensor(n * world, torch.float32)
        x.copy_(_make(r, n * world, torch.float32))
        shard = mlsl.reduce_scatter(x, scale=0.5)
        full = mlsl.allgather(shard)
        # in-place forms
        d, e = mlsl.world_distribution(), mlsl.env()
        y = mlsl.alloc_tensor(n * world, torch.float32)
        y.copy_(x)
        e.wait(d.reduce_scatter(y, y, n, 0, 0, 0, 0.5))
        z = mlsl.alloc_tensor(n * world, torch.float32)
        z[r * n:(r + 1) * n] = shard
        e.wait(d.all_gather(z, n, z, 0, 0))
        torch.cuda.current_stream().synchronize()
        return shard.cpu(), full.cpu(), y[:n].cpu(), z.cpu()

    outs = _gpu(body, world)
    ref = (_ref_reduce([_make(r, n * world, torch.float32) for r in range(world)], "sum") * 0.5).float()
    for r, (shard, full, inplace, z) in enumerate(outs):
        assert torch.allclose(shard, ref[r * n:(r + 1) * n], rtol=1e-6, atol=1e-6)
        assert torch.allclose(full, ref, rtol=1e-6, atol=1e-6)
        assert torch.equal(inplace, shard)
        assert torch.equal(z, full)


@pytest.mark.parametrize("split", ["0", "1"])
@pytest.mark.parametrize("n", [3331, 300000])
def test_alltoall_split_knobs(split, n):
    """MLSL_ALLTOALL_SPLIT / MLSL_ALLTOALLV_SPLIT (reference src/comm_ep.cpp:1192,1270): 1 = every pair message spread over
    all channels, 0 = the channels are dealt to the pairs, all pairs moving at once.  Same result either way."""
    world = 4

    def body(r, mlsl):
        a = mlsl.alloc_tensor(world * n, torch.float32)
        a.copy_(torch.arange(world * n, dtype=torch.float32) + 1000000 * r)
        out = mlsl.alltoall(a)
        counts = [(p + r) % 3 * 1000 + 7 for p in range(world)]             # what I send to p
        rcounts = [(r + p) % 3 * 1000 + 7 for p in range(world)]            # what p sends to me (same formula, symmetric)
        v = mlsl.alltoallv(a[:sum(counts)], counts, rcounts)
        torch.cuda.current_stream().synchronize()
        return out.cpu(), v.cpu()

    outs = _gpu(body, world, env={"MLSL_ALLTOALL_SPLIT": split, "MLSL_ALLTOALLV_SPLIT": split})
    for r, (out, v) in enumerate(outs):
        for p in range(world):
            want = (torch.arange(world * n, dtype=torch.float32) + 1000000 * p)[r * n:(r + 1) * n]
            assert torch.equal(out[p * n:(p + 1) * n], want), (split, r, p)
        off = 0
        for p in range(world):
            pc = [(q + p) % 3 * 1000 + 7 for q in range(world)]             # p's send counts
            so = sum(pc[:r])
            cnt = pc[r]
            want = (torch.arange(world * n, dtype=torch.float32) + 1000000 * p)[so:so + cnt]
            assert torch.equal(v[off:off + cnt], want), (split, r, p)
            off += cnt


def test_heap_pool_makes_torch_allocations_zero_copy():
    """`with mlsl.heap_pool():` - torch allocations come from the symmetric heap (pluggable allocator), so a collective on
    them passes the pointer check that rejects foreign buffers (MLSL_POINTER_CHECK=1) and needs no staging."""
    world, n = 2, 100000

    def body(r, mlsl):
        with mlsl.heap_pool():
            x = torch.full((n,), float(r + 1), device="cuda")
            y = torch.empty(n, device="cuda")
        mlsl.allreduce(x, out=y)
        foreign_rejected = False
        try:
            mlsl.allreduce(torch.ones(n, device="cuda"))
        except mlsl.MLSLError:
            foreign_rejected = True
        torch.cuda.current_stream().synchronize()
        ok = bool((y == 3.0).all().item())
        del x, y
        return ok, foreign_rejected

    assert _gpu(body, world, env={"MLSL_POINTER_CHECK": "1"}) == [(True, True)] * world


def test_bcast_reduce_gather_scatter_alltoall():
    world, n = 4, 3331

    def body(r, mlsl):
        d, e = mlsl.world_distribution(), mlsl.env()
        b = (_make(1, n, torch.float32) if r == 1 else torch.zeros(n)).cuda()
        mlsl.bcast(b, root=1)
        x = _make(r, n, torch.float64).cuda()
        red = torch.zeros(n, dtype=torch.float64, device="cuda")
        mlsl.reduce(x, out=red, root=2, op="max")
        g = torch.zeros(n * world if r == 3 else 1, device="cuda")
        mine = _make(r, n, torch.float32).cuda()
        e.wait(d.gather(mine, n, g, 0, 3, 0))
        src = (torch.arange(n * world, dtype=torch.float32) if r == 0 else torch.zeros(1)).cuda()
        sc = torch.zeros(n, device="cuda")
        e.wait(d.scatter(src, sc, n, 0, 0, 0))
        a = (torch.arange(world * n, dtype=torch.float32) + 1000 * r).cuda()
        a2a = mlsl.alltoall(a)
        sc_counts = [p + 1 for p in range(world)]
        so = [10 * p for p in range(world)]
        rc = [r + 1] * world
        ro = [p * (r + 1) for p in range(world)]
        outv = torch.zeros(world * (r + 1), device="cuda")
        e.wait(d.all_to_allv(a, sc_counts, so, outv, rc, ro, 0, 0))
        torch.cuda.current_stream().synchronize()
        return b.cpu(), red.cpu(), g.cpu(), sc.cpu(), a2a.cpu(), outv.cpu()

    outs = _gpu(body, world)
    for r, (b, red, g, sc, a2a, outv) in enumerate(outs):
        assert torch.equal(b, _make(1, n, torch.float32))
        assert torch.equal(sc, torch.arange(n * world, dtype=torch.float32)[r * n:(r + 1) * n])
        for p in range(world):
            src = torch.arange(world * n, dtype=torch.float32) + 1000 * p
            assert torch.equal(a2a[p * n:(p + 1) * n], src[r * n:(r + 1) * n])
            assert torch.equal(outv[p * (r + 1):(p + 1) * (r + 1)], src[10 * r:10 * r + r + 1])
    assert torch.equal(outs[2][1], _ref_reduce([_make(r, n, torch.float64) for r in range(world)], "max"))
    assert torch.equal(outs[3][2], torch.cat([_make(r, n, torch.float32) for r in range(world)]))


@pytest.mark.parametrize("mx", ["0", "1"])
def test_quantized_allreduce_fp8_matches_host_definition(mx):
    """The fused fp8 kernel must agree with the CPU definition of the format (csrc/core/quant.hpp) to rounding - with one fp32
    scale per 128 elements and with the MX layout (MLSL_QUANT_MX=1: one ue8m0 power-of-two scale per 32 elements)."""
    world, n = 4, 70001

    def body(r, mlsl):
        x = mlsl.alloc_tensor(n, torch.float32)
        x.copy_(_make(r, n, torch.float32) * 3)
        y = mlsl.alloc_tensor(n, torch.float32)
        mlsl.allreduce(x, out=y, compress=True, scale=0.25)
        torch.cuda.current_stream().synchronize()
        return y.cpu()

    def host_body(r, mlsl):
        x = _make(r, n, torch.float32) * 3
        y = torch.zeros(n)
        mlsl.allreduce(x, out=y, compress=True, scale=0.25)
        return y

    dev = _gpu(body, world, env={"MLSL_QUANT_MX": mx})
    host = run_ranks(world, host_body, backend="host", env={"MLSL_QUANT_MX": mx})
    ref = (_ref_reduce([_make(r, n, torch.float32) * 3 for r in range(world)], "sum") * 0.25).float()
    for r in range(world):
        assert torch.equal(dev[r], dev[0])
    assert (dev[0] - ref).abs().max() / ref.abs().max() < 0.08
    mism = (dev[0] != host[0]).float().mean().item()
    assert mism < 1e-3, "device and host fp8 paths disagree on %.4f%% of the elements" % (100 * mism)


@pytest.mark.parametrize("opt", ["sgd", "adamw"])
@pytest.mark.parametrize("pdtype", [torch.float32, torch.bfloat16])
def test_fused_distributed_update(opt, pdtype):
    """reduce-scatter + optimizer + all-gather in one kernel vs torch.optim on the summed gradient."""
    world, owned = 4, 3000
    n = owned * world

    def body(r, mlsl):
        from mlsl_b200.api import DataType, OperationType, OptimizerType
        e = mlsl.env()
        sess = e.create_session()
        sess.set_global_minibatch_size(world)
        dist = mlsl.world_distribution()
        ri = sess.create_operation_reg_info(OperationType.CC)
        ri.add_input(1, 1, DataType.FLOAT)
        ri.add_output(1, 1, DataType.FLOAT)
        ri.add_parameter_set(n, 1, DataType.FLOAT, True)
        op = sess.get_operation(sess.add_operation(ri, dist))
        sess.commit()
        ps = op.get_parameter_set(0)
        assert ps.get_owned_kernel_count() == owned
        w0 = _make(99, n, torch.float32)
        grad = mlsl.alloc_tensor(n, torch.float32)
        param = mlsl.alloc_tensor(n, pdtype)
        param.copy_(w0)
        master = w0[r * owned:(r + 1) * owned].clone().cuda()
        s1 = torch.zeros(owned, device="cuda")
        s2 = torch.zeros(owned, device="cuda")
        for step in range(1, 4):
            grad.copy_(_make(r, n, torch.float32, seed=step))
            ps.start_fused_update(grad, param, mlsl.comm.mlsl_dtype(pdtype), master, s1, s2 if opt == "adamw" else None,
                                  OptimizerType.ADAMW if opt == "adamw" else OptimizerType.SGD, lr=0.05, momentum=0.9,
                                  weight_decay=0.01, step=step, grad_scale=1.0 / world)
            ps.wait_fused_update()
        torch.cuda.current_stream().synchronize()
        out = param.float().cpu()
        e.delete_session(sess)
        return out

    outs = _gpu(body, world)
    w = _make(99, n, torch.float32).clone().requires_grad_(True)
    o = (torch.optim.AdamW([w], lr=0.05, weight_decay=0.01) if opt == "adamw"
         else torch.optim.SGD([w], lr=0.05, momentum=0.9, weight_decay=0.01))
    for step in range(1, 4):
        w.grad = sum(_make(r, n, torch.float32, seed=step) for r in range(world)) / world
        o.step()
    ref = w.detach()
    tol = 2e-5 if pdtype == torch.float32 else 2e-2
    for out in outs:
        assert torch.allclose(out, ref, rtol=tol, atol=tol)
        assert torch.equal(out, outs[0])


def test_session_graph_on_device_hybrid():
    """2x2 hybrid: activation ReduceScatter/AllGather over the model group, gradient ReduceScatter + increment
    AllGather over the data group, with the device pack/unpack kernels; index-valued tensors as in the C++ test."""
    world, M = 4, 2

    def body(r, mlsl):
        from mlsl_b200.api import DataType, OperationType
        e = mlsl.env()
        sess = e.create_session()
        sess.set_global_minibatch_size(8)
        dist = e.create_distribution(world // M, M)
        ops = []
        for l, (ifm, ofm) in enumerate([(16, 32), (32, 32)]):
            ri = sess.create_operation_reg_info(OperationType.CC)
            ri.set_name("layer_%d" % l)
            ri.add_input(ifm, 9, DataType.FLOAT)
            ri.add_output(ofm, 9, DataType.FLOAT)
            ri.add_parameter_set(ifm * ofm, 4, DataType.FLOAT, True)
            ops.append(sess.get_operation(sess.add_operation(ri, dist)))
            sess.delete_operation_reg_info(ri)
        ops[1].set_prev(ops[0], 0, 0)
        sess.commit()
        oa, ia = ops[0].get_output(0), ops[1].get_input(0)
        lmb = ops[0].get_local_minibatch_size()
        n_out = oa.get_local_fm_count() * lmb * oa.get_fm_size()
        n_in = ia.get_local_fm_count() * lmb * ia.get_fm_size()
        out = torch.arange(n_out, dtype=torch.float32, device="cuda")
        comm_o = mlsl.tensor_from_address(oa.get_comm_buf(), (oa.get_comm_buf_size() // 4,), torch.float32)
        oa.pack(out, comm_o)
        oa.start_comm(comm_o)
        got = ia.wait_comm()
        recv = mlsl.tensor_from_address(got, (n_in,), torch.float32)
        inp = torch.zeros(n_in, device="cuda")
        ia.unpack(recv, inp)
        # backward: dIn = global index, must arrive as dOut[i] = i
        lfm, fs, off = ia.get_local_fm_count(), ia.get_fm_size(), ia.get_global_fm_offset()
        idx = torch.arange(n_in, device="cuda")
        mb, fm, s = idx // (lfm * fs), (idx // fs) % lfm, idx % fs
        din = (mb * lfm * fs * M + (off + fm) * fs + s).float()
        comm_i = mlsl.tensor_from_address(ia.get_comm_buf(), (ia.get_comm_buf_size() // 4,), torch.float32)
        ia.pack(din, comm_i)
        ia.start_comm(comm_i)
        gback = oa.wait_comm()
        dout = torch.zeros(n_out, device="cuda")
        oa.unpack(mlsl.tensor_from_address(gback, (n_out,), torch.float32), dout)
        # gradients through the data group
        ps = ops[0].get_parameter_set(0)
        npar = ps.get_local_kernel_count() * ps.get_kernel_size()
        dw = mlsl.alloc_tensor(npar, torch.float32)
        dw.copy_(torch.arange(npar, dtype=torch.float32))
        ps.start_gradient_comm(dw)
        g = mlsl.tensor_from_address(ps.wait_gradient_comm(), (ps.get_owned_kernel_count() * ps.get_kernel_size(),), torch.float32)
        torch.cuda.current_stream().synchronize()
        res = (inp.cpu(), dout.cpu(), g.cpu(), ps.get_owned_kernel_offset() * ps.get_kernel_size(), lfm, fs, off, lmb)
        e.delete_session(sess)
        e.delete_distribution(dist)
        return res

    outs = _gpu(body, world)
    D = world // M
    for inp, dout, g, own_off, lfm, fs, off, lmb in outs:
        idx = torch.arange(inp.numel())
        mb, fm, s = idx // (lfm * fs), (idx // fs) % lfm, idx % fs
        assert torch.equal(inp, (M * (mb * lfm * fs * M + (off + fm) * fs + s)).float())
        assert torch.equal(dout, torch.arange(dout.numel(), dtype=torch.float32))
        assert torch.equal(g, D * (own_off + torch.arange(g.numel(), dtype=torch.float32)))


def test_host_resident_allreduce_is_pipelined_and_correct():
    """Pinned host buffers straight through the public call: chunked H2D -> all-reduce -> D2H pipeline."""
    world, n = 2, (80 << 20) // 4 + 1031          # > 2 chunks of 32 MiB, odd tail

    def body(r, mlsl):
        hin = (torch.arange(n, dtype=torch.float32) % 1000 + r).pin_memory()
        hout = torch.empty(n, dtype=torch.float32).pin_memory()
        mlsl.allreduce(hin, out=hout, scale=0.5)
        torch.cuda.current_stream().synchronize()
        return hout[::4099].clone(), hout[-5:].clone()

    outs = run_ranks(world, body, backend="cuda", env={"MLSL_HEAP_SIZE_GB": "0.5", "MLSL_WATCHDOG_SEC": "20"})
    base = torch.arange(n, dtype=torch.float32) % 1000
    ref = (base * 2 + 1) * 0.5
    for a, b in outs:
        assert torch.equal(a, ref[::4099]) and torch.equal(b, ref[-5:])


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16, torch.float64, torch.int32, torch.uint8])
def test_allreduce_low_latency_path(world, dtype):
    """<= 8 KiB: the flag-in-data (LL) kernel, on heap and on foreign buffers, several calls in a row (parity reuse)."""
    sizes = [1, 3, 257, 8192 // max(1, torch.empty((), dtype=dtype).element_size())]

    def body(r, mlsl):
        res = []
        for n in sizes:
            for op in ("sum", "max"):
                x = _make(r, n, dtype).cuda()
                y = torch.empty_like(x)
                mlsl.allreduce(x, out=y, op=op, scale=0.5 if (dtype.is_floating_point and op == "sum") else 1.0)
                h = mlsl.alloc_tensor(n, dtype)
                h.copy_(x)
                mlsl.allreduce(h, op=op)
                torch.cuda.current_stream().synchronize()
                res.append((y.cpu(), h.cpu()))
        return res

    outs = _gpu(body, world)
    k = 0
    for n in sizes:
        for op in ("sum", "max"):
            ins = [_make(r, n, dtype) for r in range(world)]
            ref = _ref_reduce(ins, op)
            if dtype == torch.uint8 and op == "sum":
                ref = ref % 256
            sc = 0.5 if (dtype.is_floating_point and op == "sum") else 1.0
            tol = 0 if not dtype.is_floating_point else {torch.float32: 1e-6, torch.float64: 1e-12}.get(dtype, 4e-2)
            for r in range(world):
                y, h = outs[r][k]
                assert torch.allclose(y.to(ref.dtype), ref * sc if sc != 1.0 else ref, rtol=tol, atol=tol * 4), (n, op)
                assert torch.allclose(h.to(ref.dtype), ref, rtol=tol, atol=tol * 4), (n, op)
                assert torch.equal(y, outs[0][k][0]) and torch.equal(h, outs[0][k][1])
            k += 1


def test_heap_tensor_explicit_free_then_recycled_address():
    """free_tensor() + an allocation that recycles the address while the old tensor object is still alive: its later
    garbage collection must not release the new owner's block (this made two live tensors alias)."""
    def body(r, mlsl):
        outs = []
        for it in range(3):
            x = mlsl.alloc_tensor(1000, torch.float32)
            x.fill_(float(r + 1))
            y = mlsl.alloc_tensor(1000, torch.float32)
            mlsl.allreduce(x, out=y)
            torch.cuda.current_stream().synchronize()
            outs.append((x.data_ptr() != y.data_ptr(), float(y[0]), float(x[0])))
            mlsl.free_tensor(x)
            mlsl.free_tensor(y)
        return outs

    for r, outs in enumerate(_gpu(body, 2)):
        for distinct, ysum, xval in outs:
            assert distinct and ysum == 3.0 and xval == float(r + 1)


@pytest.mark.parametrize("wait_mode", ["host", "stream"])
def test_statistics_carry_device_timed_durations(wait_mode, tmp_path):
    """MLSL_STATS=1 on the device: an event pair around every kernel gives the duration of the collective ON THE DEVICE
    (Statistics.get_device_comm_nanos); with stream-ordered waits that duration is what the comm counters carry.  The Chrome
    trace (MLSL_TRACE_FILE) records it per request as `device_us`."""
    import json
    import os
    world, n = 2, 1 << 20
    trace = str(tmp_path / "trace")

    def body(r, mlsl):
        from mlsl_b200.api import DataType, OperationType
        e = mlsl.env()
        sess = e.create_session()
        sess.set_global_minibatch_size(world)
        dist = e.create_distribution(world, 1)
        ri = sess.create_operation_reg_info(OperationType.CC)
        ri.add_input(8, 1, DataType.FLOAT)
        ri.add_output(8, 1, DataType.FLOAT)
        ri.add_parameter_set(n, 1, DataType.FLOAT, False)
        op = sess.get_operation(sess.add_operation(ri, dist))
        sess.commit()
        st = sess.get_stats()
        iso = st.get_total_isolation_comm_cycles()
        st.start()
        ps = op.get_parameter_set(0)
        g = mlsl.alloc_tensor(n, torch.float32)
        g.fill_(1.0)
        for _ in range(4):
            ps.start_gradient_comm(g)
            ps.wait_gradient_comm()
            torch.cuda.current_stream().synchronize()
        ps.start_gradient_comm(g)                 # the fifth run harvests the fourth in stream mode
        ps.wait_gradient_comm()
        torch.cuda.current_stream().synchronize()
        st.stop()
        res = (iso, st.get_device_comm_nanos(0), st.get_comm_nanos(0), float(g[0]))
        st.print()
        e.delete_session(sess)
        e.delete_distribution(dist)
        return res

    outs = _gpu(body, world, wait_mode=wait_mode, env={"MLSL_STATS": "1", "MLSL_STATS_ITERS": "3", "MLSL_STATS_SKIP": "1",
                                                       "MLSL_TRACE_FILE": trace})
    for iso, dev_ns, comm_ns, v in outs:
        assert v == 2.0 ** 5
        assert iso > 0
        assert dev_ns > 3 * 2000, dev_ns            # >= 3 harvested runs of a 4 MiB all-reduce, microseconds each
        assert comm_ns >= (dev_ns if wait_mode == "stream" else 1)
    tr = json.load(open(trace + ".0.json"))
    evs = [ev for ev in tr["traceEvents"] if ev.get("ph") == "X" and ev["name"] == "AllReduce"]
    assert evs and any(ev["args"].get("device_us", 0) > 0 for ev in evs)
    if os.path.exists("mlsl_stats.log"):
        os.remove("mlsl_stats.log")
