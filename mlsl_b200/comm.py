"""Tensor-level API: `mlsl.init()`, `mlsl.allreduce(t)`, `mlsl.allgather(...)`, symmetric-heap tensors.

The thin PyTorch face of the library.  Every call goes straight to the native runtime through
the C API; on the CUDA backend the collective is ordered after the work already queued on torch's current stream and
`Work.wait()` orders the current stream after the collective (no host blocking), like torch.distributed.
"""
import ctypes
import threading

import torch

from . import api
from .api import CompressionType, DataType, GroupType, ReductionType

_tls = threading.local()
_process_state = {}

_TORCH2MLSL = {
    torch.float32: DataType.FLOAT,
    torch.float64: DataType.DOUBLE,
    torch.uint8: DataType.BYTE,
    torch.int8: DataType.BYTE,
    torch.bfloat16: DataType.BF16,
    torch.float16: DataType.FP16,
    torch.int32: DataType.INT32,
}
_OPS = {"sum": ReductionType.SUM, "min": ReductionType.MIN, "max": ReductionType.MAX}
_GROUPS = {"data": GroupType.DATA, "model": GroupType.MODEL, "global": GroupType.GLOBAL}


def _state():
    ov = getattr(_tls, "override", None)
    if ov is not None:
        return ov
    if getattr(_tls, "bound", False):
        return _tls.__dict__
    return _process_state


class use_state:
    """Run a block with the library state captured on another thread (`comm._state()` there).  Autograd executes the
    backward of CUDA functions on its own device thread, which is not bound to any in-process virtual rank."""

    def __init__(self, state):
        self.state = state

    def __enter__(self):
        self.prev = getattr(_tls, "override", None)
        _tls.override = self.state
        return self

    def __exit__(self, *exc):
        _tls.override = self.prev
        return False


def bind_thread_state():
    """Give the calling thread its own library state (in-process virtual ranks call this before init())."""
    _tls.bound = True


def mlsl_dtype(t):
    try:
        return _TORCH2MLSL[t]
    except KeyError:
        raise TypeError("dtype %s is not supported by mlsl_b200" % t)


class Work:
    """Handle of an in-flight collective."""

    def __init__(self, env, req, result=None, keep=()):
        self._env, self._req, self.result, self._keep = env, req, result, keep

    def wait(self):
        if self._req is not None:
            self._env.wait(self._req)
            self._req = None
        return self.result

    def is_completed(self):
        if self._req is None:
            return True
        if self._env.test(self._req):
            self._req = None
            return True
        return False


def init(wait_mode=None):
    """Initialise the library for this process (or this virtual rank).  Returns the Environment (api.MLSL)."""
    st = _state()
    if st.get("env") is not None:
        return st["env"]
    env = api.MLSL()
    env.init()
    st["env"] = env
    st["device"] = env.is_device_backend()
    st["world_dist"] = None
    st["live"] = {}
    st["deferred"] = []
    st["stream"] = None
    if st["device"]:
        env.set_wait_mode(wait_mode or "stream")
        api._stream_hook = _sync_stream
        st["ev_pool"] = [torch.cuda.Event() for _ in range(64)]
        for ev in st["ev_pool"]:
            ev.record()          # torch creates the CUDA event lazily, on the first record
    return env


def finalize():
    st = _state()
    env = st.get("env")
    if env is None:
        return
    if st.get("world_dist") is not None:
        env.delete_distribution(st["world_dist"])
    if st.get("device"):
        torch.cuda.synchronize()
    for _, ptr in st.get("deferred", []):
        env.free(ptr)
    st["deferred"] = []
    for ptr in list(st["live"].keys()):
        env.free(ptr)
    st["live"].clear()
    env.finalize()
    st["env"] = None


def env():
    e = _state().get("env")
    if e is None:
        raise RuntimeError("mlsl_b200 is not initialised: call mlsl_b200.init() first")
    return e


def is_initialized():
    return _state().get("env") is not None


def rank():
    return env().get_process_idx()


def world_size():
    return env().get_process_count()


def is_device():
    return bool(_state().get("device"))


def world_distribution():
    """Distribution(world, 1): every rank is a data-parallel replica."""
    st = _state()
    if st.get("world_dist") is None:
        st["world_dist"] = env().create_distribution(world_size(), 1)
    return st["world_dist"]


class _CudaMem:
    """Minimal CUDA-array-interface carrier so torch can wrap library memory without copying.  torch keeps this object
    alive for as long as any tensor / view uses the storage; when `owner` is given, dropping the last of them returns
    the block to the symmetric heap - deferred until the work queued on the stream at that moment has finished."""

    def __init__(self, ptr, nbytes, owner=None):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        self._ptr, self._owner = ptr, owner
        if owner is not None:
            owner["live"][ptr] = id(self)     # the block belongs to THIS carrier (the address may be recycled later)

    def __del__(self):
        st = self._owner
        if st is None or st.get("env") is None or st.get("live", {}).get(self._ptr) != id(self):
            return                            # freed explicitly (free_tensor) or the library is already finalised
        try:
            del st["live"][self._ptr]
            # pooled event (created at init): cuEventCreate can block behind another in-process rank's pending
            # pageable copy (class 4) - recording an existing event never does
            pool = st.get("ev_pool")
            ev = pool.pop() if pool else torch.cuda.Event()
            ev.record()
            st["deferred"].append((ev, self._ptr))
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


def _sweep_deferred(st):
    """Free heap blocks whose last user has been garbage collected and whose stream work has completed."""
    pend = st.get("deferred")
    if not pend:
        return
    keep = []
    for ev, ptr in pend:
        if ev.query():
            st["env"].free(ptr)
            st.setdefault("ev_pool", []).append(ev)
        else:
            keep.append((ev, ptr))
    st["deferred"] = keep


def tensor_from_address(ptr, shape, dtype, device=None, _owner=None):
    """Zero-copy torch view of `numel*itemsize` bytes at `ptr` (device memory on the CUDA backend, host otherwise)."""
    numel = 1
    for s in shape:
        numel *= int(s)
    nbytes = max(numel, 1) * torch.empty((), dtype=dtype).element_size()
    if is_device():
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        raw = torch.as_tensor(_CudaMem(ptr, nbytes, _owner), device=dev)
    else:
        buf = (ctypes.c_uint8 * nbytes).from_address(ptr)
        raw = torch.frombuffer(buf, dtype=torch.uint8)
    return raw.view(dtype)[:numel].view(*shape)


def alloc_tensor(shape, dtype=torch.float32, zero=True):
    """Allocate a tensor in the symmetric heap: peers' kernels read/write it directly (zero-copy collectives)."""
    if isinstance(shape, int):
        shape = (shape,)
    numel = 1
    for s in shape:
        numel *= int(s)
    nbytes = max(numel, 1) * torch.empty((), dtype=dtype).element_size()
    st = _state()
    _sweep_deferred(st)
    ptr = env().alloc(nbytes, 256)
    st["live"][ptr] = 0
    t = tensor_from_address(ptr, tuple(shape), dtype, _owner=st if st.get("device") else None)
    if zero:
        t.zero_()
    return t


_HEAP_POOL = {}


def heap_pool():
    """Context manager: torch CUDA allocations made inside it come from this rank's symmetric heap (a torch.cuda.MemPool on a
    CUDAPluggableAllocator that calls Environment::Alloc / Free), so collectives on them - DDP / FSDP buckets, the "mlsl"
    torch.distributed backend, optimizer flats - run zero-copy instead of being staged through heap scratch.

        with mlsl_b200.heap_pool():
            model = DDP(model.cuda())            # gradient buckets live in the heap from now on
    """
    if not is_device():
        import contextlib
        return contextlib.nullcontext()
    st = _state()
    pool = st.get("heap_pool")
    if pool is None:
        from . import _lib
        alloc = _HEAP_POOL.get("alloc")
        if alloc is None:
            alloc = torch.cuda.memory.CUDAPluggableAllocator(_lib.LIB_PATH, "mlsl_heap_malloc", "mlsl_heap_free")
            _HEAP_POOL["alloc"] = alloc
        pool = torch.cuda.MemPool(alloc.allocator())
        st["heap_pool"] = pool
    return torch.cuda.use_mem_pool(pool)


def free_tensor(t):
    ptr = t.data_ptr()
    st = _state()
    if ptr in st["live"]:
        del st["live"][ptr]
        env().free(ptr)


def _sync_stream():
    st = _state()
    if st.get("device") and st.get("env") is not None:
        cur = torch.cuda.current_stream().cuda_stream
        if st.get("stream") != cur:
            st["env"].set_stream(cur)
            st["stream"] = cur


def _prep(t):
    if not t.is_contiguous():
        raise ValueError("mlsl_b200 collectives need contiguous tensors")
    if t.is_cuda and not is_device():
        raise TypeError("a CUDA tensor was passed but the %s backend works on host memory: initialise with "
                        "MLSL_BACKEND=cuda, or move the tensor to the CPU" % env().get_backend_name())
    return t


def _movable(t):
    """Pure data-movement collectives accept any dtype: unsupported ones travel as bytes."""
    return t if t.dtype in _TORCH2MLSL else t.view(torch.uint8)


def _dist(distribution):
    return distribution if distribution is not None else world_distribution()


def _enter(tensor, distribution):
    """Common prologue of the tensor-level collectives: (state, distribution) with one state lookup."""
    st = _state()
    if st.get("env") is None:
        raise RuntimeError("mlsl_b200 is not initialised: call mlsl_b200.init() first")
    if not tensor.is_contiguous() or (tensor.is_cuda and not st["device"]):
        _prep(tensor)                                   # raises the descriptive error
    return st, (distribution if distribution is not None else (st.get("world_dist") or world_distribution()))


def _finish(st, req, result, keep, async_op):
    e = st["env"]
    if async_op:
        return Work(e, req, result, keep)
    e.wait(req)
    return result


def _group(group):
    if isinstance(group, str):
        if group not in _GROUPS:
            raise ValueError("unknown group %r (expected one of %s)" % (group, ", ".join(sorted(_GROUPS))))
        return _GROUPS[group]
    return group


def _op(op):
    if op not in _OPS:
        raise ValueError("unknown reduction %r (expected sum, min or max)" % (op,))
    return _OPS[op]


def _check_out(what, out, numel, dtype):
    if out.numel() != numel or out.dtype != dtype:
        raise ValueError("%s: `out` must hold %d elements of %s (got %d of %s)" % (what, numel, dtype, out.numel(), out.dtype))
    return _prep(out)


def allreduce(tensor, op="sum", group="data", scale=1.0, compress=False, out=None, async_op=False, distribution=None):
    """All-reduce `tensor` (in place unless `out` is given).  `scale` and the optional fp8 `compress`ed transport are
    fused into the reduction kernel."""
    st = _state()
    e = st.get("env")
    if e is None:
        raise RuntimeError("mlsl_b200 is not initialised: call mlsl_b200.init() first")
    if not tensor.is_contiguous() or (tensor.is_cuda and not st["device"]):
        _prep(tensor)                                   # raises the descriptive error
    n = tensor.numel()
    if out is None:
        out = tensor
    elif out is not tensor:
        _check_out("allreduce", out, n, tensor.dtype)
    ct = CompressionType.QUANTIZATION if compress else CompressionType.NONE
    d = distribution if distribution is not None else (st.get("world_dist") or world_distribution())
    if not async_op:      # start + wait in one native call (the stream is bound to torch's current one inside)
        d.all_reduce_blocking(e, tensor, out, n, mlsl_dtype(tensor.dtype), _op(op), _group(group), float(scale), ct)
        return out
    req = _dist(distribution).all_reduce_ex(tensor, out, tensor.numel(), mlsl_dtype(tensor.dtype), _op(op),
                                            _group(group), float(scale), ct)
    return Work(env(), req, out, (tensor, out))


def reduce_scatter(tensor, out=None, op="sum", group="data", scale=1.0, async_op=False, distribution=None):
    """tensor: P*n elements; returns rank's n-element reduced shard."""
    st, d = _enter(tensor, distribution)
    g = _group(group)
    P = d.get_process_count(g)
    if tensor.numel() % P:
        raise ValueError("reduce_scatter: %d elements cannot be split over %d ranks" % (tensor.numel(), P))
    n = tensor.numel() // P
    if out is None:
        out = alloc_tensor((n,), tensor.dtype, zero=False) if st["device"] else torch.empty(n, dtype=tensor.dtype)
    else:
        _check_out("reduce_scatter", out, n, tensor.dtype)
    req = d.reduce_scatter(tensor, out, n, mlsl_dtype(tensor.dtype), _op(op), g, float(scale))
    return _finish(st, req, out, (tensor, out), async_op)


def allgather(tensor, out=None, group="data", async_op=False, distribution=None):
    st, d = _enter(tensor, distribution)
    g = _group(group)
    n = tensor.numel()
    P = d.get_process_count(g)
    if out is None:
        out = alloc_tensor((P * n,), tensor.dtype, zero=False) if st["device"] else torch.empty(P * n, dtype=tensor.dtype)
    else:
        _check_out("allgather", out, P * n, tensor.dtype)
    req = d.all_gather(tensor, n, out, mlsl_dtype(tensor.dtype), g)
    return _finish(st, req, out, (tensor, out), async_op)


def allgatherv(tensor, recv_counts, out=None, group="data", async_op=False, distribution=None):
    """Distribution::AllGatherv: member p contributes `recv_counts[p]` elements (this rank: all of `tensor`); returns the
    concatenation in member order."""
    _prep(tensor)
    d = _dist(distribution)
    g = _group(group)
    P, idx = d.get_process_count(g), d.get_process_idx(g)
    recv_counts = [int(c) for c in recv_counts]
    if len(recv_counts) != P or recv_counts[idx] != tensor.numel():
        raise ValueError("allgatherv: need %d counts and counts[%d] == %d (this rank's elements)" % (P, idx, tensor.numel()))
    total = sum(recv_counts)
    if out is None:
        out = torch.empty(max(total, 1), dtype=tensor.dtype, device=tensor.device)[:total]
    else:
        _check_out("allgatherv", out, total, tensor.dtype)
    _sync_stream()
    req = d.all_gatherv(tensor, tensor.numel(), out, recv_counts, mlsl_dtype(tensor.dtype), g)
    w = Work(env(), req, out, (tensor, out))
    return w if async_op else w.wait()


def alltoall(tensor, out=None, group="data", async_op=False, distribution=None):
    st, d = _enter(tensor, distribution)
    g = _group(group)
    P = d.get_process_count(g)
    if tensor.numel() % P:
        raise ValueError("alltoall: %d elements cannot be split over %d ranks" % (tensor.numel(), P))
    if out is None:
        out = alloc_tensor(tuple(tensor.shape), tensor.dtype, zero=False) if st["device"] else torch.empty_like(tensor)
    else:
        _check_out("alltoall", out, tensor.numel(), tensor.dtype)
    req = d.all_to_all(tensor, tensor.numel() // P, out, mlsl_dtype(tensor.dtype), g)
    return _finish(st, req, out, (tensor, out), async_op)


def alltoallv(tensor, send_counts, recv_counts=None, out=None, group="data", async_op=False, distribution=None):
    """Variable all-to-all (Distribution::AlltoAllv): `send_counts[p]` consecutive elements of `tensor` go to rank p;
    returns the concatenation of what the ranks sent here, in rank order.  `recv_counts` are exchanged first when the
    caller does not know them (one small all-to-all)."""
    _prep(tensor)
    d = _dist(distribution)
    g = _group(group)
    P = d.get_process_count(g)
    send_counts = [int(c) for c in send_counts]
    if len(send_counts) != P or sum(send_counts) > tensor.numel():
        raise ValueError("alltoallv: need %d send counts that fit the %d elements of the tensor" % (P, tensor.numel()))
    if recv_counts is None:
        mine = torch.tensor(send_counts, dtype=torch.int32)
        theirs = torch.empty(P, dtype=torch.int32)
        if is_device():
            mine, theirs = mine.to(tensor.device), theirs.to(tensor.device)
        alltoall(mine, out=theirs, group=group, distribution=distribution)
        recv_counts = theirs.tolist()
    recv_counts = [int(c) for c in recv_counts]
    total = sum(recv_counts)
    if out is None:
        out = torch.empty(max(total, 1), dtype=tensor.dtype, device=tensor.device)[:total]
    elif out.numel() < total or out.dtype != tensor.dtype:
        raise ValueError("alltoallv: `out` must hold %d elements of %s" % (total, tensor.dtype))
    so = [sum(send_counts[:p]) for p in range(P)]
    ro = [sum(recv_counts[:p]) for p in range(P)]
    _sync_stream()
    req = d.all_to_allv(tensor, send_counts, so, out, recv_counts, ro, mlsl_dtype(tensor.dtype), g)
    w = Work(env(), req, out, (tensor, out))
    return w if async_op else w.wait()


def ring_shift(tensor, shift=1, out=None, group="data", async_op=False, distribution=None):
    """out on rank i = tensor of rank (i - shift) mod P: every rank sends its block `shift` positions up the ring.
    One SendRecvList operation (the reference declares that op but never wires it up, src/comm.hpp:212-248); this is the
    KV rotation of ring attention and the neighbour exchange of pipeline schedules."""
    _prep(tensor)
    d = _dist(distribution)
    g = _group(group)
    P, idx = d.get_process_count(g), d.get_process_idx(g)
    n = tensor.numel()
    if out is None:
        out = alloc_tensor(tuple(tensor.shape), tensor.dtype, zero=False) if is_device() else torch.empty_like(tensor)
    dst, src = (idx + shift) % P, (idx - shift) % P
    sc = [n if p == dst else 0 for p in range(P)]
    rc = [n if p == src else 0 for p in range(P)]
    _sync_stream()
    req = d.send_recv_list(tensor, sc, [0] * P, out, rc, [0] * P, mlsl_dtype(tensor.dtype), g)
    w = Work(env(), req, out, (tensor, out))
    return w if async_op else w.wait()


def bcast(tensor, root=0, group="data", async_op=False, distribution=None):
    st, d = _enter(tensor, distribution)
    raw = tensor if (tensor.dim() == 1 and tensor.dtype in _TORCH2MLSL) else _movable(tensor.view(-1))
    req = d.bcast(raw, raw.numel(), mlsl_dtype(raw.dtype), root, _group(group))
    return _finish(st, req, tensor, (tensor, raw), async_op)


def reduce(tensor, out=None, root=0, op="sum", group="data", async_op=False, distribution=None):
    st, d = _enter(tensor, distribution)
    out = tensor if out is None else _check_out("reduce", out, tensor.numel(), tensor.dtype)
    req = d.reduce(tensor, out, tensor.numel(), mlsl_dtype(tensor.dtype), _op(op), root, _group(group))
    return _finish(st, req, out, (tensor, out), async_op)


def gather(tensor, out=None, root=0, group="data", async_op=False, distribution=None):
    """Distribution::Gather: rank `root` receives the concatenation of every member's `tensor` (in member order) in `out`
    (allocated when missing); the other ranks get None."""
    _prep(tensor)
    d = _dist(distribution)
    g = _group(group)
    P, idx = d.get_process_count(g), d.get_process_idx(g)
    raw = _movable(tensor.view(-1))
    if idx == root:
        if out is None:
            out = torch.empty(P * tensor.numel(), dtype=tensor.dtype, device=tensor.device)
        else:
            _check_out("gather", out, P * tensor.numel(), tensor.dtype)
        raw_out = _movable(out.view(-1))
    else:
        out, raw_out = None, raw          # not written on the other ranks; a valid address keeps the pointer checks quiet
    _sync_stream()
    req = d.gather(raw, raw.numel(), raw_out, mlsl_dtype(raw.dtype), root, g)
    w = Work(env(), req, out, (tensor, raw, raw_out))
    return w if async_op else w.wait()


def scatter(tensor, out=None, root=0, group="data", async_op=False, distribution=None):
    """Distribution::Scatter: member i receives block i of rank `root`'s `tensor` (P equal blocks).  On the other ranks
    `tensor` only tells the block's dtype / device; pass `out` (or a tensor shaped like one block)."""
    d = _dist(distribution)
    g = _group(group)
    P, idx = d.get_process_count(g), d.get_process_idx(g)
    if idx == root:
        _prep(tensor)
        if tensor.numel() % P:
            raise ValueError("scatter: %d elements cannot be split over %d ranks" % (tensor.numel(), P))
        n = tensor.numel() // P
    else:
        n = out.numel() if out is not None else tensor.numel()
    if out is None:
        out = torch.empty(n, dtype=tensor.dtype, device=tensor.device)
    else:
        _check_out("scatter", out, n, tensor.dtype)
    raw_out = _movable(out.view(-1))
    raw_in = _movable(tensor.view(-1)) if idx == root else raw_out
    _sync_stream()
    req = d.scatter(raw_in, raw_out, raw_out.numel(), mlsl_dtype(raw_out.dtype), root, g)
    w = Work(env(), req, out, (tensor, raw_in, raw_out))
    return w if async_op else w.wait()


def barrier(group="global", distribution=None):
    _sync_stream()
    _dist(distribution).barrier(_group(group))
