"""`torch.distributed` backend "mlsl": run unmodified PyTorch distributed code (all_reduce, DDP, FSDP-style
all_gather_into_tensor / reduce_scatter_tensor, new_group) on this library's collectives.

    import mlsl_b200.torch_backend            # registers the backend
    torch.distributed.init_process_group("mlsl", ...)

The reference stops at its own Distribution/Operation API (reference include/mlsl.hpp:1-900); frameworks that wanted it
underneath had to be patched (Intel Caffe, reference README.md).  Here the PyTorch side needs no patch: the default group
maps to the world distribution, `new_group(ranks)` to a distribution created by its members only (the torch store is the
rendezvous, Environment::CreateDistributionFromRanks), and every collective is one call into `mlsl_b200.comm`, so on the
CUDA backend it is stream-ordered like ProcessGroupNCCL (`Work.wait()` orders the current stream, it does not block the
host).  send / recv between two ranks are one SendRecvList operation (the op the reference declares but never wires up,
src/comm.hpp:212-248) on a two-member distribution made for the pair on first use.
"""
import struct

import torch
import torch.distributed as dist
from torch.distributed import ReduceOp

from . import comm
from .api import GroupType

BACKEND_NAME = "mlsl"
_NATIVE_RED = {ReduceOp.SUM: "sum", ReduceOp.MIN: "min", ReduceOp.MAX: "max"}
_LOCAL_RED = {
    ReduceOp.SUM: lambda s: s.sum(0),
    ReduceOp.AVG: lambda s: s.sum(0) / s.shape[0] if s.is_floating_point() else s.sum(0) // s.shape[0],
    ReduceOp.PRODUCT: lambda s: s.prod(0),
    ReduceOp.MIN: lambda s: s.amin(0),
    ReduceOp.MAX: lambda s: s.amax(0),
    ReduceOp.BAND: lambda s: _fold(s, torch.bitwise_and),
    ReduceOp.BOR: lambda s: _fold(s, torch.bitwise_or),
    ReduceOp.BXOR: lambda s: _fold(s, torch.bitwise_xor),
}


def _fold(stack, fn):
    acc = stack[0].clone()
    for i in range(1, stack.shape[0]):
        acc = fn(acc, stack[i])
    return acc


def _red_key(op):
    """ReduceOp objects compare equal to the enum members but do not hash like them."""
    for k in _LOCAL_RED:
        if op == k:
            return k
    raise NotImplementedError("mlsl backend: reduction %s is not supported" % (op,))


class _MLSLWork(dist._Work):
    """Completion handle.  `pending` are comm.Work handles, `finish` copies staged results back to the caller's tensors."""

    def __init__(self, pending, result, finish=None):
        super().__init__()
        self._pending, self._result, self._finish, self._fut = list(pending), result, finish, None

    def wait(self, timeout=None):
        for w in self._pending:
            w.wait()
        self._pending = []
        if self._finish is not None:
            fin, self._finish = self._finish, None
            fin()
        return True

    def is_completed(self):
        if all(w.is_completed() for w in self._pending):
            self.wait()          # nothing left to wait for: run the copy-backs
            return True
        return False

    def is_success(self):
        return True

    def exception(self):
        return None

    def synchronize(self):
        self.wait()

    def result(self):
        self.wait()
        return self._result

    def get_future(self):
        if self._fut is None:
            self.wait()
            devs = sorted({t.device for t in _flatten(self._result) if t.is_cuda}, key=str)
            self._fut = torch.futures.Future(devices=devs) if devs else torch.futures.Future()
            self._fut.set_result(self._result)
        return self._fut


def _flatten(x):
    if isinstance(x, torch.Tensor):
        return [x]
    out = []
    for y in x or ():
        out.extend(_flatten(y))
    return out


class MLSLProcessGroup(dist.ProcessGroup):
    def __init__(self, rank, size, distribution, owns_distribution, owns_library, store=None, ranks=None):
        super().__init__(rank, size)
        self._d, self._owns_d, self._owns_lib = distribution, owns_distribution, owns_library
        self._state = comm._state()   # collectives may be issued from autograd's threads (DDP hooks)
        self._store, self._ranks = store, list(ranks) if ranks is not None else list(range(size))
        self._pairs = {}              # peer (group rank) -> (two-member distribution, my index in it)

    # ---- plumbing -------------------------------------------------------------------------------------------
    def getBackendName(self):
        return BACKEND_NAME

    # torch keeps the group name on the registered C++ backends; a Python process group has none
    def _set_group_name(self, name):
        self._name = name
        super()._set_group_name(name)

    @property
    def group_name(self):
        return self._name

    def _set_group_desc(self, desc):
        self._desc = desc
        super()._set_group_desc(desc)

    @property
    def group_desc(self):
        return self._desc

    def _kw(self):
        return {"group": "data", "distribution": self._d, "async_op": True}

    def _staged(self, t):
        """(contiguous tensor to communicate on, copy-back closure or None)"""
        if t.is_contiguous():
            return t, None
        c = t.contiguous()
        return c, (lambda: t.copy_(c))

    def _done(self, pending, result, finishers=()):
        fins = [f for f in finishers if f is not None]
        finish = (lambda: [f() for f in fins]) if fins else None
        return _MLSLWork(pending, result, finish)

    def _gather_flat(self, flat):
        """All-gather a flat contiguous tensor of any dtype -> [P, n] tensor (blocking on the host backend, stream-ordered
        on the CUDA backend)."""
        P = self.size()
        raw = flat.view(torch.uint8) if flat.dtype not in comm._TORCH2MLSL else flat
        out = torch.empty(P * raw.numel(), dtype=raw.dtype, device=raw.device)
        comm.allgather(raw, out=out, group="data", distribution=self._d)
        return out.view(flat.dtype).view(P, flat.numel())

    # ---- collectives ----------------------------------------------------------------------------------------
    def allreduce(self, tensors, opts=None):
        op = _red_key(opts.reduceOp) if opts is not None else ReduceOp.SUM
        with comm.use_state(self._state):
            pending, fins = [], []
            for t in tensors:
                c, back = self._staged(t)
                if self.size() == 1:
                    pass
                elif c.dtype in comm._TORCH2MLSL and c.dtype not in (torch.uint8, torch.int8) and (op in _NATIVE_RED or (
                        op == ReduceOp.AVG and c.is_floating_point())):
                    if op == ReduceOp.AVG:
                        pending.append(comm.allreduce(c.view(-1), op="sum", scale=1.0 / self.size(), **self._kw()))
                    else:
                        pending.append(comm.allreduce(c.view(-1), op=_NATIVE_RED[op], **self._kw()))
                else:   # exact for every dtype and operator: gather, reduce locally
                    c.view(-1).copy_(_LOCAL_RED[op](self._gather_flat(c.view(-1))))
                fins.append(back)
            return self._done(pending, list(tensors), fins)

    def allreduce_coalesced(self, tensors, opts=None):
        return self.allreduce(tensors, opts)

    def broadcast(self, tensors, opts=None):
        root = opts.rootRank if opts is not None else 0
        with comm.use_state(self._state):
            pending, fins = [], []
            for t in tensors:
                c, back = self._staged(t)
                if self.size() > 1 and c.numel():
                    pending.append(comm.bcast(c, root=root, **self._kw()))
                fins.append(back if self.rank() != root else None)
            return self._done(pending, list(tensors), fins)

    def _allgather_base(self, output, input, opts=None):
        with comm.use_state(self._state):
            o, back = self._staged(output)
            i = input.contiguous()
            if self.size() == 1:
                o.view(-1).copy_(i.view(-1))
                pending = []
            else:
                raw_i = i.view(-1) if i.dtype in comm._TORCH2MLSL else i.view(-1).view(torch.uint8)
                raw_o = o.view(-1) if o.dtype in comm._TORCH2MLSL else o.view(-1).view(torch.uint8)
                pending = [comm.allgather(raw_i, out=raw_o, **self._kw())]
            return self._done(pending, output, [back])

    def allgather(self, output_tensors, input_tensors, opts=None):
        with comm.use_state(self._state):
            pending, fins = [], []
            for outs, inp in zip(output_tensors, input_tensors):
                i = inp.contiguous()
                flat = torch.empty((self.size(),) + tuple(i.shape), dtype=i.dtype, device=i.device)
                w = self._allgather_base(flat, i)
                pending.append(w)
                fins.append(lambda outs=outs, flat=flat: [o.copy_(flat[k]) for k, o in enumerate(outs)])
            return self._done(pending, [list(o) for o in output_tensors], fins)

    def allgather_into_tensor_coalesced(self, outputs, inputs, opts=None):
        works = [self._allgather_base(o, i) for o, i in zip(outputs, inputs)]
        return self._done(works, list(outputs))

    def allgather_coalesced(self, output_lists, input_list, opts=None):
        return self.allgather(output_lists, input_list, opts)

    def _reduce_scatter_base(self, output, input, opts=None):
        op = _red_key(opts.reduceOp) if opts is not None else ReduceOp.SUM
        P = self.size()
        with comm.use_state(self._state):
            o, back = self._staged(output)
            i = input.contiguous().view(-1)
            if i.numel() != P * o.numel():
                raise ValueError("reduce_scatter: input must hold world_size * output elements")
            if P == 1:
                o.view(-1).copy_(i)
                pending = []
            elif i.dtype in comm._TORCH2MLSL and i.dtype not in (torch.uint8, torch.int8) and (op in _NATIVE_RED or (
                    op == ReduceOp.AVG and i.is_floating_point())):
                scale = 1.0 / P if op == ReduceOp.AVG else 1.0
                pending = [comm.reduce_scatter(i, out=o.view(-1), op=_NATIVE_RED.get(op, "sum"), scale=scale, **self._kw())]
            else:
                n = o.numel()
                full = _LOCAL_RED[op](self._gather_flat(i))
                o.view(-1).copy_(full[self.rank() * n:(self.rank() + 1) * n])
                pending = []
            return self._done(pending, output, [back])

    def reduce_scatter(self, output_tensors, input_tensors, opts=None):
        works = []
        for out, ins in zip(output_tensors, input_tensors):
            works.append(self._reduce_scatter_base(out, torch.cat([x.reshape(-1) for x in ins]), opts))
        return self._done(works, list(output_tensors))

    def reduce_scatter_tensor_coalesced(self, outputs, inputs, opts=None):
        works = [self._reduce_scatter_base(o, i, opts) for o, i in zip(outputs, inputs)]
        return self._done(works, list(outputs))

    def reduce(self, tensors, opts=None):
        op = _red_key(opts.reduceOp) if opts is not None else ReduceOp.SUM
        root = opts.rootRank if opts is not None else 0
        with comm.use_state(self._state):
            pending, fins = [], []
            for t in tensors:
                c, back = self._staged(t)
                if self.size() == 1:
                    continue
                if c.dtype in comm._TORCH2MLSL and c.dtype not in (torch.uint8, torch.int8) and op in _NATIVE_RED:
                    # non-root ranks keep their input: reduce into scratch there
                    out = c.view(-1) if self.rank() == root else torch.empty_like(c.view(-1))
                    pending.append(comm.reduce(c.view(-1), out=out, root=root, op=_NATIVE_RED[op], **self._kw()))
                else:
                    full = _LOCAL_RED[op](self._gather_flat(c.view(-1)))
                    if self.rank() == root:
                        c.view(-1).copy_(full)
                fins.append(back if self.rank() == root else None)
            return self._done(pending, list(tensors), fins)

    def alltoall_base(self, output, input, output_split_sizes, input_split_sizes, opts=None):
        P = self.size()
        with comm.use_state(self._state):
            o, back = self._staged(output)
            i = input.contiguous()
            if P == 1:
                o.view(-1).copy_(i.view(-1))
                return self._done([], output, [back])
            row = i[0].numel() if i.dim() > 0 and i.shape[0] else 1
            esz = i.element_size() if i.dtype not in comm._TORCH2MLSL else 1
            raw_i = i.view(-1) if esz == 1 else i.view(-1).view(torch.uint8)
            raw_o = o.view(-1) if esz == 1 else o.view(-1).view(torch.uint8)
            if not output_split_sizes and not input_split_sizes:
                return self._done([comm.alltoall(raw_i, out=raw_o, **self._kw())], output, [back])
            ins = list(input_split_sizes) or [i.shape[0] // P] * P
            outs = list(output_split_sizes) or [o.shape[0] // P] * P
            sc = [s * row * esz for s in ins]
            rc = [s * row * esz for s in outs]
            so = [sum(sc[:k]) for k in range(P)]
            ro = [sum(rc[:k]) for k in range(P)]
            comm._prep(raw_i), comm._prep(raw_o)
            comm._sync_stream()
            req = self._d.send_recv_list(raw_i, sc, so, raw_o, rc, ro, comm.mlsl_dtype(raw_i.dtype), comm._group("data"))
            return self._done([comm.Work(comm.env(), req, raw_o, (raw_i, raw_o))], output, [back])

    def alltoall(self, output_tensors, input_tensors, opts=None):
        if not input_tensors:
            return self._done([], list(output_tensors))
        dt, dev = input_tensors[0].dtype, input_tensors[0].device
        src = torch.cat([t.reshape(-1) for t in input_tensors])
        dst = torch.empty(sum(t.numel() for t in output_tensors), dtype=dt, device=dev)
        w = self.alltoall_base(dst, src, [t.numel() for t in output_tensors], [t.numel() for t in input_tensors])

        def scatter_back():
            off = 0
            for t in output_tensors:
                t.copy_(dst[off:off + t.numel()].view(t.shape))
                off += t.numel()
        return self._done([w], list(output_tensors), [scatter_back])

    def gather(self, output_tensors, input_tensors, opts=None):
        root = opts.rootRank if opts is not None else 0
        with comm.use_state(self._state):
            pending, fins = [], []
            for k, inp in enumerate(input_tensors):
                flat = inp.contiguous().view(-1)
                if self.size() == 1:
                    output_tensors[k][0].copy_(inp)
                    continue
                w = comm.gather(flat, root=root, **self._kw())
                pending.append(w)
                if self.rank() == root:
                    fins.append(lambda w=w, outs=output_tensors[k], n=flat.numel(): [
                        o.copy_(w.result[r * n:(r + 1) * n].view(o.shape)) for r, o in enumerate(outs)])
            return self._done(pending, [list(o) for o in output_tensors], fins)

    def scatter(self, output_tensors, input_tensors, opts=None):
        root = opts.rootRank if opts is not None else 0
        with comm.use_state(self._state):
            pending, fins = [], []
            for k, out in enumerate(output_tensors):
                if self.size() == 1:
                    out.copy_(input_tensors[k][0])
                    continue
                o, back = self._staged(out)
                src = torch.cat([t.reshape(-1) for t in input_tensors[k]]) if self.rank() == root else o.view(-1)
                pending.append(comm.scatter(src, out=o.view(-1), root=root, **self._kw()))
                fins.append(back)
            return self._done(pending, list(output_tensors), fins)

    def barrier(self, opts=None):
        with comm.use_state(self._state):
            if self.size() > 1:
                comm.barrier(group="data", distribution=self._d)
            return self._done([], None)

    # ---- point to point -------------------------------------------------------------------------------------
    # A pair of ranks that exchanges messages gets a two-member distribution of its own the first time it does (created by
    # the two members only, the group's store is the rendezvous - like new_group); a send and the matching recv are then ONE
    # SendRecvList operation on it, so messages between a pair match in program order (tags are not consulted), both calls
    # complete together (rendezvous semantics), and batch_isend_irecv maps every (send, recv) couple of a pair to the same
    # operation on both sides.
    def _pair(self, peer):
        me = self.rank()
        if peer == me or not 0 <= peer < self.size():
            raise ValueError("mlsl backend: invalid peer rank %d for rank %d of %d" % (peer, me, self.size()))
        ent = self._pairs.get(peer)
        if ent is None:
            env = comm.env()
            lo, hi = min(me, peer), max(me, peer)
            key = "mlsl_p2p/%d_%d/%%d" % (lo, hi)
            rows, mark = env.get_group_state()
            self._store.set(key % me, struct.pack("<QQ", rows, mark))
            peer_rows, peer_mark = struct.unpack("<QQ", self._store.get(key % peer))
            d = env.create_distribution_from_ranks([self._ranks[lo], self._ranks[hi]], rows | peer_rows, max(mark, peer_mark))
            ent = self._pairs[peer] = (d, 0 if me == lo else 1)
        return ent

    def _p2p(self, tensor, peer, sending):
        with comm.use_state(self._state):
            if self._store is None:
                raise NotImplementedError("mlsl backend: this process group was made without a store: no point-to-point")
            d, idx = self._pair(peer)
            c, back = self._staged(tensor)
            raw = c.view(-1) if c.dtype in comm._TORCH2MLSL else c.view(-1).view(torch.uint8)
            n = raw.numel()
            mine = [0, 0]
            mine[1 - idx] = n
            sc, rc = (mine, [0, 0]) if sending else ([0, 0], mine)
            comm._sync_stream()
            req = d.send_recv_list(raw, sc, [0, 0], raw, rc, [0, 0], comm.mlsl_dtype(raw.dtype), GroupType.DATA)
            return self._done([comm.Work(comm.env(), req, None, (raw, c))], [tensor], [None if sending else back])

    def send(self, tensors, dstRank, tag):
        return self._p2p(tensors[0], dstRank, True)

    def recv(self, tensors, srcRank, tag):
        return self._p2p(tensors[0], srcRank, False)

    def recv_anysource(self, tensors, tag):
        raise NotImplementedError("mlsl backend: recv from any source is not provided (every message is a rendezvous of a pair)")

    # ---- lifetime -------------------------------------------------------------------------------------------
    def shutdown(self):
        """Called by torch.distributed.destroy_process_group (newest group first, the same order on every rank)."""
        with comm.use_state(self._state):
            if comm.is_initialized():
                me = self.rank()                          # collective over the pair: one global order of the pairs
                for peer in sorted(self._pairs, key=lambda q: (min(me, q), max(me, q))):
                    comm.env().delete_distribution(self._pairs[peer][0])
            self._pairs = {}
            if self._d is not None and self._owns_d and comm.is_initialized():
                comm.env().delete_distribution(self._d)   # collective over the members
            self._d = None
            if self._owns_lib and comm.is_initialized():
                comm.finalize()
                self._owns_lib = False

    abort = shutdown


def _create(opts, pg_options=None):
    """Backend constructor (extended API): only the members of the new group get here."""
    owns_lib = not comm.is_initialized()
    env = comm.init()
    size, rank = opts.group_size, opts.group_rank
    ranks = [int(r) for r in (opts.global_ranks_in_group or [])] or list(range(size))
    world = env.get_process_count()
    if size > world or max(ranks) >= world:
        raise RuntimeError("mlsl backend: group of %d ranks (max rank %d) does not fit the library's %d processes"
                           % (size, max(ranks), world))
    if ranks[rank] != env.get_process_idx():
        raise RuntimeError("mlsl backend: torch rank %d is library process %d - launch both from the same RANK/"
                           "WORLD_SIZE (mlslrun or torchrun)" % (ranks[rank], env.get_process_idx()))
    if ranks == list(range(world)):
        return MLSLProcessGroup(rank, size, comm.world_distribution(), False, owns_lib, opts.store, ranks)
    rows, mark = env.get_group_state()
    opts.store.set("mlsl_group_state/%d" % rank, struct.pack("<QQ", rows, mark))
    for r in range(size):
        peer_rows, peer_mark = struct.unpack("<QQ", opts.store.get("mlsl_group_state/%d" % r))
        rows |= peer_rows
        mark = max(mark, peer_mark)
    return MLSLProcessGroup(rank, size, env.create_distribution_from_ranks(ranks, rows, mark), True, owns_lib, opts.store, ranks)


def compressed_allreduce_hook(process_group, bucket):
    """DDP communication hook: average the gradient bucket with the library's quantised all-reduce (block-scaled FP8 with
    error feedback on the CUDA backend, the QuantParams plug-in or the built-in codec on the host backend - the
    reference's CT_QUANTIZATION, src/quant/quant.cpp:1-200), scale fused into the same operation.

        ddp.register_comm_hook(None, mlsl_b200.torch_backend.compressed_allreduce_hook)
    """
    pg = process_group if process_group is not None else dist.group.WORLD
    if not isinstance(pg, MLSLProcessGroup):
        raise TypeError("compressed_allreduce_hook needs a process group of the mlsl backend")
    buf = bucket.buffer()
    with comm.use_state(pg._state):
        if pg.size() > 1:
            comm.allreduce(buf, scale=1.0 / pg.size(), compress=buf.dtype == torch.float32, group="data",
                           distribution=pg._d, async_op=True).wait()
    fut = torch.futures.Future(devices=[buf.device]) if buf.is_cuda else torch.futures.Future()
    fut.set_result(buf)
    return fut


def register():
    if BACKEND_NAME.upper() not in getattr(dist.Backend, "_plugins", {}):
        dist.Backend.register_backend(BACKEND_NAME, _create, extended_api=True, devices=["cpu", "cuda"])


register()
