"""GPU jobs that span nodes: the library's kernels inside each NVSwitch domain, torch.distributed (NCCL over the node
interconnect; gloo on CPUs) between the domains.

The reference leaves node boundaries to MPI (one flat communicator, src/comm_ep.cpp:1791-1830).  On NVLink machines the
two levels differ by an order of magnitude in bandwidth, so the collective is split the way the hardware is:

    all-reduce  = reduce-scatter inside the node (peer-memory kernel, 1/N scale fused)
                  -> all-reduce of the 1/L shard between nodes, one group per local rank, all L groups in parallel
                  -> all-gather inside the node
    so every byte crosses the slow link once, as 1/L of the message per GPU.

Launch with torchrun on every node (RANK / WORLD_SIZE / LOCAL_RANK / LOCAL_WORLD_SIZE / MASTER_*):

    hc = mlsl_b200.parallel.multinode.init_hybrid()      # instead of mlsl.init()
    hc.allreduce(grad_bucket, scale=1.0 / hc.world_size)
"""
import os

import torch
import torch.distributed as dist

from .. import comm

_RED = {"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}


def _env_int(name, default=None):
    v = os.environ.get(name)
    if v is None:
        if default is None:
            raise RuntimeError("init_hybrid: %s is not set - start the job with torchrun (or export the torchrun variables)"
                               % name)
        return default
    return int(v)


class HybridComm:
    """Two-level collectives: `comm` (this library) inside the node, `inter` (a torch.distributed group of the ranks that
    share my local rank) between nodes."""

    def __init__(self, inter, node, nnodes, local_rank, local_world):
        self.inter, self.node, self.nnodes, self.local_rank, self.local_world = inter, node, nnodes, local_rank, local_world
        self.rank = node * local_world + local_rank
        self.world_size = nnodes * local_world

    def _scratch(self, n, like):
        if comm.is_device():
            return comm.alloc_tensor((n,), like.dtype, zero=False)
        return torch.empty(n, dtype=like.dtype)

    def allreduce(self, tensor, op="sum", scale=1.0):
        """In place.  `tensor`: contiguous, on the device of the backend (any length - the tail is handled by padding)."""
        if op not in _RED:
            raise ValueError("unknown reduction %r (expected sum, min or max)" % (op,))
        L = self.local_world
        flat = tensor.view(-1)
        if self.nnodes == 1:
            return comm.allreduce(flat, op=op, scale=scale)
        if L == 1:
            dist.all_reduce(flat, _RED[op], group=self.inter)
            if scale != 1.0:
                flat.mul_(scale)
            return tensor
        n = flat.numel()
        per = -(-n // L)
        if per * L != n:   # pad once so that every local rank owns an equal slice
            padded = self._scratch(per * L, flat)
            padded[:n].copy_(flat)
            padded[n:].zero_()
        else:
            padded = flat
        shard = self._scratch(per, flat)
        comm.reduce_scatter(padded, out=shard, op=op, scale=scale if op == "sum" else 1.0)
        dist.all_reduce(shard, _RED[op], group=self.inter)
        comm.allgather(shard, out=padded)
        if padded is not flat:
            flat.copy_(padded[:n])
            if comm.is_device():
                comm.free_tensor(padded)
        if comm.is_device():
            comm.free_tensor(shard)
        return tensor

    def bcast(self, tensor, root=0):
        """`root` is a global rank.  The root's node-mates get the data over NVLink, the other nodes over the column of the
        root's local rank, then inside their node."""
        root_local = root % self.local_world
        flat = tensor.view(-1)
        if self.nnodes > 1 and self.local_rank == root_local:
            dist.broadcast(flat, src=root, group=self.inter)
        if self.local_world > 1:
            comm.bcast(flat, root=root_local)
        return tensor

    def allgather(self, tensor, out=None):
        """out = concatenation over GLOBAL ranks (node-major).  Each rank's block crosses the node link once: gathered
        along its column first, then the columns are exchanged over NVLink."""
        flat = tensor.contiguous().view(-1)
        n, L, N = flat.numel(), self.local_world, self.nnodes
        if out is None:
            out = torch.empty(self.world_size * n, dtype=flat.dtype, device=flat.device)
        if N > 1:
            parts = torch.empty(N * n, dtype=flat.dtype, device=flat.device)        # [node] of my column
            dist.all_gather_into_tensor(parts, flat, group=self.inter)
        else:
            parts = flat
        if L > 1:
            cols = torch.empty(L * N * n, dtype=flat.dtype, device=flat.device)      # [local rank][node]
            comm.allgather(parts, out=cols)
            out.view(N, L, n).copy_(cols.view(L, N, n).permute(1, 0, 2))
        else:
            out.view(-1).copy_(parts)
        return out

    def barrier(self):
        comm.barrier()
        if self.nnodes > 1 and self.local_rank == 0:
            dist.barrier(group=self.inter)
        comm.barrier()

    def finalize(self):
        comm.finalize()


def init_hybrid(inter_backend=None, init_method=None, timeout=None):
    """Initialise this library for the ranks of MY node and torch.distributed for the whole job; returns a HybridComm."""
    rank, world = _env_int("RANK"), _env_int("WORLD_SIZE")
    local_world = _env_int("LOCAL_WORLD_SIZE", world)
    local_rank = _env_int("LOCAL_RANK", rank % local_world)
    if world % local_world:
        raise RuntimeError("init_hybrid: WORLD_SIZE %d is not a multiple of LOCAL_WORLD_SIZE %d" % (world, local_world))
    nnodes, node = world // local_world, rank // local_world
    if rank % local_world != local_rank:
        raise RuntimeError("init_hybrid: ranks must be node-major (RANK = node * LOCAL_WORLD_SIZE + LOCAL_RANK)")
    # the library sees one node: its own rank numbering and a rendezvous name per node
    os.environ["MLSL_RANK"], os.environ["MLSL_WORLD_SIZE"] = str(local_rank), str(local_world)
    os.environ["MLSL_LOCAL_RANK"] = str(local_rank)
    base = os.environ.get("MLSL_JOB_ID") or "hy%s_%s" % (os.environ.get("TORCHELASTIC_RUN_ID", ""), os.environ.get("MASTER_PORT", "0"))
    os.environ["MLSL_JOB_ID"] = "".join(c if c.isalnum() else "_" for c in "%s_n%d" % (base, node))
    comm.init()
    if not dist.is_initialized():
        backend = inter_backend or ("nccl" if comm.is_device() else "gloo")
        kw = {"timeout": timeout} if timeout is not None else {}
        dist.init_process_group(backend, init_method=init_method or "env://", rank=rank, world_size=world, **kw)
    inter = None
    for l in range(local_world):      # every rank takes part in the creation of every column group
        g = dist.new_group([nd * local_world + l for nd in range(nnodes)]) if nnodes > 1 else None
        if l == local_rank:
            inter = g
    return HybridComm(inter, node, nnodes, local_rank, local_world)
