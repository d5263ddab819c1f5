"""DistributedOptimizer: data-parallel training on top of the Session / ParameterSet graph API.

What a framework integration of the reference does by hand (reference tests/examples/mlsl_test/mlsl_test.cpp:
per layer StartGradientComm during backward, WaitGradientComm before the update, optional distributed update with
StartIncrementComm / WaitIncrementComm) packaged for PyTorch:

  * parameters and gradients are flattened into buckets that live in the symmetric heap (zero-copy for the peer
    kernels); every parameter's .data / .grad is a view into its bucket;
  * one ParameterSet per bucket is registered in a Session (one Operation per bucket, so statistics are per bucket);
  * a post-accumulate-grad hook counts the gradients of a bucket and starts its communication the moment the last
    one is produced - while autograd is still computing the earlier layers (the overlap MLSL was designed for);
  * mode "fused" (default on the CUDA backend): ReduceScatter + optimizer step on the owned shard + AllGather of the
    new parameters is ONE kernel per bucket (ParameterSet.start_fused_update), with fp32 master weights and optimizer
    state sharded over the data group (the reference's "distributed update", src/mlsl_impl.cpp:401-433);
  * mode "allreduce": gradient AllReduce with the 1/N scale (and optionally the fp8 transport) fused in, followed by
    a local torch optimizer.
"""
import torch

from . import comm
from .api import CompressionType, OperationType, OptimizerType


class _Bucket:
    __slots__ = ("params", "numel", "padded", "grad", "flat", "ps", "op", "pending", "master", "state1", "state2",
                 "started", "offsets", "accumulating")


class DistributedOptimizer:
    def __init__(self, params, lr=0.1, momentum=0.0, weight_decay=0.0, optimizer="sgd", betas=(0.9, 0.999), eps=1e-8,
                 bucket_mb=64, mode=None, compress=False, distribution=None, average=True, hybrid=None):
        self.env = comm.env()
        self._state = comm._state()   # gradient hooks fire on autograd's device thread
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        self.dist = distribution if distribution is not None else comm.world_distribution()
        self.world = self.dist.get_process_count(0)
        self.rank = self.dist.get_process_idx(0)
        # hybrid: a parallel.multinode.HybridComm - gradients are averaged over ALL nodes by the two-level all-reduce
        # (reduce-scatter in the node, shard all-reduce between nodes, all-gather in the node), the update is local
        self.hybrid = hybrid
        self.mode = mode or ("allreduce" if hybrid is not None else "fused")
        assert self.mode in ("fused", "allreduce")
        assert hybrid is None or self.mode == "allreduce", "hybrid (multi-node) training uses mode='allreduce'"
        self.kind = optimizer
        assert optimizer in ("sgd", "adamw")
        self.lr, self.momentum, self.weight_decay, self.betas, self.eps = lr, momentum, weight_decay, betas, eps
        self.compress = bool(compress) and self.mode == "allreduce"
        self.scale = 1.0 / (hybrid.world_size if hybrid is not None else self.world) if average else 1.0
        self.steps = 0
        self._sync = True
        self._build(bucket_mb)
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(p)) for p in self.params]
        if self.mode == "allreduce":
            cls = torch.optim.AdamW if optimizer == "adamw" else torch.optim.SGD
            kw = dict(lr=lr, weight_decay=weight_decay)
            kw.update(dict(betas=betas, eps=eps) if optimizer == "adamw" else dict(momentum=momentum))
            self.local = cls(self.params, **kw)

    # ------------------------------------------------------------------------------------------------------------
    def _build(self, bucket_mb):
        dtype = self.params[0].dtype
        assert all(p.dtype == dtype for p in self.params), "mixed parameter dtypes: create one optimizer per dtype"
        self.dtype = dtype
        limit = int(bucket_mb * (1 << 20)) // self.params[0].element_size()
        # buckets are filled in REVERSE parameter order: gradients arrive last-layer-first
        groups, cur, cur_n = [], [], 0
        for p in reversed(self.params):
            if cur and cur_n + p.numel() > limit:
                groups.append(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            groups.append(cur)
        self.session = self.env.create_session()
        self.session.set_global_minibatch_size(self.world)
        mdt = comm.mlsl_dtype(dtype)
        self.buckets, self._bucket_of, self._offset_of = [], {}, {}
        align = 64  # elements: keeps every parameter view 16-byte aligned and the owned shards vector friendly
        for gi, plist in enumerate(groups):
            b = _Bucket()
            b.params, b.offsets, off = plist, [], 0
            for p in plist:
                b.offsets.append(off)
                off += (p.numel() + align - 1) // align * align
            b.numel = off
            b.padded = (off + self.world * align - 1) // (self.world * align) * (self.world * align)
            ri = self.session.create_operation_reg_info(OperationType.CC)
            ri.set_name("bucket_%d" % gi)
            ri.add_input(1, 1, mdt)
            ri.add_output(1, 1, mdt)
            ri.add_parameter_set(b.padded, 1, mdt, self.mode == "fused",
                                 CompressionType.QUANTIZATION if self.compress else CompressionType.NONE)
            b.op = self.session.get_operation(self.session.add_operation(ri, self.dist))
            self.session.delete_operation_reg_info(ri)
            self.buckets.append(b)
        self.session.commit()
        dev = self.params[0].device
        for b in self.buckets:
            b.ps = b.op.get_parameter_set(0)
            b.grad = comm.alloc_tensor(b.padded, dtype)
            b.flat = comm.alloc_tensor(b.padded, dtype)
            for p, off in zip(b.params, b.offsets):
                view = b.flat[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = b.grad[off:off + p.numel()].view_as(p)
                self._bucket_of[p] = b
                self._offset_of[p] = off
            b.pending, b.started, b.accumulating = len(b.params), False, False
            b.master = b.state1 = b.state2 = None
            if self.mode == "fused":
                owned = b.ps.get_owned_kernel_count() * b.ps.get_kernel_size()
                lo = b.ps.get_owned_kernel_offset() * b.ps.get_kernel_size()
                if dtype != torch.float32:
                    b.master = b.flat[lo:lo + owned].float().clone()
                if self.kind == "adamw" or self.momentum != 0.0:
                    b.state1 = torch.zeros(owned, dtype=torch.float32, device=dev)
                if self.kind == "adamw":
                    b.state2 = torch.zeros(owned, dtype=torch.float32, device=dev)
            else:
                b.ps.set_gradient_scale(self.scale)

    def _reattach(self, p, b):
        """`model.zero_grad()` (set_to_none=True by default) or `p.grad = None` makes autograd allocate a fresh gradient
        outside the bucket; the exchange would then ship a stale bucket and the replicas would silently diverge.  Move
        the gradient into the bucket (accumulating under no_sync, where the bucket holds earlier micro-batches) and
        point p.grad at its bucket view again."""
        off = self._offset_of[p]
        view = b.grad[off:off + p.numel()].view_as(p)
        g = p.grad
        if g is not None and g.data_ptr() != view.data_ptr():
            if getattr(b, "accumulating", False):
                view.add_(g.to(view.dtype))
            else:
                view.copy_(g)
            p.grad = view

    def _make_hook(self, p):
        def hook(_):
            b = self._bucket_of[p]
            self._reattach(p, b)
            b.pending -= 1
            if b.pending == 0:
                if not self._sync:          # gradient accumulation: keep adding into the bucket, exchange later
                    b.pending = len(b.params)
                    b.accumulating = True
                    return
                b.accumulating = False
                with comm.use_state(self._state):
                    self._start(b)
        return hook

    def no_sync(self):
        """Context manager for gradient accumulation: backward passes inside it only accumulate into the buckets; the
        exchange of the accumulated gradients starts in the first backward after it (or in step())."""
        opt = self

        class _NoSync:
            def __enter__(self):
                opt._sync = False

            def __exit__(self, *exc):
                opt._sync = True
                return False

        return _NoSync()

    def _start(self, b):
        if self.mode == "fused":
            b.ps.start_fused_update(b.grad, b.flat, comm.mlsl_dtype(self.dtype), b.master, b.state1, b.state2,
                                    OptimizerType.ADAMW if self.kind == "adamw" else OptimizerType.SGD, lr=self.lr,
                                    momentum=self.momentum, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps,
                                    weight_decay=self.weight_decay, step=self.steps + 1, grad_scale=self.scale)
        elif self.hybrid is not None:
            self.hybrid.allreduce(b.grad, scale=self.scale)
        else:
            b.ps.start_gradient_comm(b.grad)
        b.started = True

    # ------------------------------------------------------------------------------------------------------------
    def step(self):
        """Finish the communication started during backward and apply the update."""
        for b in self.buckets:
            if not b.started:          # gradients that autograd never produced (unused parameters): start now
                self._start(b)
        for b in self.buckets:
            if self.mode == "fused":
                b.ps.wait_fused_update()
            elif self.hybrid is None:
                b.ps.wait_gradient_comm()
            b.pending, b.started = len(b.params), False
        if self.mode == "allreduce":
            self.local.step()
        self.steps += 1

    def zero_grad(self, set_to_none=False):
        """Zero the gradient buckets and (re)point every p.grad at its bucket view - also after a `model.zero_grad()`
        that set the gradients to None."""
        for b in self.buckets:
            b.grad.zero_()
            b.accumulating = False
            for p, off in zip(b.params, b.offsets):
                view = b.grad[off:off + p.numel()].view_as(p)
                if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                    p.grad = view

    def set_lr(self, lr):
        self.lr = lr
        if self.mode == "allreduce":
            for g in self.local.param_groups:
                g["lr"] = lr

    def gather_full_state(self):
        """Checkpoint helper (the reference leaves snapshot assembly to Distribution::AllGather): returns
        the full fp32 master weights of every bucket on every rank."""
        out = []
        for b in self.buckets:
            if self.mode == "fused" and b.master is not None:
                out.append(comm.allgather(b.master.contiguous(), distribution=self.dist))
            else:
                out.append(b.flat.float().clone())
        return out

    # ---- checkpoint / resume (the reference has none; SURVEY 5.4 leaves snapshot assembly to AllGather) ------------
    def state_dict(self):
        """Collective.  The COMPLETE optimizer state on every rank: fp32 master weights and moments are all-gathered
        from their owners and stripped of the world-size dependent padding, so a checkpoint written by one rank can be
        resumed on a different number of ranks (the shards are re-cut in load_state_dict)."""
        sd = {"version": 1, "mode": self.mode, "optimizer": self.kind, "steps": self.steps, "lr": self.lr, "buckets": []}
        for b in self.buckets:
            e = {"numel": b.numel, "shapes": [tuple(p.shape) for p in b.params]}

            def full(shard):
                if shard is None:
                    return None
                return comm.allgather(shard.contiguous(), distribution=self.dist)[:b.numel].float().cpu().clone()

            if self.mode == "fused":
                e["master"] = full(b.master) if b.master is not None else b.flat[:b.numel].float().cpu().clone()
                e["state1"], e["state2"] = full(b.state1), full(b.state2)
            else:
                e["master"] = b.flat[:b.numel].float().cpu().clone()
            sd["buckets"].append(e)
        if self.mode == "allreduce":
            import copy
            sd["local"] = copy.deepcopy(self.local.state_dict())   # torch hands out references to the live moments
        return sd

    def load_state_dict(self, sd):
        """Restore parameters (they are views of the buckets) and optimizer state; works across world sizes."""
        assert sd.get("version") == 1 and sd["mode"] == self.mode and sd["optimizer"] == self.kind, \
            "checkpoint was written by a different optimizer configuration"
        assert len(sd["buckets"]) == len(self.buckets), "different bucket layout (bucket_mb / parameter list changed)"
        self.steps = int(sd["steps"])
        self.set_lr(sd["lr"])
        dev = self.params[0].device
        for b, e in zip(self.buckets, sd["buckets"]):
            assert e["numel"] == b.numel and e["shapes"] == [tuple(p.shape) for p in b.params], "parameter shapes changed"
            with torch.no_grad():
                b.flat[:b.numel].copy_(e["master"].to(self.dtype))
            if self.mode != "fused":
                continue
            owned = b.ps.get_owned_kernel_count() * b.ps.get_kernel_size()
            lo = b.ps.get_owned_kernel_offset() * b.ps.get_kernel_size()

            def shard(full):
                if full is None:
                    return None
                padded = torch.zeros(b.padded, dtype=torch.float32)
                padded[:b.numel] = full
                return padded[lo:lo + owned].to(dev).contiguous()

            if b.master is not None:
                b.master = shard(e["master"])
            b.state1 = shard(e["state1"]) if b.state1 is not None else None
            b.state2 = shard(e["state2"]) if b.state2 is not None else None
        if self.mode == "allreduce" and "local" in sd:
            import copy
            self.local.load_state_dict(copy.deepcopy(sd["local"]))   # torch keeps references when dtype/device match

    def close(self):
        """Release the session and the buckets.  The parameters get storage of their own back (they were views into the
        buckets), so the model stays usable - e.g. for another optimizer."""
        for h in self._hooks:
            h.remove()
        self.env.delete_session(self.session)
        with torch.no_grad():
            for p in self.params:
                p.data = p.data.clone()
                p.grad = None
        for b in self.buckets:
            comm.free_tensor(b.grad)
            comm.free_tensor(b.flat)
