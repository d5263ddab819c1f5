"""Pipeline parallelism on collective neighbour exchanges.

The reference partitions a layer's feature maps (model groups) or the minibatch (data groups), never the layer sequence
(reference include/mlsl.hpp:563-620); deep models on NVLink domains also want the third axis.  The library has no
point-to-point primitive, and does not need one: the stages of a pipeline form a group and advance in lock step - at
every clock tick each stage runs (at most) one micro-batch and then ALL stages take part in one `SendRecvList` in which
stage s hands its output to stage s+1 (backward: its input gradient to stage s-1).  One collective per tick keeps the
order of communication identical on all ranks by construction - the rule every collective of the library relies on -
and on the CUDA backend the exchange is a single peer-memory kernel ordered on the compute stream.

Schedule: GPipe (all forwards, then all backwards), M micro-batches over S stages = M + S - 1 ticks each way, bubble
fraction (S - 1) / (M + S - 1); activations of the M micro-batches are kept for the backward pass.
"""
import torch

from .. import comm
from ..api import GroupType


def _neighbour_shift(tensor, out, direction, dist, group):
    """out on stage i = tensor of stage i - direction; the ends of the pipeline send / receive nothing."""
    g = comm._group(group)
    P, idx = dist.get_process_count(g), dist.get_process_idx(g)
    n = tensor.numel()
    dst, src = idx + direction, idx - direction
    sc = [n if p == dst else 0 for p in range(P)]
    rc = [n if p == src else 0 for p in range(P)]
    comm._prep(tensor), comm._prep(out)
    comm._sync_stream()
    req = dist.send_recv_list(tensor, sc, [0] * P, out, rc, [0] * P, comm.mlsl_dtype(tensor.dtype), g)
    comm.Work(comm.env(), req, out, (tensor, out)).wait()
    return out


class PipelineStage:
    """One stage of a pipeline: `module` maps an activation of shape `in_shape` (per micro-batch) to `out_shape`.

        stage = PipelineStage(my_layers, in_shape=(mb, d), out_shape=(mb, d), group="model", distribution=dist)
        loss = stage.step(micro_inputs if stage.is_first else None, loss_fn=f if stage.is_last else None,
                          targets=micro_targets if stage.is_last else None)      # gradients are in .grad afterwards
    """

    def __init__(self, module, in_shape, out_shape, dtype=torch.float32, group="model", distribution=None, device=None):
        self.module, self.group = module, group
        self.dist = distribution if distribution is not None else comm.world_distribution()
        g = comm._group(group)
        self.stages, self.stage = self.dist.get_process_count(g), self.dist.get_process_idx(g)
        self.is_first, self.is_last = self.stage == 0, self.stage == self.stages - 1
        self.in_shape, self.out_shape, self.dtype = tuple(in_shape), tuple(out_shape), dtype
        if device is None:
            p = next(module.parameters(), None)
            device = p.device if p is not None else torch.device("cpu")
        self.device = device

    def _buf(self, shape):
        if comm.is_device():
            return comm.alloc_tensor(shape, self.dtype, zero=True)
        return torch.zeros(shape, dtype=self.dtype, device=self.device)

    def step(self, micro_inputs=None, loss_fn=None, targets=None, num_micro=None):
        """One training step over M micro-batches: forward through all stages, backward through all stages.  Gradients
        accumulate into the parameters' .grad (mean over the micro-batches).  Returns the mean loss on the last stage,
        None elsewhere.  `num_micro` must be given on stages that see neither inputs nor targets."""
        M = num_micro if num_micro is not None else len(micro_inputs if micro_inputs is not None else targets)
        S, s = self.stages, self.stage
        if self.is_first and (micro_inputs is None or len(micro_inputs) != M):
            raise ValueError("the first stage needs the %d micro-batch inputs" % M)
        if self.is_last and (loss_fn is None or targets is None or len(targets) != M):
            raise ValueError("the last stage needs loss_fn and the %d micro-batch targets" % M)
        recv_act, send_act = self._buf(self.in_shape), self._buf(self.out_shape)
        recv_grad, send_grad = self._buf(self.out_shape), self._buf(self.in_shape)
        saved = [None] * M
        total = None
        # ---- forward: at tick t stage s holds micro-batch t - s ----
        for t in range(M + S - 1):
            m = t - s
            if 0 <= m < M:
                if self.is_first:
                    x = micro_inputs[m]
                else:
                    x = recv_act.clone().requires_grad_(True)
                y = self.module(x)
                saved[m] = (x, y)
                if not self.is_last:
                    send_act.copy_(y.detach())
            if S > 1:
                _neighbour_shift(send_act, recv_act, +1, self.dist, self.group)
        # ---- backward: the last stage starts with the last micro-batch; at tick t stage s holds M-1-(t-(S-1-s)) ----
        for t in range(M + S - 1):
            m = M - 1 - (t - (S - 1 - s))
            if 0 <= m < M:
                x, y = saved[m]
                if self.is_last:
                    loss = loss_fn(y, targets[m]) / M
                    total = loss.detach() if total is None else total + loss.detach()
                    loss.backward()
                else:
                    torch.autograd.backward(y, recv_grad.view_as(y).clone())
                if not self.is_first:
                    send_grad.copy_(x.grad)
                saved[m] = None
            if S > 1:
                _neighbour_shift(send_grad, recv_grad, -1, self.dist, self.group)
        if comm.is_device():
            for b in (recv_act, send_act, recv_grad, send_grad):
                comm.free_tensor(b)
        return total


def bubble_fraction(stages, num_micro):
    return (stages - 1) / float(num_micro + stages - 1)


__all__ = ["PipelineStage", "bubble_fraction", "GroupType"]
