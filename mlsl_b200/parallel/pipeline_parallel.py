"""Pipeline parallelism on collective neighbour exchanges.

The reference partitions a layer's feature maps (model groups) or the minibatch (data groups), never the layer sequence
(reference include/mlsl.hpp:563-620); deep models on NVLink domains also want the third axis.  The library has no
point-to-point primitive, and does not need one: the stages of a pipeline form a group and advance in lock step - at
every clock tick each stage runs (at most) one micro-batch and then ALL stages take part in one `SendRecvList` in which
stage s hands its output to stage s+1 (backward: its input gradient to stage s-1).  One collective per tick keeps the
order of communication identical on all ranks by construction - the rule every collective of the library relies on -
and on the CUDA backend the exchange is a single peer-memory kernel ordered on the compute stream.

Schedules (both give the gradients of the unsplit model, mean over the micro-batches):
  "gpipe" : all forwards, then all backwards; M + S - 1 ticks each way, the activations of all M micro-batches are alive;
  "1f1b"  : a stage runs the forward of micro-batch t - s and the backward of micro-batch t - 2(S-1) + s in the same tick,
            and BOTH directions travel in the one exchange of that tick (activation to s+1, gradient to s-1);
            M + 2(S-1) ticks, at most 2(S-1-s) + 1 micro-batches alive on stage s.
Bubble fraction (S - 1) / (M + S - 1) in both.
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


def _both_ways(send, recv, n_fwd_out, n_bwd_out, n_fwd_in, n_bwd_in, dist, group):
    """One exchange, two directions: send = [activation for s+1 | gradient for s-1], recv = [activation from s-1 |
    gradient from s+1]."""
    g = comm._group(group)
    P, idx = dist.get_process_count(g), dist.get_process_idx(g)
    sc, so, rc, ro = [0] * P, [0] * P, [0] * P, [0] * P
    if idx + 1 < P:
        sc[idx + 1], so[idx + 1] = n_fwd_out, 0
        rc[idx + 1], ro[idx + 1] = n_bwd_in, n_fwd_in
    if idx - 1 >= 0:
        sc[idx - 1], so[idx - 1] = n_bwd_out, n_fwd_out
        rc[idx - 1], ro[idx - 1] = n_fwd_in, 0
    comm._prep(send), comm._prep(recv)
    comm._sync_stream()
    req = dist.send_recv_list(send, sc, so, recv, rc, ro, comm.mlsl_dtype(send.dtype), g)
    comm.Work(comm.env(), req, recv, (send, recv)).wait()


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

    def step(self, micro_inputs=None, loss_fn=None, targets=None, num_micro=None, schedule="gpipe"):
        """One training step over M micro-batches: forward through all stages, backward through all stages.  Gradients
        accumulate into the parameters' .grad (mean over the micro-batches).  Returns the mean loss on the last stage,
        None elsewhere.  `num_micro` must be given on stages that see neither inputs nor targets."""
        M = num_micro if num_micro is not None else len(micro_inputs if micro_inputs is not None else targets)
        S, s = self.stages, self.stage
        if self.is_first and (micro_inputs is None or len(micro_inputs) != M):
            raise ValueError("the first stage needs the %d micro-batch inputs" % M)
        if self.is_last and (loss_fn is None or targets is None or len(targets) != M):
            raise ValueError("the last stage needs loss_fn and the %d micro-batch targets" % M)
        if schedule == "1f1b":
            return self._step_1f1b(micro_inputs, loss_fn, targets, M)
        if schedule != "gpipe":
            raise ValueError("unknown schedule %r (expected 'gpipe' or '1f1b')" % (schedule,))
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

    def _step_1f1b(self, micro_inputs, loss_fn, targets, M):
        S, s = self.stages, self.stage
        n_in = int(torch.Size(self.in_shape).numel())
        n_out = int(torch.Size(self.out_shape).numel())
        # send = [my output | my input gradient], recv = [my input | my output gradient]
        send, recv = self._buf((n_out + n_in,)), self._buf((n_in + n_out,))
        saved, total = {}, None
        self.max_alive = 0
        for t in range(M + 2 * (S - 1)):
            f, b = t - s, t - 2 * (S - 1) + s
            if 0 <= f < M:
                x = micro_inputs[f] if self.is_first else recv[:n_in].view(self.in_shape).clone().requires_grad_(True)
                y = self.module(x)
                saved[f] = (x, y)
                self.max_alive = max(self.max_alive, len(saved))
                if not self.is_last:
                    send[:n_out].copy_(y.detach().reshape(-1))
            if 0 <= b < M:
                x, y = saved.pop(b)
                if self.is_last:
                    loss = loss_fn(y, targets[b]) / M
                    total = loss.detach() if total is None else total + loss.detach()
                    loss.backward()
                else:
                    torch.autograd.backward(y, recv[n_in:].view_as(y).clone())
                if not self.is_first:
                    send[n_out:].copy_(x.grad.reshape(-1))
            if S > 1:
                _both_ways(send, recv, n_out, n_in, n_in, n_out, self.dist, self.group)
        if comm.is_device():
            comm.free_tensor(send)
            comm.free_tensor(recv)
        return total


def bubble_fraction(stages, num_micro):
    return (stages - 1) / float(num_micro + stages - 1)


__all__ = ["PipelineStage", "bubble_fraction", "GroupType"]
